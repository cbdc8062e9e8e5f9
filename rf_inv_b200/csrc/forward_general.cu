// forward_kernel variants for transform lengths that are not powers of two (Bluestein's chirp-z form of the inverse FFT,
// forward.cu: bluestein_inverse) -- a translation unit of its own so the two sets of instantiations compile side by side.
#define RFINV_FWD_GENERAL_TU 1
#include "forward.cu"
