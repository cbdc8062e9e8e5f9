// Parallel-tempering MCMC on device (filled in below).
#include "rfinv_handle.h"

void rfinv_handle::free_pt() {}
