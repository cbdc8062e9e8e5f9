"""Debug tool: builds librfinv_b200_prof.so with -DRFINV_PHASE_TIMING and prints the average cycles a
forward_kernel CTA spends per phase (thread 0's clock64 between phase boundaries)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rf_inv_b200 import build as rbuild, capi, workloads

so = os.path.join(ROOT, "rf_inv_b200", "librfinv_b200_prof.so")
src = [os.path.join(rbuild.CSRC, s) for s in rbuild.SOURCES]
if "--build" in sys.argv or not os.path.exists(so):
    flags = [f for f in rbuild.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    subprocess.check_call([rbuild.nvcc_path(), "-ccbin", "/usr/bin/g++"] + flags + ["-DRFINV_PHASE_TIMING", "-shared", "-o", so] + src + ["-lcudart", "-ldl"])
    if "--build" in sys.argv:
        sys.exit(0)
lib = capi.load(so)
capi._lib = lib
from rf_inv_b200.evaluator import Evaluator
wl = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "target"
cfg = workloads.make_config(wl)
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = np.zeros((cfg.ntrc, cfg.nsmp, cfg.nsmp))
# --k N: every model with N interfaces (e.g. --k 3: the layer counts of the chains inside the PT loop, k_mean 2.5 - 3)
k_fixed = int(sys.argv[sys.argv.index("--k") + 1]) if "--k" in sys.argv else None
m = workloads.draw_models(cfg, 4096, seed=100, dvs_scale=0.3, k_fixed=k_fixed)
names = ["stage consts", "trig tables", "layer loop", "epilogue(surface)", "deconv+Z build", "fft", "max/shift/output"]
with Evaluator(cfg) as ev:
    ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    out = (C.c_ulonglong * 16)()
    lib.rfinv_debug_get_phases(out)
    ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    lib.rfinv_debug_get_phases(out)
n = out[15]
tot = sum(out[i] for i in range(7))
print(f"CTAs {n}, mean cycles per CTA {tot / n:.0f}, k_mean {m['k'].mean():.2f}")
for i, nm in enumerate(names):
    print(f"  {nm:20s} {out[i] / n:9.0f} cycles  {100.0 * out[i] / tot:5.1f}%")
pnames = ["load + sort (warp 0)", "shared layer physics (warp 0)", "barrier + per-ray physics", "interfaces + scans", "ray constants", "edge bins (serial, warp 0)", "edge surface + store"]
ptot = sum(out[8 + i] for i in range(7))
print(f"prep_kernel, warp-cycles per model {ptot / len(m['k']):.0f} (summed over the warps of a CTA)")
for i, nm in enumerate(pnames):
    print(f"  {nm:32s} {out[8 + i] / len(m['k']):9.0f} cycles  {100.0 * out[8 + i] / ptot:5.1f}%")
