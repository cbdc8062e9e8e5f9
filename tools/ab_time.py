"""A/B kernel timing of several builds of the library in ONE process, interleaved (box-to-box clock differences cancel):
    python tools/ab_time.py [--workload target] [--chains 16384] [--rounds 12] lib1.so lib2.so ...
Device-resident inputs (rfinv_eval_batch_device), CUDA-event times of prep+forward / quadratic form, L2 flushed between calls."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from rf_inv_b200 import capi, workloads
from rf_inv_b200 import evaluator as evmod

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--workload", default="target")
ap.add_argument("--chains", type=int, default=16384)
ap.add_argument("--rounds", type=int, default=12)
ap.add_argument("--dvs-scale", type=float, default=0.3)
ap.add_argument("--no-inner-timing", action="store_true", help="no CUDA events between the kernels of an evaluation")
a = ap.parse_args()
cfg = workloads.make_config(a.workload)
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)
m = workloads.draw_models(cfg, a.chains, seed=100, dvs_scale=a.dvs_scale)
dev = torch.device("cuda", 0)
d = {k: torch.from_numpy(v).to(dev) for k, v in workloads.to_soa(m).items()}
logl = torch.empty(a.chains, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
evs = []
for so in a.libs:
    lib = capi.load(so)
    capi._lib = lib                       # Evaluator picks the library up at construction
    ev = evmod.Evaluator(cfg)
    ev.set_stream(torch.cuda.current_stream().cuda_stream)
    capi.check(lib.rfinv_set_timing(ev.handle, 0 if a.no_inner_timing else 1))
    evs.append((so, lib, ev))
res = {so: [] for so in a.libs}
ref = None
for r in range(a.rounds + 2):
    for so, lib, ev in evs:
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ev.calc_likelihood_device(a.chains, d["k"].data_ptr(), d["z"].data_ptr(), d["dvp"].data_ptr(), d["dvs"].data_ptr(),
                                  d["sig"].data_ptr(), logl.data_ptr())
        e1.record(); e1.synchronize()
        t = (C.c_double * 3)()
        if not a.no_inner_timing: capi.check(lib.rfinv_get_timing(ev.handle, t))
        if r >= 2: res[so].append([t[0], t[1], t[2], e0.elapsed_time(e1)])
        out = logl.cpu().numpy()
        if ref is None: ref = out
        elif r == 0: print(f"{so}: max |logL - logL(first lib)| / |logL| = {np.nanmax(np.abs(out - ref) / np.abs(ref)):.2e}")
print(f"workload {a.workload}, {a.chains} chains, k_mean {m['k'].mean():.2f}; median (min) ms over {a.rounds} interleaved rounds")
for so in a.libs:
    x = np.array(res[so]); med = np.median(x, axis=0); mn = np.min(x, axis=0)
    print(f"{os.path.basename(so):32s} prep+forward {med[0]:.4f} ({mn[0]:.4f})  quadform {med[1]:.4f} ({mn[1]:.4f})  step {med[3]:.4f} ({mn[3]:.4f})  -> {a.chains / med[3] / 1e3:.3f} M evals/s")
