"""prep+forward / quadform device time of the target shape with its transform length changed: powers of two run the
shared-memory FFT directly, any other length Bluestein's convolution on twice the next power of two.
    python tools/any_length_time.py [n_models] [nfft ...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rf_inv_b200 import capi, workloads
from rf_inv_b200.evaluator import Evaluator
nC = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lengths = [int(a) for a in sys.argv[2:]] or [1024, 1000, 2048, 2000]
for nfft in lengths:
    cfg = workloads.make_config("target")
    cfg.nfft = nfft; cfg.nsmp = min(512, nfft)
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)
    m = workloads.draw_models(cfg, nC, seed=100, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        capi.check(capi.load().rfinv_set_timing(ev.handle, 1))
        ts = []
        for i in range(6):
            ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
            t = (C.c_double * 3)(); capi.check(capi.load().rfinv_get_timing(ev.handle, t)); ts.append(list(t))
    f, q, _ = np.mean(ts[2:], axis=0)
    print(f"nfft {nfft:5d}: prep+forward {f:8.4f} ms  quadform {q:7.4f} ms  -> {nC / (f + q) * 1e3:.3e} evals/s ({nC} models x {cfg.ntrc} traces)")
