"""Experiment helper: time prep+forward / quadform with an alternative build of the library."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rf_inv_b200 import capi, workloads
so = sys.argv[1]
capi._lib = capi.load(so)
from rf_inv_b200.evaluator import Evaluator
cfg = workloads.make_config("target")
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)   # the real R^-1: the quadratic form may use its factor
nC = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
m = workloads.draw_models(cfg, nC, seed=100, dvs_scale=0.3)
with Evaluator(cfg) as ev:
    capi.check(capi._lib.rfinv_set_timing(ev.handle, 1))
    ts = []
    for i in range(8):
        ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
        t = (C.c_double * 3)(); capi.check(capi._lib.rfinv_get_timing(ev.handle, t)); ts.append(list(t))
print(so, "forward/quadform/loglik ms:", np.round(np.mean(ts[3:], axis=0), 4))
