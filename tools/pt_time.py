"""PT-MCMC iterations/s at the target shape (16384 chains on one GPU): python tools/pt_time.py [iters] [lib.so]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rf_inv_b200 import capi, workloads
if len(sys.argv) > 2:
    capi._lib = capi.load(sys.argv[2])
from rf_inv_b200.evaluator import Evaluator
from rf_inv_b200.pt import ParallelTempering
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = workloads.make_config(os.environ.get("WORKLOAD", "target"))
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)
tm = workloads.true_model(cfg)
with Evaluator(cfg) as ev:
    _, rft, _ = ev.calc_likelihood(tm["k"], tm["z"], tm["dvp"], tm["dvs"], tm["sig"], want_rft=True)
cfg.obs = (rft[0, :, :cfg.nsmp] + np.random.default_rng(7).normal(0.0, 0.01, (cfg.ntrc, cfg.nsmp))).astype(np.float32).astype(np.float64)
pt = ParallelTempering(cfg, int(os.environ.get("CHAINS", "16384")) // cfg.nchains)
pt.run(10)
t0 = time.perf_counter(); pt.run(iters); dt = time.perf_counter() - t0
print(f"{iters} iterations: {iters / dt:.1f} iters/s, {dt / iters * 1e6:.1f} us per iteration, k_mean {pt.state()['k'].mean():.2f}, n_eval {pt.counters()['n_eval']}")
pt.close()
