"""Times the blocking rfinv_eval_batch (host buffers, pinned) at the target shape: python tools/e2e_sync_time.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
cfg = workloads.make_config("target")
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)
m = workloads.draw_models(cfg, 16384, seed=100, dvs_scale=0.3)
p = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in m.items()}
with Evaluator(cfg) as ev:
    for _ in range(3): ev.calc_likelihood(p["k"], p["z"], p["dvp"], p["dvs"], p["sig"])
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); ev.calc_likelihood(p["k"], p["z"], p["dvp"], p["dvs"], p["sig"]); ts.append(time.perf_counter() - t0)
print("pieces", os.environ.get("RFINV_UPLOAD_PIECES_N", "4"), "overlap", os.environ.get("RFINV_UPLOAD_OVERLAP", "1"),
      "median ms %.4f min %.4f -> %.3f M evals/s" % (np.median(ts) * 1e3, np.min(ts) * 1e3, 16384 / np.median(ts) / 1e6))
