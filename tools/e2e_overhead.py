"""Where the end-to-end call spends its time: Evaluator.calc_likelihood (Python) vs rfinv_eval_batch (C) vs the device step."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from rf_inv_b200 import capi, workloads
from rf_inv_b200.evaluator import Evaluator, _p
cfg = workloads.make_config("target")
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)
n = 16384
m = workloads.draw_models(cfg, n, seed=100, dvs_scale=0.3)
pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in m.items()}
logl = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
lib = capi.load()
with Evaluator(cfg) as ev:
    args = (ev.handle, n, _p(pin["k"], capi.i32p), _p(pin["z"], capi.dp), _p(pin["dvp"], capi.dp), _p(pin["dvs"], capi.dp),
            _p(pin["sig"], capi.dp), _p(logl, capi.dp), C.cast(None, capi.dp), C.cast(None, capi.u8p))
    for name, fn in (("python calc_likelihood", lambda: ev.calc_likelihood(pin["k"], pin["z"], pin["dvp"], pin["dvs"], pin["sig"])),
                     ("C rfinv_eval_batch (pinned logl)", lambda: lib.rfinv_eval_batch(*args))):
        for _ in range(3): fn()
        ts = []
        for _ in range(30):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        print(f"{name}: {np.mean(ts) * 1e3:.4f} ms (min {np.min(ts) * 1e3:.4f})")
