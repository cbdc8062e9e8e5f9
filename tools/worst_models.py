"""Diagnostics: the models of a parity-report configuration on which the CUDA path differs most from the C oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers, oracle_c
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cfg = helpers.attach_obs_and_rinv(workloads.make_config(name), noise=0.01)
m = workloads.draw_models(cfg, n, seed=2024, dvs_scale=0.5)
ll_o, rft_o, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
with Evaluator(cfg) as ev:
    ll_g, rft_g, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
scale = np.max(np.abs(rft_o), axis=-1)
err = np.max(np.abs(rft_g - rft_o), axis=-1) / scale          # [model][trace]
order = np.argsort(err.max(axis=1))[::-1]
for i in order[:8]:
    print(int(i), "k", int(m["k"][i]), "err per trace", ["%.2e" % e for e in err[i]])
print("models above 1e-10:", int((err.max(axis=1) > 1e-10).sum()), "above 1e-12:", int((err.max(axis=1) > 1e-12).sum()), "of", n)
np.save(os.path.join(ROOT, "gpurun_out", f"worst_{name}.npy"), err)
