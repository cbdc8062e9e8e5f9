"""First GPU parity probe (not a test): CUDA path vs C oracle on a few small configurations."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers, oracle_c
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator

res = {}
cases = {
    "land_P": dict(), "sea_P": dict(sdep=2.0), "land_S": dict(ipha=[-1, -1], rayps=[0.10, 0.12]),
    "sea_S_deconv": dict(sdep=1.0, ipha=[-1, -1], deconv_mode=1), "P_deconv": dict(deconv_mode=1),
    "common": dict(rayps=[0.06, 0.06], a_gus=[2.0, 4.0]), "vp1_tstart": dict(vp_mode=1, t_start=-3.0),
    "n1024": dict(nfft=1024, nsmp=512, k_max=20, z_max=40.0), "n2048": dict(nfft=2048, nsmp=1000, k_max=30, z_max=40.0),
    "n128": dict(nfft=128, nsmp=64), "n64": dict(nfft=64, nsmp=40),
    "n512": dict(nfft=512, nsmp=200),
}
for name, kw in cases.items():
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(**kw))
    m = workloads.draw_models(cfg, 64, seed=3, dvs_scale=0.3)
    t0 = time.time()
    ll_o, rft_o, val_o = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    t1 = time.time()
    with Evaluator(cfg) as ev:
        ll_g, rft_g, val_g = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True, want_valid=True)
    e_rft = helpers.rel_err_rft(rft_g, rft_o)
    e_ll = float(np.max(np.abs(ll_g - ll_o) / np.abs(ll_o)))
    res[name] = dict(rft=e_rft, logl=e_ll, valid_equal=bool(np.array_equal(val_g, val_o)), oracle_s=t1 - t0)
    print(name, res[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "first_check.json"), "w"), indent=1)
