"""Parity report quoted with every benchmark: max relative error of rft and logL of the CUDA path vs the
C oracle on >= 1000 random models per configuration, and fixed-seed PT sequence identity on C1.  Run on the GPU box:
    python tools/parity_report.py > gpurun_out/parity_report.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers, oracle_c
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
from rf_inv_b200.pt import ParallelTempering

out = {"tolerance": 1e-9, "configs": {}}
# "<workload>@<nfft>": the workload with its transform length changed to one that is not a power of two (Bluestein path)
for name, n in [("sample", 8192), ("c2", 8192), ("c3", 8192), ("c3_buried", 8192), ("c3_deconv", 8192), ("c3_common", 8192), ("c4", 8192), ("c4_laplace", 4096), ("c5", 4096), ("target", 8192),
                ("target@1000", 2048), ("c3@600", 2048), ("c5@2000", 1024), ("sample@375", 4096)]:
    base = workloads.make_config(name.split("@")[0])
    if "@" in name:
        base.nfft = int(name.split("@")[1]); base.nsmp = min(base.nsmp, base.nfft // 2)
    cfg = helpers.attach_obs_and_rinv(base, noise=0.01)
    m = workloads.draw_models(cfg, n, seed=2024, dvs_scale=0.5)
    t0 = time.time()
    ll_o, rft_o, val_o, cond = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_cond=True)
    t_cpu = time.time() - t0
    with Evaluator(cfg) as ev:
        ll_g, rft_g, val_g = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True, want_valid=True)
    out["configs"][name] = {"models": n, "nfft": cfg.nfft, "ntrc": cfg.ntrc, "k_max": cfg.k_max, "nsmp": cfg.nsmp,
                            "rft_max_rel_err": helpers.rel_err_rft(rft_g, rft_o),
                            "logl_max_rel_err": helpers.logl_err(cfg, ll_g, ll_o, m["sig"]),
                            "valid_flags_equal": bool(np.array_equal(val_g, val_o)), "nan_logl": int(np.isnan(ll_g).sum()),
                            "oracle_seconds": round(t_cpu, 2)}
    # Conditioning of the reference's normalisation (rft = rx / maxval(rx_vertical), src/forward.f90:197-203): where the main
    # pulse of the vertical trace is negative, maxval() is a small ripple and any two fp64 evaluations of the reference's own
    # formulas only agree to eps * cond, cond = max|rx| / maxval(rx) (oracle/rfinv_oracle.c orc_eval_batch_cond).
    e_mt = np.max(np.abs(rft_g - rft_o), axis=-1) / np.max(np.abs(rft_o), axis=-1)          # [model][trace]
    well = cond <= 1.0e3
    out["configs"][name].update({
        "norm_condition_max": float(np.max(cond)), "traces_with_norm_condition_above_1e3": int((~well).sum()),
        "rft_max_rel_err_where_condition_le_1e3": float(np.max(e_mt[well])),
        "rft_max_rel_err_over_condition": float(np.max(e_mt / cond)),
        "logl_max_rel_err_where_condition_le_1e3": helpers.logl_err(cfg, ll_g[well.all(axis=1)], ll_o[well.all(axis=1)], m["sig"][well.all(axis=1)])})
cfg = helpers.attach_obs_and_rinv(workloads.make_config("sample"), noise=0.01)
n_iter, nproc = 3000, 20
pt = ParallelTempering(cfg, nproc); pt.set_logging(n_iter); pt.run(n_iter)
fl, ty, sw = pt.log(n_iter); cnt = pt.counters(); st = pt.state(); pt.close()
orc = oracle_c.OraclePT(cfg, nproc); ofl, oty, osw = orc.run(n_iter); ocnt = orc.counters(n_iter); ost = orc.state()
out["pt_fixed_seed_C1"] = {"iterations": n_iter, "chains": nproc * cfg.nchains,
                           "accept_flags_identical": bool(np.array_equal(fl, ofl)), "proposal_types_identical": bool(np.array_equal(ty, oty)),
                           "swaps_identical": bool(np.array_equal(sw, osw)), "nprop": cnt["nprop"].tolist(), "naccept": cnt["naccept"].tolist(),
                           "nprop_oracle": ocnt["nprop"].tolist(), "naccept_oracle": ocnt["naccept"].tolist(),
                           "logl_max_rel_err": helpers.logl_err(cfg, st["logl"], ost["logl"], ost["sig"]),
                           "first_divergence_iteration": int(np.argmax((fl != ofl).any(axis=1))) if (fl != ofl).any() else None}
# the same identity with a buried station (sample configuration, receiver 0.5 km below the sea floor)
cfg = workloads.make_config("sample"); cfg.bdep = 0.5
cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
n_iter = 200
pt = ParallelTempering(cfg, nproc); pt.set_logging(n_iter); pt.run(n_iter)
fl, ty, sw = pt.log(n_iter); st = pt.state(); pt.close()
orc = oracle_c.OraclePT(cfg, nproc); ofl, oty, osw = orc.run(n_iter); ost = orc.state()
out["pt_fixed_seed_C1_buried"] = {"iterations": n_iter, "chains": nproc * cfg.nchains, "bdep": cfg.bdep,
                                  "accept_flags_identical": bool(np.array_equal(fl, ofl)),
                                  "proposal_types_identical": bool(np.array_equal(ty, oty)), "swaps_identical": bool(np.array_equal(sw, osw)),
                                  "logl_max_rel_err": helpers.logl_err(cfg, st["logl"], ost["logl"], ost["sig"])}
print(json.dumps(out, indent=1))
