"""End-to-end check of the science: the sample configuration's full run (N_BURN 3000 + N_ITER 8000, 20 ranks x 5 chains, as
sample_syn/params.in) on synthetic data of the data-generating model; posterior-mean Vs profile vs the true one."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers
from rf_inv_b200 import workloads
from rf_inv_b200.pt import ParallelTempering
cfg = helpers.attach_obs_and_rinv(workloads.make_config("sample"), noise=0.01)
t0 = time.time()
pt = ParallelTempering(cfg, 20)
pt.run(cfg.nburn + cfg.niter)
h = pt.hist(); cnt = pt.counters(); pt.close()
print("run time %.1f s, sampled models %d, naccept/nprop %s / %s" % (time.time() - t0, h["nmod"], cnt["naccept"].tolist(), cnt["nprop"].tolist()))
dz = (cfg.z_max - 0.0) / cfg.nbin_z
tm = workloads.true_model(cfg)
z1, z2 = tm["z"][0, 0], tm["z"][0, 1]
vs_true = lambda z: 0.0 if z < cfg.sdep else (2.89 + tm["dvs"][0, 0] if z < z1 else (2.89 + tm["dvs"][0, 1] if z < z2 else 2.89 + tm["dvs"][0, -1]))
print("interfaces at", z1, z2, " k histogram:", (h["nk"] / h["nk"].sum()).round(3).tolist())
for zq in (2.5, 3.5, 4.5, 6.0, 7.0, 8.5, 9.0, 11.0, 14.0, 18.0):
    ib = int(zq / dz)
    print("  z %.1f km  vs_mean %.3f  true %.3f" % (zq, h["vs_mean"][ib], vs_true(zq)))
