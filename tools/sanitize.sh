#!/bin/bash
# compute-sanitizer over one small evaluation + a few PT iterations (memcheck, racecheck, synccheck); logs in gpurun_out/
cat > /tmp/san_case.py <<'PY'
import sys, os, numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import helpers
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
from rf_inv_b200.pt import ParallelTempering
for kw in (dict(sdep=2.0, ntrc=2), dict(nfft=1024, nsmp=300, k_max=12, ntrc=2, rayps=[0.05, 0.07], a_gus=[2.0, 4.0], sig_min=[0.01, 0.01], sig_max=[0.02, 0.01]),
           dict(rayps=[0.06, 0.06], a_gus=[2.0, 4.0]), dict(deconv_mode=1), dict(bdep=1.0, sdep=2.0), dict(bdep=25.0),
           dict(bdep=6.0, rayps=[0.06, 0.06], nsmp=100),
           dict(nfft=250, nsmp=101, sdep=1.0), dict(nfft=375, nsmp=150, ipha=[1, -1], rayps=[0.06, 0.11]),     # Bluestein path, even / odd
           dict(nfft=600, nsmp=200, deconv_mode=1), dict(nfft=1000, nsmp=300, bdep=2.0, k_max=12)):
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(**kw), noise=0.01)
    m = workloads.draw_models(cfg, 40, seed=3, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        ll, rft, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    assert np.isfinite(ll).all()
cfg = helpers.attach_obs_and_rinv(helpers.small_config(nchains=4, ncool=1, sig_min=[0.005, 0.01], sig_max=[0.05, 0.01], nburn=2, niter=6, ncorr=2,
                                                       nbin_z=20, nbin_vs=20, nbin_vp=20, nbin_vpvs=20, nbin_sig=10, nbin_amp=20, amp_min=-1.0, amp_max=1.0), noise=0.01)
pt = ParallelTempering(cfg, 6); pt.run(8); print("pt ok", pt.counters()["nprop"].sum()); pt.close()
cfg.nchains = 16; cfg.bdep = 0.7
pt = ParallelTempering(cfg, 70); pt.run(3); print("pt 1120 chains ok", pt.counters()["nprop"].sum()); pt.close()   # two rounds of pt_compact_kernel
m = workloads.draw_models(cfg, 8192 + 5, seed=4, dvs_scale=0.3)
with Evaluator(cfg) as ev:      # two-piece upload on the second stream
    ll, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
assert np.isfinite(ll).all()
print("case ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|case ok' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
