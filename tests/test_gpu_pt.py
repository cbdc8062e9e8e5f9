"""Fixed-seed PT-MCMC parity: the device-resident chains must reproduce the oracle's proposal types, null /
accept / reject sequence, swap proposals and counters EXACTLY, and its state to 1e-9 (BASELINE.json north_star)."""
import numpy as np
import pytest

import helpers
import oracle_c
from rf_inv_b200 import workloads
from rf_inv_b200.pt import ParallelTempering

pytestmark = pytest.mark.gpu


def run_both(cfg, nproc, n_iter):
    pt = ParallelTempering(cfg, nproc)
    pt.set_logging(n_iter)
    st0 = pt.state()
    pt.run(n_iter)
    flags, itypes, swaps = pt.log(n_iter)
    st, cnt = pt.state(), pt.counters()
    pt.close()
    orc = oracle_c.OraclePT(cfg, nproc)
    so0 = orc.state()
    o_flags, o_itypes, o_swaps = orc.run(n_iter)
    so, ocnt = orc.state(), orc.counters(n_iter)
    return dict(flags=flags, itypes=itypes, swaps=swaps, st0=st0, st=st, cnt=cnt), \
        dict(flags=o_flags, itypes=o_itypes, swaps=o_swaps, st0=so0, st=so, cnt=ocnt, n_eval=orc.n_eval)


def check(g, o, cfg):
    # initial state: init_model / init_sig / temperatures use the same mt19937 streams
    assert np.array_equal(g["st0"]["k"], o["st0"]["k"])
    for key in ("z", "dvp", "dvs", "sig", "temps"):
        assert np.allclose(g["st0"][key], o["st0"][key], rtol=1e-14, atol=0), key
    assert helpers.logl_err(cfg, g["st0"]["logl"], o["st0"]["logl"], o["st0"]["sig"]) < 1e-9
    # sequences: exact
    assert np.array_equal(g["itypes"], o["itypes"])
    assert np.array_equal(g["flags"], o["flags"])
    assert np.array_equal(g["swaps"], o["swaps"])
    assert np.array_equal(g["cnt"]["nprop"], o["cnt"]["nprop"])
    assert np.array_equal(g["cnt"]["naccept"], o["cnt"]["naccept"])
    assert g["cnt"]["n_eval"] == o["n_eval"]
    # final state
    assert np.array_equal(g["st"]["k"], o["st"]["k"])
    for key in ("z", "dvp", "dvs", "sig", "temps"):
        assert np.allclose(g["st"][key], o["st"][key], rtol=1e-12, atol=1e-14), key
    assert helpers.logl_err(cfg, g["st"]["logl"], o["st"]["logl"], o["st"]["sig"]) < 1e-9
    lh_g, lh_o = g["cnt"]["likelihood_hist"], o["cnt"]["likelihood_hist"]
    assert np.max(np.abs(lh_g - lh_o) / (np.abs(lh_o) + cfg.nsmp * cfg.ntrc)) < 1e-9


def test_sample_syn_config_200_iterations():
    """C1: sample_syn/params.in as shipped (SEA_DEP 2.0, 2 traces, Gaussian prior, sigma fixed), 20 ranks x 5 chains."""
    cfg = workloads.make_config("sample")
    cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
    g, o = run_both(cfg, nproc=20, n_iter=200)
    check(g, o, cfg)
    assert g["cnt"]["nprop"].sum() == 200 * 20            # one cold chain per rank
    assert (g["flags"] == -1).any() and (g["flags"] == 1).any() and (g["flags"] == 0).any()
    assert g["swaps"][:, 2].sum() > 0


@pytest.mark.parametrize("variant", ["sigma_solved_vp_solved", "laplace_prior_joint_PS", "single_chain_per_rank"])
def test_proposal_variants(variant):
    if variant == "sigma_solved_vp_solved":           # 6 proposal types, sea layer, common rays
        cfg = helpers.small_config(sdep=2.0, vp_mode=1, rayps=[0.06, 0.06], a_gus=[2.5, 4.0], nfft=128, nsmp=64,
                                   sig_min=[0.005, 0.01], sig_max=[0.05, 0.01], nchains=4, ncool=2, t_high=10.0, iseed=777)
        nproc, n_iter = 6, 120
    elif variant == "laplace_prior_joint_PS":          # Laplace sampler (data-dependent draw count), P + S traces
        cfg = helpers.small_config(prior_mode=1, dvs_prior=0.3, dvp_prior=0.1, vp_mode=1, ipha=[1, -1], rayps=[0.06, 0.10],
                                   nfft=128, nsmp=64, nchains=3, ncool=1, t_high=5.0, iseed=4242)
        nproc, n_iter = 5, 120
    else:                                               # nchains < 2: the swap step is skipped entirely (pt_mcmc.f90:498)
        cfg = helpers.small_config(nfft=64, nsmp=40, nchains=1, ncool=1, iseed=99)
        nproc, n_iter = 8, 60
    cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
    g, o = run_both(cfg, nproc, n_iter)
    check(g, o, cfg)
    if variant == "single_chain_per_rank":
        assert g["swaps"].sum() == 0


def test_split_driver_matches_single_driver():
    """local_step / swap_table / apply_swap with two handles on one GPU (two 'processes', tables gathered by hand)
    gives the same chain as rfinv_pt_run with one handle: the multi-GPU protocol without a second device."""
    import torch
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(sdep=2.0, nfft=128, nsmp=64, nchains=3, ncool=1, t_high=8.0), noise=0.01)
    nproc, n_iter = 4, 60
    one = ParallelTempering(cfg, nproc)
    one.set_logging(n_iter)
    one.run(n_iter)
    f1, t1, s1 = one.log(n_iter)
    st1 = one.state()
    one.close()
    parts = [ParallelTempering(cfg, nproc, world=2, rank=r) for r in range(2)]
    for p in parts:
        p.set_logging(n_iter)
    lib = parts[0]._lib
    from rf_inv_b200 import capi
    from rf_inv_b200.pt import _tensor_from_ptr
    dev = torch.device("cuda:0")
    tabs = []
    for p in parts:
        ptr, n = p.swap_table()
        tabs.append(_tensor_from_ptr(torch, ptr, n, dev))
    for _ in range(n_iter):
        for p in parts:
            capi.check(lib.rfinv_pt_local_step(p.ev.handle))
        for p in parts:
            p.ev.synchronize()
        gathered = torch.cat(tabs).contiguous()
        for p in parts:
            capi.check(lib.rfinv_pt_apply_swap(p.ev.handle, gathered.data_ptr(), 2))
        for p in parts:
            p.ev.synchronize()
    logs = [p.log(n_iter) for p in parts]
    assert np.array_equal(np.concatenate([l[0] for l in logs], axis=1), f1)
    assert np.array_equal(np.concatenate([l[1] for l in logs], axis=1), t1)
    assert np.array_equal(logs[0][2], s1) and np.array_equal(logs[1][2], s1)
    st2 = [p.state() for p in parts]
    for key in ("k", "z", "dvs", "logl", "temps"):
        assert np.array_equal(np.concatenate([s[key] for s in st2]), st1[key]), key
    for p in parts:
        p.close()
