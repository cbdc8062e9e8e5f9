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
    assert pt.exchange_mode == "none"        # a single process exchanges nothing
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


@pytest.mark.parametrize("variant", ["sigma_solved_vp_solved", "laplace_prior_joint_PS", "single_chain_per_rank", "transform_length_250",
                                     "transform_length_375_odd"])
def test_proposal_variants(variant):
    if variant == "sigma_solved_vp_solved":           # 6 proposal types, sea layer, common rays
        cfg = helpers.small_config(sdep=2.0, vp_mode=1, rayps=[0.06, 0.06], a_gus=[2.5, 4.0], nfft=128, nsmp=64,
                                   sig_min=[0.005, 0.01], sig_max=[0.05, 0.01], nchains=4, ncool=2, t_high=10.0, iseed=777)
        nproc, n_iter = 6, 120
    elif variant == "laplace_prior_joint_PS":          # Laplace sampler (data-dependent draw count), P + S traces
        cfg = helpers.small_config(prior_mode=1, dvs_prior=0.3, dvp_prior=0.1, vp_mode=1, ipha=[1, -1], rayps=[0.06, 0.10],
                                   nfft=128, nsmp=64, nchains=3, ncool=1, t_high=5.0, iseed=4242)
        nproc, n_iter = 5, 120
    elif variant.startswith("transform_length"):       # N_FFT that is not a power of two: the Bluestein kernels inside the PT loop
        n = 250 if variant.endswith("250") else 375
        cfg = helpers.small_config(sdep=1.0, nfft=n, nsmp=101, nchains=4, ncool=1, t_high=8.0, iseed=31337)
        nproc, n_iter = 5, 80
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


def test_posterior_bookkeeping_matches_oracle():
    """nk, nz, nsig, namp, nvpz, nvsz, nvpvsz (integers: exact) and the profile means (1e-12) of the
    non-tempered chains, recorded every ncorr iterations after nburn (src/pt_mcmc.f90:204-286)."""
    cfg = helpers.small_config(sdep=2.0, nfft=128, nsmp=64, nchains=4, ncool=2, t_high=8.0, nburn=20, niter=100, ncorr=5,
                               sig_min=[0.005, 0.01], sig_max=[0.05, 0.01], nbin_z=40, nbin_vs=25, nbin_vp=20,
                               nbin_vpvs=30, nbin_sig=10, nbin_amp=32, vp_mode=1, iseed=31337)
    cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
    nproc, n_iter = 5, 120
    pt = ParallelTempering(cfg, nproc)
    pt.set_logging(n_iter)
    pt.run(n_iter)
    hg = pt.hist()
    vp_g, vs_g = pt.models()
    flags_g = pt.log(n_iter)[0]
    pt.close()
    orc = oracle_c.OraclePT(cfg, nproc)
    flags_o, _, _ = orc.run(n_iter, record=True)
    ho = orc.hist()
    nmod_o = orc.counters(n_iter)["nmod"]
    assert np.array_equal(flags_g, flags_o)
    assert hg["nmod"] == nmod_o == 20 * nproc * cfg.ncool
    for key in ("nk", "nz", "nsig", "namp", "nvpz", "nvsz", "nvpvsz"):
        assert np.array_equal(hg[key], ho[key]), key
        assert hg[key].sum() > 0, key
    for key in ("vp_mean", "vs_mean", "vpvs_mean"):
        assert np.allclose(hg[key], ho[key], rtol=1e-12, atol=0), key
    # ocean bins carry the reference's assignment quirk (src/pt_mcmc.f90:260-263)
    assert hg["vs_mean"][0] == cfg.vs_min and hg["vpvs_mean"][0] == cfg.vpvs_min
    assert vp_g.shape == (hg["nmod"], cfg.nbin_z) and (vs_g[:, 0] == cfg.vs_min).all()


def test_run_driver_writes_reference_output_set(tmp_path):
    """params.in + SAC + velmod in, the reference's 12 output files out (src/rf_inv.f90:28-108, src/mcmc_out.f90)."""
    import os
    from rf_inv_b200 import io as rio, run as rrun
    cfg = helpers.attach_obs_and_rinv(workloads.make_config("sample"), noise=0.01)
    d = tmp_path
    (d / "data").mkdir(); (d / "model").mkdir()
    for t in range(2):
        rio.write_sac(str(d / "data" / f"s{t + 1}.trc"), cfg.obs[t], cfg.delta, 0.0)
    with open(d / "model" / "ref.velmod", "w") as f:
        for i in range(61):
            f.write(f"{0.5 * i} 5.00 2.89\n")
    vals = ["'./rslt'", 40, 120, 10, 5, 1, 15.0, 12345678, 2, 0.06, 0.08, 4.0, 4.0, 1, 1, 256, "'data/s1.trc'", "'data/s2.trc'",
            "0.0 5.0", 0, 2.0, '"model/ref.velmod"', 0, "1 10", "0.0 20.0", 0.05, 2, 2.0, 0.2, "0.01 0.01", "0.01 0.01", 0.02, 0.02,
            0.02, 0.002, 100, 50, 50, 100, 50, 100, "-0.8 0.8", "0.1 8.6", "0.001 5.0", "0.0 5.0"]
    (d / "params.in").write_text("# test problem in the reference's params.in format\n" + "\n".join(str(v) for v in vals) + "\n")
    hist, cnt = rrun.run(str(d / "params.in"), nproc=4, verbose=False)
    out = d / "rslt"
    expected = {"params.in.copy", "all_models", "likelihood", "num_interface.ppd", "syn_trace.ppd", "interface_depth.ppd",
                "sigma.ppd", "vs_z.ppd", "vp_z.ppd", "vpvs_z.ppd", "vs_z.mean", "vp_z.mean", "vpvs_z.mean"}
    assert expected <= set(os.listdir(out))
    assert os.path.exists(d / "input01") and os.path.exists(d / "input02")
    assert hist["nmod"] == 12 * 4 and cnt["nprop"].sum() == 160 * 4
    lk = np.loadtxt(out / "likelihood")
    assert lk.shape == (160, 2) and np.isfinite(lk).all()
    nk = np.loadtxt(out / "num_interface.ppd")
    assert abs(nk[:, 1].sum() - 1.0) < 1e-12
    syn = np.loadtxt(out / "syn_trace.ppd")
    assert syn.shape == (2 * 101 * 100, 4) and abs(syn[:100, 2].sum() - 1.0) < 1e-3
    assert os.path.getsize(out / "sigma.ppd") == 0                    # sigma fixed for both traces: nothing written


def test_full_length_sample_run_recovers_the_data_generating_structure():
    """The sample configuration at the reference's own run length (sample_syn/params.in: N_BURN 3000 + N_ITER 8000, 20
    ranks x 5 chains = 1.1 M chain steps, a few seconds here) on synthetic data of true.velmod's structure (fast layer over
    a slow layer over a fast half space below a 2 km sea): the posterior-mean Vs profile (vs_z.mean = vs_mean / nmod)
    must show that structure.  Fixed seed and deterministic kernels: the numbers are reproducible."""
    cfg = helpers.attach_obs_and_rinv(workloads.make_config("sample"), noise=0.01)
    pt = ParallelTempering(cfg, 20)
    pt.run(cfg.nburn + cfg.niter)
    h = pt.hist(); cnt = pt.counters(); pt.close()
    assert h["nmod"] == (cfg.niter // cfg.ncorr) * 20 * cfg.ncool
    assert cnt["nprop"].sum() == (cfg.nburn + cfg.niter) * 20 * cfg.ncool
    dz = cfg.z_max / cfg.nbin_z
    mean = lambda z: h["vs_mean"][int(z / dz)] / h["nmod"]
    tm = workloads.true_model(cfg)
    top, mid, bot = 2.89 + tm["dvs"][0, 0], 2.89 + tm["dvs"][0, 1], 2.89 + tm["dvs"][0, -1]      # 4.76, 3.54, 4.68 km/s
    v_top, v_mid, v_bot = mean(2.5), mean(7.0), mean(14.0)
    assert v_top - v_mid > 0.4 and v_bot - v_mid > 0.4                     # high - low - high
    assert abs(v_top - top) < 0.5 and abs(v_mid - mid) < 0.5 and abs(v_bot - bot) < 0.5
