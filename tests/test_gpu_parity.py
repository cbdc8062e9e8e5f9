"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the same inputs.

Tolerances (BASELINE.json north_star): synthetic RFs and log-likelihoods within 1e-9 relative in fp64.
  * rft:  max |gpu - oracle| <= 1e-9 * max|rft| per (model, trace)
  * logL: |gpu - oracle| <= 1e-9 * (|logL| + nsmp * sum_t |log sig_t|)   (the terms logL is summed from)
format_model outputs (layer stack, validity flag) are bit-exact.
"""
import os

import numpy as np
import pytest

import helpers
import oracle_c
from rf_inv_b200 import capi, workloads
from rf_inv_b200.evaluator import Evaluator

pytestmark = pytest.mark.gpu

RTOL = 1e-9
V = np.load(os.path.join(helpers.GOLDEN, "oracle_vectors.npz"))
G = np.load(os.path.join(helpers.GOLDEN, "sample_syn.npz"))

VARIANTS = {
    "land_P": dict(), "sea_P": dict(sdep=2.0), "land_S": dict(ipha=[-1, -1], rayps=[0.10, 0.12]),
    "sea_S_deconv": dict(sdep=1.0, ipha=[-1, -1], deconv_mode=1), "P_deconv": dict(deconv_mode=1),
    "common": dict(rayps=[0.06, 0.06], a_gus=[2.0, 4.0]), "vp1_tstart": dict(vp_mode=1, t_start=-3.0),
    "three_traces_mixed": dict(ntrc=3, rayps=[0.05, 0.06, 0.11], a_gus=[2.0, 4.0, 8.0], ipha=[1, 1, -1],
                               sig_min=[0.01, 0.02, 0.03], sig_max=[0.01, 0.02, 0.03], sdep=1.5),
    "n64": dict(nfft=64, nsmp=40), "n128": dict(nfft=128, nsmp=64), "n512": dict(nfft=512, nsmp=200),
    "n1024_k20": dict(nfft=1024, nsmp=512, k_max=20, z_max=40.0),
    "n2048_k30": dict(nfft=2048, nsmp=1000, k_max=30, z_max=40.0),
    "n4096": dict(nfft=4096, nsmp=300, k_max=12),
    "nsmp_eq_nfft": dict(nfft=128, nsmp=128),
    # transform lengths that are not powers of two (FFTW takes any, src/fftw.f90:43-45): Bluestein path
    "n2000_k30": dict(nfft=2000, nsmp=1000, k_max=30, z_max=40.0), "n1000_sea": dict(nfft=1000, nsmp=400, sdep=2.0, k_max=16),
    "n600_S_deconv": dict(nfft=600, nsmp=250, sdep=1.0, ipha=[-1, -1], deconv_mode=1),
    "n250_common": dict(nfft=250, nsmp=101, rayps=[0.06, 0.06], a_gus=[2.0, 4.0]),
    "n100_tstart": dict(nfft=100, nsmp=100, t_start=-1.0), "n375_odd": dict(nfft=375, nsmp=150, ipha=[1, -1], rayps=[0.06, 0.11]),
    "n1501_odd_deconv": dict(nfft=1501, nsmp=500, deconv_mode=1), "n1536_mixed_widths": dict(nfft=1536, nsmp=512, ntrc=3,
        rayps=[0.05, 0.06, 0.07], a_gus=[2.0, 4.0, 8.0], ipha=[1, 1, 1], sig_min=[0.01] * 3, sig_max=[0.01] * 3),
    "buried_n360": dict(bdep=2.5, sdep=1.0, nfft=360, nsmp=150),
    # buried station (BOREHOLE_DEP; commented out in the reference, src/forward.f90:289-338, 493-516)
    "buried_land": dict(bdep=1.0), "buried_sea": dict(bdep=1.0, sdep=2.0),
    "buried_S": dict(bdep=7.3, ipha=[-1, -1], rayps=[0.10, 0.12]), "buried_half_space": dict(bdep=25.0),
    "buried_sea_deconv": dict(bdep=3.0, sdep=1.0, deconv_mode=1), "buried_common": dict(bdep=4.0, rayps=[0.06, 0.06], a_gus=[2.0, 4.0]),
    "buried_n1024_k20": dict(bdep=2.5, sdep=1.0, nfft=1024, nsmp=512, k_max=20, z_max=40.0),
    "buried_n2048_k30": dict(bdep=11.0, nfft=2048, nsmp=1000, k_max=30, z_max=40.0),
}


def gpu_vs_oracle(cfg, models):
    ll_o, rft_o, val_o = oracle_c.eval_batch(cfg, models["k"], models["z"], models["dvp"], models["dvs"], models["sig"])
    with Evaluator(cfg) as ev:
        ll_g, rft_g, val_g = ev.calc_likelihood(models["k"], models["z"], models["dvp"], models["dvs"], models["sig"],
                                                want_rft=True, want_valid=True)
        assert ev.last_launch_count >= 3
    return (ll_g, rft_g, val_g), (ll_o, rft_o, val_o)


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_eval_batch_matches_oracle(name):
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(**VARIANTS[name]), noise=0.01)
    m = workloads.draw_models(cfg, 96, seed=3, dvs_scale=0.3)
    (ll_g, rft_g, val_g), (ll_o, rft_o, val_o) = gpu_vs_oracle(cfg, m)
    assert np.array_equal(val_g, val_o)
    assert helpers.rel_err_rft(rft_g, rft_o) < RTOL
    assert helpers.logl_err(cfg, ll_g, ll_o, m["sig"]) < RTOL


@pytest.mark.parametrize("nfft", [96, 250, 375, 1000, 2000])
def test_lengths_that_are_not_powers_of_two_against_numpy_irfft(nfft):
    """The complete RF (prop_rft(nfft, ntrc)) of the Bluestein path against the numpy restatement, whose c2r is
    numpy.fft.irfft(x, n) * n for any n (the C restatement sums the defining series instead): three independent
    evaluations of the same transform."""
    import rfinv_oracle as pyo
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(nfft=nfft, nsmp=min(101, nfft), sdep=1.0, ipha=[1, -1], rayps=[0.06, 0.10]), noise=0.01)
    m = workloads.draw_models(cfg, 6, seed=11, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        ll, rft, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    pc = helpers.py_config(cfg)
    flt = pyo.init_filter(pc)
    rinv = np.transpose(cfg.r_inv, (2, 1, 0))
    for i in range(6):
        ll_n, rft_n = pyo.calc_likelihood(pc, flt, rinv, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i], m["sig"][i])
        assert helpers.rel_err_rft(rft[i][None], rft_n.T[None]) < RTOL
        assert helpers.logl_err(cfg, ll[i:i + 1], np.array([ll_n]), m["sig"][i:i + 1]) < RTOL


def test_golden_traces_from_reference_fixture():
    """The reference's own fixture: true.velmod -> sample_{1,2}.trc (float32); the CUDA path reproduces every sample
    bit-exactly after rounding to float32, like the oracle does."""
    cfg = helpers.small_config(delta=float(G["delta"]))
    cfg.obs = np.stack([G["trc1"], G["trc2"]]).astype(np.float64)
    cfg.r_inv = helpers.scipy_r_inv(cfg)
    tm = workloads.true_model(cfg)
    with Evaluator(cfg) as ev:
        ll, rft, valid = ev.calc_likelihood(tm["k"], tm["z"], tm["dvp"], tm["dvs"], tm["sig"], want_rft=True, want_valid=True)
        nlay, alpha, beta, rho, h, ok = ev.format_model(tm["k"], tm["z"], tm["dvp"], tm["dvs"])
    assert valid[0] and ok[0] and nlay[0] == 3
    vm = G["true_velmod"]
    assert np.allclose(alpha[0, :3], vm[:, 0], rtol=0, atol=0) and np.allclose(beta[0, :3], vm[:, 1], rtol=0, atol=1e-15)
    assert rho[0, 0] == 2.5347508187769563                      # float32 Brocher coefficients
    assert np.array_equal(rft[0, 0, :101].astype(np.float32), G["trc1"])
    assert np.array_equal(rft[0, 1, :101].astype(np.float32), G["trc2"])
    # observed == synthetic up to float32 storage: phi ~ 0, logL ~ -sum nsmp log(sig)
    assert abs(ll[0] - (-2 * 101 * np.log(0.01))) < 1e-3 * abs(2 * 101 * np.log(0.01))


GOLDEN_KW = dict(VARIANTS, buried_sea_P=dict(bdep=1.5, sdep=2.0), buried_land_S=dict(bdep=6.0, ipha=[-1, -1], rayps=[0.10, 0.12]))


@pytest.mark.parametrize("name", ["land_P", "sea_P", "land_S", "sea_S_deconv", "P_deconv", "common", "vp1_tstart", "buried_sea_P",
                                  "buried_land_S", "buried_half_space"])
def test_committed_numpy_oracle_vectors(name):
    cfg = helpers.small_config(**GOLDEN_KW[name])
    cfg.obs = V[name + "/obs"]
    cfg.r_inv = helpers.scipy_r_inv(cfg)
    m = {k: V[f"{name}/{k}"] for k in ("k", "z", "dvp", "dvs", "sig")}
    with Evaluator(cfg) as ev:
        ll, rft, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    assert helpers.rel_err_rft(rft, V[name + "/rft"]) < RTOL
    assert helpers.logl_err(cfg, ll, V[name + "/logl"], m["sig"]) < RTOL


def test_format_model_bit_exact_incl_invalid_models():
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(sdep=2.0, vp_mode=1))
    m = workloads.draw_models(cfg, 200, seed=5)
    m["z"][:60] *= 0.08                      # interfaces above the sea floor / too thin layers
    m["dvs"][60:90] -= 3.5                   # velocities out of bounds
    with Evaluator(cfg) as ev:
        nlay, alpha, beta, rho, h, ok = ev.format_model(m["k"], m["z"], m["dvp"], m["dvs"])
    n_bad = 0
    for i in range(200):
        o = oracle_c.format_model(cfg, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i])
        assert o[0] == nlay[i] and o[5] == ok[i]
        n = o[0]
        assert np.array_equal(o[1], alpha[i, :n]) and np.array_equal(o[2], beta[i, :n])
        assert np.array_equal(o[3], rho[i, :n]) and np.array_equal(o[4], h[i, :n])
        n_bad += not o[5]
    assert 20 < n_bad < 200


def test_edge_cases_k1_kmax_minus_1_single_model_and_empty():
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(k_max=6))
    km = cfg.k_max
    k = np.array([1, km - 1, 1], dtype=np.int32)
    z = np.zeros((3, km - 1)); dvp = np.zeros((3, km)); dvs = np.zeros((3, km))
    z[0, 0] = 5.0
    z[1, :km - 1] = [12.0, 3.0, 9.0, 6.0, 15.0]       # unsorted on purpose
    dvs[1, :km - 1] = [0.3, -0.2, 0.1, 0.4, 0.0]
    z[2, 0] = 19.999                                  # interface just above z_max
    sig = np.full((3, 2), 0.01)
    m = dict(k=k, z=z, dvp=dvp, dvs=dvs, sig=sig)
    (ll_g, rft_g, val_g), (ll_o, rft_o, val_o) = gpu_vs_oracle(cfg, m)
    assert np.array_equal(val_g, val_o) and helpers.rel_err_rft(rft_g, rft_o) < RTOL
    assert helpers.logl_err(cfg, ll_g, ll_o, sig) < RTOL
    with Evaluator(cfg) as ev:
        ll, _, _ = ev.calc_likelihood(k[:0], z[:0], dvp[:0], dvs[:0], sig[:0])       # empty batch
        assert ll.shape == (0,)
        ll1, _, _ = ev.calc_likelihood(k[:1], z[:1], dvp[:1], dvs[:1], sig[:1])      # ragged: batch of one
        assert abs(ll1[0] - ll_g[0]) == 0.0
        with pytest.raises(capi.RfinvError):                                         # k outside [1, k_max-1]
            ev.calc_likelihood(np.array([km], dtype=np.int32), z[:1], dvp[:1], dvs[:1], sig[:1])


def test_post_critical_ray_gives_nan_like_the_reference():
    # 1/alpha^2 < p^2 -> sqrt of a negative number -> NaN logL (SURVEY.md 7 "post-critical rays"): a value, not an error
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(rayps=[0.06, 0.21]))
    m = workloads.draw_models(cfg, 4, seed=1, dvs_scale=0.1)
    (ll_g, _, _), (ll_o, _, _) = gpu_vs_oracle(cfg, m)
    assert np.isnan(ll_o).all() and np.isnan(ll_g).all()


def test_internal_r_inv_matches_lapack_route():
    """r_inv = NULL: the library builds R^-1 itself (Householder + QL eigen-solver) instead of the caller's LAPACK dgesvd.
    The two matrices are two backward-stable evaluations of the same truncated pseudo-inverse and agree to < 1e-9 of the
    largest entry; logL inherits exactly that input difference, |d phi| <= ||dR^-1||_2 ||m||^2 -- asserted below -- which
    for entries of R^-1 up to 1/1e-3 amounts to up to ~1e-8 of the logL scale.  (A drop-in host passes the LAPACK matrix
    and gets 1e-9; this route exists for hosts without LAPACK.)"""
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(a_gus=[4.0, 2.5]), noise=0.01)
    ref = cfg.r_inv.copy()
    m = workloads.draw_models(cfg, 16, seed=8, dvs_scale=0.3)
    ll_o, rft_o, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    cfg.r_inv = None
    with Evaluator(cfg) as ev:
        own = ev.r_inv()
        ll_g, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    assert np.max(np.abs(own - ref)) / np.max(np.abs(ref)) < 1e-9
    # the logL difference is the difference of the two input matrices, not of the arithmetic: bound it by that
    mis = rft_o[:, :, :cfg.nsmp] - cfg.obs[None, :, :]
    bound = sum(np.linalg.norm(own[t] - ref[t], 2) * np.sum(mis[:, t, :] ** 2, axis=1) / (2.0 * m["sig"][:, t] ** 2) for t in range(cfg.ntrc))
    assert (np.abs(ll_g - ll_o) <= bound + 1e-9 * np.abs(ll_o)).all()
    assert helpers.logl_err(cfg, ll_g, ll_o, m["sig"]) < 1e-8


def test_device_resident_entry_and_determinism():
    import torch
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(nfft=512, nsmp=200), noise=0.01)
    m = workloads.draw_models(cfg, 300, seed=12, dvs_scale=0.3)
    soa = workloads.to_soa(m)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(v).to(dev) for k, v in soa.items()}
    logl = torch.empty(300, dtype=torch.float64, device=dev)
    rft = torch.empty((cfg.ntrc, 300, cfg.nsmp), dtype=torch.float64, device=dev)
    valid = torch.empty(300, dtype=torch.uint8, device=dev)
    with Evaluator(cfg) as ev:
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        runs = []
        for _ in range(2):
            ev.calc_likelihood_device(300, d["k"].data_ptr(), d["z"].data_ptr(), d["dvp"].data_ptr(), d["dvs"].data_ptr(),
                                      d["sig"].data_ptr(), logl.data_ptr(), rft.data_ptr(), valid.data_ptr())
            torch.cuda.synchronize()
            runs.append((logl.cpu().numpy().copy(), rft.cpu().numpy().copy()))
        ll_h, rft_h, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])   # no atomics: bit-stable
    assert np.array_equal(runs[0][0], ll_h)
    assert np.array_equal(np.transpose(runs[0][1], (1, 0, 2)), rft_h[:, :, :cfg.nsmp])
    assert valid.cpu().numpy().all()


@pytest.mark.parametrize("workload,n_models", [("c2", 4096), ("target", 2048), ("c3_buried", 1024), ("c5", 512)])
def test_full_size_properties(workload, n_models):
    """BASELINE.json sizes: the oracle checks a seeded subsample; size-independent properties cover the rest:
    permutation equivariance, sigma scaling of logL (phi is sigma independent), time-shift property of t_start."""
    cfg = workloads.make_config(workload)
    cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
    m = workloads.draw_models(cfg, n_models, seed=21, dvs_scale=0.5)
    with Evaluator(cfg) as ev:
        ll, _, valid = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_valid=True)
        perm = np.random.default_rng(0).permutation(n_models)
        ll_p, _, _ = ev.calc_likelihood(m["k"][perm], m["z"][perm], m["dvp"][perm], m["dvs"][perm], m["sig"][perm])
        sig2 = m["sig"] * 2.0
        ll_2, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], sig2)
    assert valid.all() and np.isfinite(ll).all()
    assert np.array_equal(ll[perm], ll_p)
    # logL(2 sig) = -phi/(8 sig^2) - S sum log(2 sig)  ==>  recover phi-term from both and compare
    S = cfg.nsmp
    q1 = ll + S * np.sum(np.log(m["sig"]), axis=1)          # = -0.5 sum phi/sig^2
    q2 = ll_2 + S * np.sum(np.log(sig2), axis=1)            # = q1 / 4
    assert np.max(np.abs(q2 * 4.0 - q1) / (np.abs(q1) + S)) < 1e-9
    sub = np.arange(0, n_models, max(1, n_models // 48))
    ll_o, _, _ = oracle_c.eval_batch(cfg, m["k"][sub], m["z"][sub], m["dvp"][sub], m["dvs"][sub], m["sig"][sub], want_rft=False)
    assert helpers.logl_err(cfg, ll[sub], ll_o, m["sig"][sub]) < RTOL


def test_band_limit_changes_nothing_above_rounding(tmp_path):
    """forward_kernel skips frequency-bin groups whose Gaussian-filter weight is below 2^-52 of the filter's mass
    (band_limits(), capi.cu).  With RFINV_FULL_BAND=1 every bin is propagated: both runs must agree to rounding
    level (here: 1e-13 of the trace maximum, four orders of magnitude inside the 1e-9 bar), on the target shape
    (the band limit drops one bin group of four) and on mixed Gaussian widths (2, 3 and 4 groups kept per trace)."""
    import subprocess
    import sys
    script = r'''
import sys, numpy as np
sys.path[:0] = [%r, %r, %r]
import helpers
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
out = {}
for name in ("target", "c3"):
    cfg = workloads.make_config(name)
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = np.zeros((cfg.ntrc, cfg.nsmp, cfg.nsmp))
    m = workloads.draw_models(cfg, 64, seed=5, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        _, rft, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    out[name] = rft
np.savez(sys.argv[1], **out)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
       os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    res = {}
    for tag, val in (("band", "0"), ("full", "1")):
        env = dict(os.environ, RFINV_FULL_BAND=val)
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env)
        res[tag] = np.load(path)
    for name in ("target", "c3"):
        a, b = res["band"][name], res["full"][name]
        assert np.isfinite(b).all()
        err = helpers.rel_err_rft(a, b)
        assert err < 1e-13, (name, err)
        assert not np.array_equal(a, b) or name != "target"   # the switch really changes the computation


def test_factored_quadratic_form_matches_dense_form(tmp_path):
    """quadform_kernel evaluates phi = |W^T m|^2 with R^-1 = W W^T where the factor pays (rank ~0.4 S at a = 4), the
    dense symmetric form m^T R^-1 m otherwise (RFINV_QF_DENSE=1 forces it).  Both against the LAPACK-route R^-1 the
    reference would pass; they must agree far inside the 1e-9 bar (the factor drops eigenvalues < 1e-12 lambda_max)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path[:0] = [%r, %r, %r]
import helpers
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
cfg = helpers.attach_obs_and_rinv(workloads.make_config("target"), noise=0.01)
cfg.r_inv = workloads.lapack_r_inv(cfg)
m = workloads.draw_models(cfg, 200, seed=8, dvs_scale=0.3)
with Evaluator(cfg) as ev:
    ll, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
np.savez(sys.argv[1], ll=ll, sig=m["sig"])
''' % (root, os.path.join(root, "tests"), os.path.join(root, "oracle"))
    res = {}
    for tag, val in (("factor", "0"), ("dense", "1")):
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", script, path], check=True, env=dict(os.environ, RFINV_QF_DENSE=val))
        res[tag] = np.load(path)
    cfg = workloads.make_config("target")
    err = helpers.logl_err(cfg, res["factor"]["ll"], res["dense"]["ll"], res["dense"]["sig"])
    assert err < 1e-11, err
    assert not np.array_equal(res["factor"]["ll"], res["dense"]["ll"])   # two different summations


def test_buried_station_at_vanishing_depth_reproduces_the_surface_station_on_the_gpu():
    """The buried-station path (split layer, station pass, combination of the two propagated vectors) against the plain
    path of the same kernels: a station 1e-9 km deep must give the surface result to ~1e-9."""
    for kw in (dict(), dict(sdep=2.0)):
        top = helpers.attach_obs_and_rinv(helpers.small_config(**kw), noise=0.01)
        m = workloads.draw_models(top, 64, seed=2, dvs_scale=0.3)
        bur = helpers.small_config(bdep=1e-9, **kw)
        bur.obs, bur.r_inv = top.obs, top.r_inv
        with Evaluator(top) as ev:
            _, rft_top, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
        with Evaluator(bur) as ev:
            _, rft_bur, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
        assert helpers.rel_err_rft(rft_bur, rft_top) < 1e-8


def test_split_quadratic_form_matches_the_plain_factor_form(tmp_path):
    """R^-1 of the reference commutes with the exchange matrix, so quadform_kernel contracts sums / differences of mirrored
    samples against two half-length factors; RFINV_QF_NOSPLIT=1 keeps the factor over the whole
    window.  Even and odd window lengths, LAPACK-route R^-1: the two must agree far inside the 1e-9 bar."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path[:0] = [%r, %r, %r]
import helpers
from rf_inv_b200 import workloads
from rf_inv_b200.evaluator import Evaluator
out = {}
for name, kw in (("even", dict()), ("odd", dict(nsmp=501))):
    cfg = workloads.make_config("target")
    for k, v in kw.items(): setattr(cfg, k, v)
    cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
    cfg.r_inv = workloads.lapack_r_inv(cfg)
    m = workloads.draw_models(cfg, 200, seed=8, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        ll, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
        out[name + "_form"] = np.array(ev.quadform_form())
    out[name] = ll; out[name + "_sig"] = m["sig"]
np.savez(sys.argv[1], **out)
''' % (root, os.path.join(root, "tests"), os.path.join(root, "oracle"))
    res = {}
    for tag, val in (("split", "0"), ("plain", "1")):
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", script, path], check=True, env=dict(os.environ, RFINV_QF_NOSPLIT=val))
        res[tag] = np.load(path)
    for name, S in (("even", 512), ("odd", 501)):
        cfg = workloads.make_config("target"); cfg.nsmp = S
        assert res["split"][name + "_form"][:, 2].all() and not res["plain"][name + "_form"][:, 2].any()
        assert (res["split"][name + "_form"][:, 0] == res["plain"][name + "_form"][:, 0]).all()      # same total rank
        err = helpers.logl_err(cfg, res["split"][name], res["plain"][name], res["plain"][name + "_sig"])
        assert err < 1e-11, (name, err)
        assert not np.array_equal(res["split"][name], res["plain"][name])   # two different summations


def test_large_host_batch_with_overlapped_upload_equals_the_device_resident_entry():
    """rfinv_eval_batch uploads the models as the caller holds them (chain slowest; prep_kernel reads that layout), batches
    of >= 8192 models in four pieces on a second stream, prep_kernel launched per piece behind the event that marks the
    piece's arrival.  Same arithmetic: logL, validity flags and traces must equal the device-resident (chain fastest, no upload)
    evaluation to the last bit; ragged size, invalid models included."""
    import torch
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(sdep=2.0), noise=0.01)
    n = 8192 + 77
    m = workloads.draw_models(cfg, n, seed=17, dvs_scale=0.3)
    m["z"][5000:5040] *= 0.05                       # some invalid models in the tail piece
    soa = workloads.to_soa(m)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(v).to(dev) for k, v in soa.items()}
    logl = torch.empty(n, dtype=torch.float64, device=dev)
    valid = torch.empty(n, dtype=torch.uint8, device=dev)
    with Evaluator(cfg) as ev:
        ll_h, rft_h, val_h = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True, want_valid=True)
        n_launch = ev.last_launch_count
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        ev.calc_likelihood_device(n, d["k"].data_ptr(), d["z"].data_ptr(), d["dvp"].data_ptr(), d["dvs"].data_ptr(),
                                  d["sig"].data_ptr(), logl.data_ptr(), 0, valid.data_ptr())
        torch.cuda.synchronize()
        ll_h2, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])      # again, other stream, warm buffers
    assert n_launch == 7                            # prep_kernel per piece of the upload (4), forward, quadform (+ logL), sig layout
    assert np.array_equal(ll_h, logl.cpu().numpy(), equal_nan=True) and np.array_equal(ll_h, ll_h2, equal_nan=True)
    assert np.array_equal(val_h, valid.cpu().numpy().astype(bool)) and not val_h[5000:5040].all()
    sub = np.arange(0, n, 97)
    _, rft_o, _ = oracle_c.eval_batch(cfg, m["k"][sub], m["z"][sub], m["dvp"][sub], m["dvs"][sub], m["sig"][sub])
    ok = val_h[sub]
    assert helpers.rel_err_rft(rft_h[sub][ok], rft_o[ok]) < RTOL


@pytest.mark.parametrize("kind,expect", [("low_rank", "factor"), ("full_rank", "dense"), ("toeplitz", "split")])
def test_quadratic_form_picks_a_form_that_fits_the_matrix_it_is_given(kind, expect):
    """The caller may pass any R^-1 (src/likelihood.f90 only ever builds the Toeplitz one).  A positive semi-definite matrix
    without the mirror symmetry gets the plain factor form when its rank is low and the dense form otherwise; the split
    form is only taken when R^-1 commutes with the exchange matrix.  All three against the oracle's dense m^T R^-1 m."""
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(nfft=512, nsmp=200), noise=0.01)
    rng = np.random.default_rng(4)
    S = cfg.nsmp
    if kind != "toeplitz":
        r = 40 if kind == "low_rank" else S
        mats = []
        for t in range(cfg.ntrc):
            a = rng.normal(size=(S, r)) / np.sqrt(S)
            mats.append(a @ a.T * 50.0)
        cfg.r_inv = np.stack(mats)
    m = workloads.draw_models(cfg, 128, seed=9, dvs_scale=0.3)
    ll_o, _, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    with Evaluator(cfg) as ev:
        form = ev.quadform_form()
        ll_g, _, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    for rank, rank_s, split in form:
        got = "split" if split else ("factor" if rank > 0 else "dense")
        assert got == expect, form
    if kind == "low_rank":
        assert all(rank == 40 for rank, _, _ in form)
    assert helpers.logl_err(cfg, ll_g, ll_o, m["sig"]) < RTOL


def test_forward_test_geometry_of_the_reference_with_borehole_depth():
    """src/forward_test.f90's model (20 km layer alpha 5 / beta 2.5 over a half space alpha 8 / beta 4, p = 0.06, S phase,
    a = 8, t_start = -3, nfft = 1024, borehole depth 3 km) through the model interface (densities from vp_to_rho: the
    boundary has no density input), surface and buried station, against the oracle."""
    from test_oracle_golden import forward_test_config
    for bdep in (0.0, 3.0):
        cfg = forward_test_config(bdep)
        km = cfg.k_max
        m = dict(k=np.array([1], dtype=np.int32), z=np.zeros((1, km - 1)), dvp=np.zeros((1, km)), dvs=np.zeros((1, km)),
                 sig=np.full((1, 1), 0.01))
        m["z"][0, 0] = 20.0
        m["dvp"][0, km - 1] = 3.0; m["dvs"][0, km - 1] = 1.5
        (ll_g, rft_g, val_g), (ll_o, rft_o, val_o) = gpu_vs_oracle(cfg, m)
        assert val_g[0] and val_o[0]
        assert helpers.rel_err_rft(rft_g, rft_o) < RTOL


def test_parity_on_the_joint_p_and_s_workload_is_bounded_by_the_conditioning_of_the_reference_normalisation():
    """deconv_mode 0 divides every trace by maxval(rx) of the filtered vertical trace (src/forward.f90:197-203): the largest
    POSITIVE sample.  For S incidence on a nearly transparent model the main pulse of that trace is negative and the
    divisor is a ripple 1e-6 .. 1e-7 of it, so ANY two fp64 evaluations of the reference's own formulas agree to
    eps * cond only, cond = max|rx| / maxval(rx) (the numpy and the C oracle differ by 7e-11 on such models).  The bar:
    1e-9 wherever cond <= 1e3 (measured ~1e-14), 1e-12 * cond everywhere -- and such models exist in the sample."""
    cfg = helpers.attach_obs_and_rinv(workloads.make_config("c4"), noise=0.01)
    m = workloads.draw_models(cfg, 8192, seed=2024, dvs_scale=0.5)
    ll_o, rft_o, _, cond = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_cond=True)
    with Evaluator(cfg) as ev:
        ll_g, rft_g, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    err = np.max(np.abs(rft_g - rft_o), axis=-1) / np.max(np.abs(rft_o), axis=-1)
    well = cond <= 1.0e3
    assert (~well).sum() > 0 and cond.max() > 1.0e6                    # the ill-conditioned cases are in the sample
    assert np.max(err[well]) < 1e-12
    assert np.max(err / cond) < 1e-12
    ok = well.all(axis=1)
    assert helpers.logl_err(cfg, ll_g[ok], ll_o[ok], m["sig"][ok]) < 1e-11


def test_async_slots_match_the_synchronous_call():
    """rfinv_eval_batch_begin / _end (two groups of chains in flight, each on its own stream and workspace) return exactly what
    rfinv_eval_batch returns for the same models, whatever the order the slots are begun and ended in."""
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(sdep=2.0, ntrc=2), noise=0.01)
    m = workloads.draw_models(cfg, 600, seed=11, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        ll_ref, _, val_ref = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_valid=True)
        halves = [slice(0, 350), slice(350, 600)]
        parts = [{k: np.ascontiguousarray(v[s]) for k, v in m.items()} for s in halves]
        for order in ((0, 1), (1, 0)):
            out = [np.full(p["k"].shape[0], np.nan) for p in parts]
            val = [np.zeros(p["k"].shape[0], dtype=np.uint8) for p in parts]
            for rep in range(3):          # slots are reusable
                for s in (0, 1):
                    p = parts[s]
                    ev.calc_likelihood_begin(s, p["k"], p["z"], p["dvp"], p["dvs"], p["sig"], out[s], val[s])
                for s in order:
                    ev.calc_likelihood_end(s)
            assert np.array_equal(np.concatenate(out), ll_ref)
            assert np.array_equal(np.concatenate(val).astype(bool), val_ref)
        # a slot in flight cannot be begun again
        p = parts[0]
        o = np.empty(p["k"].shape[0])
        ev.calc_likelihood_begin(0, p["k"], p["z"], p["dvp"], p["dvs"], p["sig"], o)
        with pytest.raises(capi.RfinvError):
            ev.calc_likelihood_begin(0, p["k"], p["z"], p["dvp"], p["dvs"], p["sig"], o)
        ev.calc_likelihood_end(0)
        with pytest.raises(capi.RfinvError):
            ev.calc_likelihood_begin(2, p["k"], p["z"], p["dvp"], p["dvs"], p["sig"], o)


def test_large_host_batch_goes_up_in_pieces_and_matches_small_batches():
    """>= 8192 models: rfinv_eval_batch uploads in pieces and launches prep_kernel per piece behind the piece's event; the
    result equals the evaluation of the same models in small batches (one piece, no overlap) bit for bit."""
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(), noise=0.01)
    m = workloads.draw_models(cfg, 9000, seed=5, dvs_scale=0.3)
    with Evaluator(cfg) as ev:
        ll_big, _, val_big = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_valid=True)
        ll_small = np.concatenate([ev.calc_likelihood(m["k"][s:s + 1500], m["z"][s:s + 1500], m["dvp"][s:s + 1500], m["dvs"][s:s + 1500],
                                                      m["sig"][s:s + 1500])[0] for s in range(0, 9000, 1500)])
    assert np.array_equal(ll_big, ll_small)
    ok = oracle_c.eval_batch(cfg, m["k"][:64], m["z"][:64], m["dvp"][:64], m["dvs"][:64], m["sig"][:64])
    assert helpers.logl_err(cfg, ll_big[:64], ok[0], m["sig"][:64]) < RTOL and np.array_equal(val_big[:64], ok[2])


def test_per_model_fwd_flag_matches_the_oracles_cached_rf_branch():
    """rfinv_eval_batch_flags against calc_likelihood(fwd_flag) of the oracle (src/likelihood.f90:74-82): models with fwd_flag
    are propagated, sigma-only proposals (fwd_flag = .false.) re-use the chain's cached RF -- here through the cached
    quadratic forms the first call returned."""
    import rfinv_oracle as pyo
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(ntrc=2, sig_min=[0.005, 0.005], sig_max=[0.05, 0.05]), noise=0.01)
    n = 48
    cur = workloads.draw_models(cfg, n, seed=9, dvs_scale=0.3)        # the chains' current states
    rng = np.random.default_rng(4)
    phi = np.zeros((n, cfg.ntrc))
    with Evaluator(cfg) as ev:
        ll0, _ = ev.calc_likelihood_flags(np.ones(n, dtype=np.uint8), cur["k"], cur["z"], cur["dvp"], cur["dvs"], cur["sig"], phi)
        ll_ref, rft_ref, _ = ev.calc_likelihood(cur["k"], cur["z"], cur["dvp"], cur["dvs"], cur["sig"], want_rft=True)
        assert np.array_equal(ll0, ll_ref)
        # proposals: every other chain proposes a new sigma only (fwd_flag false), the rest a new model
        flags = (np.arange(n) % 2).astype(np.uint8)
        prop = workloads.draw_models(cfg, n, seed=10, dvs_scale=0.3)
        prop["sig"] = rng.uniform(0.005, 0.05, (n, cfg.ntrc))
        phi_cached = phi.copy()
        ll1, valid = ev.calc_likelihood_flags(flags, prop["k"], prop["z"], prop["dvp"], prop["dvs"], prop["sig"], phi, want_valid=True)
        assert np.array_equal(phi[flags == 0], phi_cached[flags == 0]) and not np.array_equal(phi[flags == 1], phi_cached[flags == 1])
    pc = helpers.py_config(cfg)
    flt = pyo.init_filter(pc)
    r_inv = np.ascontiguousarray(np.transpose(cfg.r_inv, (2, 1, 0)))
    ll_o = np.empty(n)
    for c in range(n):
        cached = np.ascontiguousarray(rft_ref[c].T)                   # the chain's rft(nfft, ntrc) as the reference keeps it
        ll_o[c], _ = pyo.calc_likelihood(pc, flt, r_inv, int(prop["k"][c]), prop["z"][c], prop["dvp"][c], prop["dvs"][c], prop["sig"][c],
                                         fwd_flag=bool(flags[c]), cached_rft=cached)
    assert helpers.logl_err(cfg, ll1, ll_o, prop["sig"]) < RTOL
    assert valid[flags == 0].all()


def _arbiter_variants():
    path = os.path.join(helpers.GOLDEN, "arbiter_vectors.npz")
    return sorted({k.split("/")[0] for k in np.load(path).files})


@pytest.mark.parametrize("name", _arbiter_variants())
def test_cuda_against_the_extended_precision_arbiter(name):
    """The CUDA path against the mpmath arbiter (oracle/arbiter_mp.py; fixture tests/golden/arbiter_vectors.npz), next to the C
    oracle on the same models: the CUDA error is within the north star's 1e-9 wherever the reference's normalisation is well
    conditioned (cond <= 1e3), and nowhere worse than the dense fp64 restatement of the reference plus 1e-12 of the scale --
    on the ill-conditioned S traces of the joint P + S workload (cond up to 1e7) BOTH fp64 evaluations sit eps * cond away
    from the exact value of the reference's formulas."""
    import test_arbiter as ta
    cfg, m, mp_ = ta.load_variant(name)
    ll_o, rft_o, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    with Evaluator(cfg) as ev:
        ll_g, rft_g, _ = ev.calc_likelihood(m["k"], m["z"], m["dvp"], m["dvs"], m["sig"], want_rft=True)
    eg_rft, eg_ll = ta.errors_vs_arbiter(cfg, m, mp_, rft_g, ll_g)
    eo_rft, eo_ll = ta.errors_vs_arbiter(cfg, m, mp_, rft_o, ll_o)
    c = np.maximum(mp_["cond"], 1.0)
    cm = c.max(axis=1)
    assert (eg_rft[c <= 1e3] <= 1e-9).all() and (eg_ll[cm <= 1e3] <= 1e-9).all()
    assert (eg_rft <= eo_rft + 1e-12 * c).all(), float(np.max((eg_rft - eo_rft) / c))
    assert (eg_ll <= eo_ll + 1e-11 * cm).all(), float(np.max((eg_ll - eo_ll) / cm))
    assert (eg_rft <= 1e-12 * c).all() and (eg_ll <= 1e-11 * cm).all()      # and in absolute terms: eps * cond with a small constant
