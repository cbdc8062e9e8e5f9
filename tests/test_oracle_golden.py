"""Pins the oracle (both restatements) against the reference's own fixtures and against itself.  CPU only."""
import os

import numpy as np
import pytest

import helpers
import oracle_c
import rfinv_oracle as pyo
from rf_inv_b200 import workloads

G = np.load(os.path.join(helpers.GOLDEN, "sample_syn.npz"))
V = np.load(os.path.join(helpers.GOLDEN, "oracle_vectors.npz"))


def golden_cfg():
    # sample_syn/data/sample_{1,2}.trc were made from true/true.velmod WITHOUT a sea layer
    return helpers.small_config(delta=float(G["delta"]), sig_min=[0.01, 0.01], sig_max=[0.01, 0.01])


def test_golden_traces_c_oracle_bit_exact_at_float32():
    vm = G["true_velmod"]
    rft = oracle_c.calc_rf_layers(golden_cfg(), vm[:, 0], vm[:, 1], vm[:, 2], vm[:, 3])
    for t, key in enumerate(("trc1", "trc2")):
        assert np.array_equal(rft[t, :101].astype(np.float32), G[key])
        assert np.max(np.abs(rft[t, :101] - G[key].astype(np.float64))) < 3e-8   # float32 storage of O(0.5) values


def test_golden_traces_numpy_oracle_bit_exact_at_float32():
    vm = G["true_velmod"]
    pc = helpers.py_config(golden_cfg())
    rft = pyo.calc_rf(pc, pyo.init_filter(pc), 3, vm[:, 0], vm[:, 1], vm[:, 2], vm[:, 3])
    assert np.array_equal(rft[:101, 0].astype(np.float32), G["trc1"])
    assert np.array_equal(rft[:101, 1].astype(np.float32), G["trc2"])


def test_density_of_true_velmod_is_bit_exact():
    # true.velmod prints rho(Vp=5.0) = 2.5347508187769563: only reproduced with float32 Brocher coefficients (F8)
    assert G["true_velmod"][0, 2] == 2.5347508187769563
    assert oracle_c.vp_to_rho(5.0) == 2.5347508187769563
    assert pyo.vp_to_rho(5.0) == 2.5347508187769563


def test_sac_header_and_params_fixture():
    assert float(G["delta"]) == float(np.float32(0.05)) and int(G["npts"]) == 101 and float(G["b"]) == 0.0
    p = list(G["params"])
    assert p[1:8] == ["3000", "8000", "10", "5", "1", "15.0", "12345678"] and p[15] == "256"


def test_mt19937_known_answers():
    # two independent implementations agreed on these (not from the reference: parity unpinned)
    kat = {4357: [0.8173300598282367, 0.9990608994849026, 0.510354372439906],
           12345678: [0.5835216229315847, 0.40791033022105694, 0.8269997001625597],
           12355701: [0.6917868393938988, 0.9083909299224615, 0.5067440385464579]}
    for seed, exp in kat.items():
        assert list(oracle_c.mt_sequence(seed, 3)) == exp
        r = pyo.MT19937(seed)
        assert [r.grnd() for _ in range(3)] == exp
    a = oracle_c.mt_sequence(99, 2000)          # crosses the 624-word reload three times
    r = pyo.MT19937(99)
    assert np.array_equal(a, np.array([r.grnd() for _ in range(2000)]))
    assert pyo.rank_seed(12345678, 1) == 12355701


def test_mt19937_matches_numpy_core():
    seed = 2024
    key = np.zeros(624, dtype=np.uint32)
    key[0] = seed
    for i in range(1, 624):
        key[i] = (69069 * int(key[i - 1])) & 0xFFFFFFFF
    bg = np.random.MT19937()
    st = bg.state
    st["state"]["key"] = key
    st["state"]["pos"] = 624
    bg.state = st
    exp = bg.random_raw(1500) / 4294967296.0
    assert np.array_equal(oracle_c.mt_sequence(seed, 1500), exp)


@pytest.mark.parametrize("kind", ["gauss", "laplace"])
def test_deviates_c_vs_numpy_oracle(kind):
    a = oracle_c.deviates(777, kind, 500)
    r = pyo.MT19937(777)
    f = pyo.gauss if kind == "gauss" else pyo.laplace
    b = np.array([f(r) for _ in range(500)])
    assert np.array_equal(a, b)          # same libm on the same host: bit identical
    assert abs(np.mean(a)) < 0.2 and 0.7 < np.std(a) < (1.7 if kind == "laplace" else 1.2)


CASES = ["land_P", "sea_P", "land_S", "sea_S_deconv", "P_deconv", "common", "vp1_tstart", "buried_sea_P", "buried_land_S",
         "buried_half_space"]
KW = {"land_P": dict(), "sea_P": dict(sdep=2.0), "land_S": dict(ipha=[-1, -1], rayps=[0.10, 0.12]),
      "sea_S_deconv": dict(sdep=1.0, ipha=[-1, -1], deconv_mode=1), "P_deconv": dict(deconv_mode=1),
      "common": dict(rayps=[0.06, 0.06], a_gus=[2.0, 4.0]), "vp1_tstart": dict(vp_mode=1, t_start=-3.0),
      "buried_sea_P": dict(bdep=1.5, sdep=2.0), "buried_land_S": dict(bdep=6.0, ipha=[-1, -1], rayps=[0.10, 0.12]),
      "buried_half_space": dict(bdep=25.0)}


def case_cfg(name):
    cfg = helpers.small_config(**KW[name])
    cfg.obs = V[name + "/obs"]
    cfg.r_inv = helpers.scipy_r_inv(cfg)
    return cfg


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_committed_numpy_vectors(name):
    cfg = case_cfg(name)
    m = {k: V[f"{name}/{k}"] for k in ("k", "z", "dvp", "dvs", "sig")}
    ll, rft, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    assert helpers.rel_err_rft(rft, V[name + "/rft"]) < 1e-12      # restatement vs restatement
    assert helpers.logl_err(cfg, ll, V[name + "/logl"], m["sig"]) < 1e-11


@pytest.mark.parametrize("name", ["land_P", "sea_S_deconv"])
def test_numpy_oracle_reproduces_committed_vectors(name):
    cfg = case_cfg(name)
    pc = helpers.py_config(cfg)
    flt = pyo.init_filter(pc)
    rinv = np.transpose(cfg.r_inv, (2, 1, 0))
    for i in range(2):
        ll, rft = pyo.calc_likelihood(pc, flt, rinv, int(V[name + "/k"][i]), V[name + "/z"][i], V[name + "/dvp"][i],
                                      V[name + "/dvs"][i], V[name + "/sig"][i])
        assert helpers.rel_err_rft(rft.T[None], V[name + "/rft"][i][None]) < 1e-13
        assert helpers.logl_err(cfg, np.array([ll]), V[name + "/logl"][i:i + 1], V[name + "/sig"][i:i + 1]) < 1e-11


@pytest.mark.parametrize("nfft", [96, 250, 375, 1000])
def test_c_oracle_matches_numpy_oracle_for_lengths_that_are_not_powers_of_two(nfft):
    """FFTW takes any length (src/fftw.f90:43-45).  numpy restatement: numpy.fft.irfft(x, n) * n; C restatement: the
    defining sum over the half spectrum -- independent of each other and of the product's Bluestein path."""
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(nfft=nfft, nsmp=min(101, nfft), sdep=1.0, ipha=[1, -1], rayps=[0.06, 0.10]), noise=0.01)
    m = workloads.draw_models(cfg, 4, seed=11, dvs_scale=0.3)
    ll, rft, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    pc = helpers.py_config(cfg)
    flt = pyo.init_filter(pc)
    rinv = np.transpose(cfg.r_inv, (2, 1, 0))
    for i in range(4):
        ll_n, rft_n = pyo.calc_likelihood(pc, flt, rinv, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i], m["sig"][i])
        assert helpers.rel_err_rft(rft[i][None], rft_n.T[None]) < 1e-12
        assert helpers.logl_err(cfg, ll[i:i + 1], np.array([ll_n]), m["sig"][i:i + 1]) < 1e-10


def test_format_model_c_vs_numpy_bit_exact():
    cfg = helpers.small_config(sdep=2.0, vp_mode=1)
    pc = helpers.py_config(cfg)
    m = workloads.draw_models(cfg, 50, seed=5)
    m["z"][:10] *= 0.05      # force some invalid thin layers
    for i in range(50):
        a = pyo.format_model(pc, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i])
        b = oracle_c.format_model(cfg, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i])
        assert a[0] == b[0] and a[5] == b[5]
        for x, y in zip(a[1:5], b[1:5]):
            assert np.array_equal(x, y)


def test_sea_layer_of_vanishing_thickness_reproduces_land():
    # analytic consistency check for the un-fixtured water-layer boundary condition
    land = helpers.attach_obs_and_rinv(helpers.small_config())
    m = workloads.draw_models(land, 8, seed=2, dvs_scale=0.3)
    _, rft_land, _ = oracle_c.eval_batch(land, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    sea = helpers.small_config(sdep=1e-9)
    sea.obs, sea.r_inv = land.obs, land.r_inv
    _, rft_sea, _ = oracle_c.eval_batch(sea, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    assert helpers.rel_err_rft(rft_sea, rft_land) < 1e-8


def test_data_generating_model_has_zero_misfit():
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(sdep=2.0))
    tm = workloads.true_model(cfg)
    ll, rft, valid = oracle_c.eval_batch(cfg, tm["k"], tm["z"], tm["dvp"], tm["dvs"], tm["sig"])
    assert valid[0]
    assert np.max(np.abs(rft[0, :, :cfg.nsmp] - cfg.obs)) < 1e-7          # float32 storage of obs
    expected = -sum(cfg.nsmp * np.log(s) for s in cfg.sig_min)
    assert abs(ll[0] - expected) < 1e-3 * abs(expected)


def test_pt_mcmc_c_vs_numpy_oracle_identical_sequences():
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(sdep=2.0, nchains=3, ncool=1, t_high=15.0, nfft=128, nsmp=64,
                                                           sig_min=[0.01, 0.005], sig_max=[0.01, 0.05]), noise=0.01)
    nproc, n_iter = 2, 25
    pc = helpers.py_config(cfg)
    flt = pyo.init_filter(pc)
    rinv = np.transpose(cfg.r_inv, (2, 1, 0))
    ranks, res = pyo.pt_run(pc, flt, rinv, nproc, n_tot_iter=n_iter)
    pt = oracle_c.OraclePT(cfg, nproc, nthreads=2)
    flags, itypes, swaps = pt.run(n_iter)
    assert np.array_equal(flags, np.array(res.accept_bits, dtype=np.int8))
    assert np.array_equal(itypes, np.array(res.itypes, dtype=np.int8))
    assert [tuple(s) for s in swaps] == res.swaps
    cnt = pt.counters(n_iter)
    assert np.array_equal(cnt["nprop"], res.nprop) and np.array_equal(cnt["naccept"], res.naccept)
    st = pt.state()
    logl_py = np.concatenate([r.logl for r in ranks])
    assert np.max(np.abs(st["logl"] - logl_py) / np.abs(logl_py)) < 1e-10
    assert np.array_equal(st["temps"], np.concatenate([r.temps for r in ranks]))
    assert np.array_equal(st["k"], np.concatenate([r.k for r in ranks]))


# ---- buried station (BOREHOLE_DEP): commented out in the reference (src/forward.f90:289-338, 493-516), no fixture ----
BURIED = {"land": dict(bdep=1.0), "sea": dict(bdep=1.0, sdep=2.0), "S": dict(bdep=7.3, ipha=[-1, -1], rayps=[0.10, 0.12]),
          "half_space": dict(bdep=25.0), "sea_deconv": dict(bdep=3.0, sdep=1.0, deconv_mode=1),
          "common": dict(bdep=4.0, rayps=[0.06, 0.06], a_gus=[2.0, 4.0])}


def _py_rft(cfg, m, i, **over):
    pc = helpers.py_config(cfg)
    for key, val in over.items():
        setattr(pc, key, val)
    nlay, a, b, r, h, _ = pyo.format_model(pc, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i])
    return pyo.calc_rf(pc, pyo.init_filter(pc), nlay, a, b, r, h).T


@pytest.mark.parametrize("name", sorted(BURIED))
def test_buried_station_c_vs_numpy_oracle(name):
    cfg = helpers.small_config(**BURIED[name])
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = np.zeros((cfg.ntrc, cfg.nsmp, cfg.nsmp))
    m = workloads.draw_models(cfg, 6, seed=3, dvs_scale=0.3)
    _, rft_c, _ = oracle_c.eval_batch(cfg, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
    for i in range(6):
        assert helpers.rel_err_rft(_py_rft(cfg, m, i)[None], rft_c[i][None]) < 1e-12


def test_buried_station_at_vanishing_depth_reproduces_the_surface_station():
    for kw in (dict(), dict(sdep=2.0), dict(ipha=[-1, -1], rayps=[0.10, 0.12])):
        top = helpers.small_config(**kw)
        top.obs = np.zeros((top.ntrc, top.nsmp)); top.r_inv = np.zeros((top.ntrc, top.nsmp, top.nsmp))
        m = workloads.draw_models(top, 8, seed=2, dvs_scale=0.3)
        _, rft_top, _ = oracle_c.eval_batch(top, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
        bur = helpers.small_config(bdep=1e-9, **kw)
        bur.obs, bur.r_inv = top.obs, top.r_inv
        _, rft_bur, _ = oracle_c.eval_batch(bur, m["k"], m["z"], m["dvp"], m["dvs"], m["sig"])
        assert helpers.rel_err_rft(rft_bur, rft_top) < 1e-8


def test_buried_station_restatement_equals_the_commented_block_where_that_block_is_consistent():
    """The numpy oracle can transcribe the commented block literally (bdep_literal): no exit after the station's layer,
    bdep instead of the remaining distance in the half space, v(i0) for every layer below the station.  With the station
    inside the LAST solid layer none of the three matters, and literal == restated to the last bit; with a layer below
    it they differ visibly (which is why the literal form is not what the CUDA path implements)."""
    cfg = helpers.small_config(k_max=6)
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = np.zeros((cfg.ntrc, cfg.nsmp, cfg.nsmp))
    km = cfg.k_max
    m = dict(k=np.array([3, 3], dtype=np.int32), z=np.zeros((2, km - 1)), dvp=np.zeros((2, km)), dvs=np.zeros((2, km)),
             sig=np.full((2, 2), 0.01))
    m["z"][:, :3] = [4.0, 9.0, 15.0]
    m["dvs"][:, :3] = [-0.6, 0.2, 0.5]; m["dvs"][:, km - 1] = 1.2
    cfg.bdep = 12.0                                  # inside the third (last) solid layer
    a = _py_rft(cfg, m, 0); b = _py_rft(cfg, m, 0, bdep_literal=True)
    assert np.array_equal(a, b)
    cfg.bdep = 6.0                                   # second layer: the literal loop runs on through the third
    a = _py_rft(cfg, m, 0); b = _py_rft(cfg, m, 0, bdep_literal=True)
    assert helpers.rel_err_rft(b[None], a[None]) > 1e-3


def forward_test_config(bdep):
    """src/forward_test.f90:36-58: one S receiver function, a = 8, delta = 0.05, t_start = -3, nfft = 1024, borehole 3 km."""
    vp_ref, vs_ref = workloads.reference_velmod(60.0, vp=5.0, vs=2.5)
    from rf_inv_b200.config import RFConfig
    cfg = RFConfig(ntrc=1, nfft=1024, nsmp=512, delta=0.05, t_start=-3.0, rayps=[0.06], a_gus=[8.0], ipha=[-1], deconv_mode=0,
                   sdep=0.0, bdep=bdep, vp_ref=vp_ref, vs_ref=vs_ref, vp_mode=1, k_min=1, k_max=4, z_max=40.0, vp_max=8.6,
                   vs_max=5.0, sig_min=[0.01], sig_max=[0.01])
    cfg.obs = np.zeros((1, 512)); cfg.r_inv = np.zeros((1, 512, 512))
    return cfg


def test_forward_test_program_of_the_reference_with_its_borehole_depth():
    """The reference's own (stale, F5) smoke program: 2 layers alpha = [5, 8], beta = [2.5, 4], rho = [3.0, 3.3],
    h = [20, -10], p = 0.06, S phase, bdep = 3.  No output of it exists (it does not compile against params.f90 any
    more); the two restatements must agree on exactly its inputs, with and without the borehole depth, and the station
    3 km down must see the direct S earlier than the surface station."""
    alpha, beta, rho, h = [5.0, 8.0], [2.5, 4.0], [3.0, 3.3], [20.0, -10.0]
    out = {}
    for bdep in (0.0, 3.0):
        cfg = forward_test_config(bdep)
        rc = oracle_c.calc_rf_layers(cfg, alpha, beta, rho, h)
        pc = helpers.py_config(cfg)
        rp = pyo.calc_rf(pc, pyo.init_filter(pc), 2, np.array(alpha), np.array(beta), np.array(rho), np.array(h)).T
        assert np.isfinite(rc).all() and helpers.rel_err_rft(rp[None], rc[None]) < 1e-12
        out[bdep] = rc
    assert helpers.rel_err_rft(out[3.0][None], out[0.0][None]) > 1e-2      # the borehole depth matters


def test_maxval_normalisation_condition_number_explains_oracle_vs_oracle_differences():
    """src/forward.f90:197-203 divides by maxval(rx) (largest POSITIVE sample).  On nearly transparent models under S
    incidence that is a ripple; the two restatements of the same dense algorithm then differ by eps * cond.  Pins the
    diagnostic (orc_eval_batch_cond) on two such models of the c4 sample and on ordinary ones."""
    cfg = helpers.attach_obs_and_rinv(workloads.make_config("c4"), noise=0.01)
    m = workloads.draw_models(cfg, 8192, seed=2024, dvs_scale=0.5)
    idx = np.array([7734, 4056, 100, 101])
    sub = {k: v[idx] for k, v in m.items()}
    _, rft_c, _, cond = oracle_c.eval_batch(cfg, sub["k"], sub["z"], sub["dvp"], sub["dvs"], sub["sig"], want_cond=True)
    assert cond[0, 2] > 1e6 and cond[1, 2] > 1e6 and np.all(cond[:, :2] == 1.0) and np.all(cond[2:] < 10.0)
    pc = helpers.py_config(cfg)
    flt = pyo.init_filter(pc)
    for j in range(4):
        nlay, a, b, r, h, _ = pyo.format_model(pc, int(sub["k"][j]), sub["z"][j], sub["dvp"][j], sub["dvs"][j])
        rp = pyo.calc_rf(pc, flt, nlay, a, b, r, h).T
        err = np.max(np.abs(rp - rft_c[j]), axis=-1) / np.max(np.abs(rft_c[j]), axis=-1)
        assert np.all(err < 1e-12 * cond[j])
        if j < 2:
            assert err[2] > 1e-12            # far above rounding level: the amplification is real
