"""make_syn (src/make_syn.f90) through the C-ABI against the numpy restatement (oracle/rfinv_oracle.py: make_syn).
No fixture of make_syn exists in the reference (parity unpinned): the check is restatement vs CUDA path."""
import os

import numpy as np
import pytest

import helpers
import rfinv_oracle as pyo
from rf_inv_b200 import make_syn, workloads
from rf_inv_b200 import io as rio

pytestmark = pytest.mark.gpu


def _compare(cfg):
    ref = pyo.make_syn(helpers.py_config(cfg))
    got = make_syn.synthesize(cfg)
    assert got["k"] == ref["k"] and got["nlay"] == ref["nlay"]
    kk = ref["k"]
    np.testing.assert_allclose(got["z"][:kk], ref["z"][:kk], rtol=1e-13, atol=0)
    np.testing.assert_allclose(got["dvs"], ref["dvs"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(got["dvp"], ref["dvp"], rtol=1e-12, atol=1e-15)
    for key in ("alpha", "beta", "rho", "h"):
        np.testing.assert_allclose(got[key], ref[key], rtol=1e-12, atol=0)
    np.testing.assert_allclose(got["noise_sigma"], ref["noise_sigma"], rtol=1e-14, atol=0)
    scale = np.max(np.abs(ref["noise"]), axis=1, keepdims=True)
    assert np.max(np.abs(got["noise"] - ref["noise"]) / scale) < 1e-11     # white noise + two FFTs of length nfft
    assert helpers.rel_err_rft(got["rft"], ref["rft"]) < 1e-9
    assert np.max(np.abs(got["noisy"] - ref["noisy"]) / np.max(np.abs(ref["noisy"]), axis=1, keepdims=True)) < 1e-9
    return got, ref


@pytest.mark.parametrize("variant", ["distinct_rays_sigma_range", "common_ray", "sea_laplace_n1024", "n1000_any_length", "n375_odd_length"])
def test_synthesize_matches_restatement(variant):
    kw = {
        "distinct_rays_sigma_range": dict(sig_min=[0.005, 0.01], sig_max=[0.05, 0.01], nchains=3, iseed=4242),
        "common_ray": dict(rayps=[0.06, 0.06], a_gus=[2.0, 4.0], sig_min=[0.01, 0.01], sig_max=[0.03, 0.03], nchains=2),
        "sea_laplace_n1024": dict(nfft=1024, nsmp=400, sdep=1.5, prior_mode=1, k_max=14, z_max=30.0, nchains=2,
                                  sig_min=[0.01, 0.02], sig_max=[0.02, 0.04], iseed=99),
        "n1000_any_length": dict(nfft=1000, nsmp=400, nchains=2, sig_min=[0.01, 0.02], sig_max=[0.02, 0.04], iseed=7),
        "n375_odd_length": dict(nfft=375, nsmp=150, rayps=[0.06, 0.06], nchains=2, sig_min=[0.01, 0.01], sig_max=[0.03, 0.03], iseed=8),
    }[variant]
    cfg = helpers.attach_obs_and_rinv(helpers.small_config(**kw))
    got, _ = _compare(cfg)
    assert np.all(got["noise_sigma"] >= np.asarray(cfg.sig_min)) and np.all(got["noise_sigma"] <= np.asarray(cfg.sig_max))


def test_run_writes_reference_file_set(tmp_path):
    """params.in -> test_vel, test_traceNN, test_traceNNwn with the SAC header words of make_syn.f90:121-137."""
    from test_host_io import write_sample_problem
    work = tmp_path / "job"
    work.mkdir()
    params = write_sample_problem(work)
    res = make_syn.run(params, out_dir=str(tmp_path / "out"), verbose=False)
    cfg = rio.load_problem(params)
    lines = open(tmp_path / "out" / "test_vel").read().splitlines()
    assert len(lines) == res["nlay"] and all(len(line) == 104 for line in lines)
    first = [float(x) for x in lines[0].split()]
    np.testing.assert_allclose(first, [res["alpha"][0], res["beta"][0], res["rho"][0], res["h"][0]], rtol=1e-15)
    for t in range(cfg.ntrc):
        for suffix, key in (("", "rft"), ("wn", "noisy")):
            raw = np.fromfile(tmp_path / "out" / f"test_trace{t + 1:02d}{suffix}", dtype=np.float32)
            ints = raw.view(np.int32)
            assert raw.size == 158 + cfg.nsmp and ints[79] == cfg.nsmp and ints[76] == 6 and ints[85] == 1 and ints[105] == 1
            assert raw[0] == np.float32(cfg.delta) and raw[5] == np.float32(cfg.t_start)
            assert np.array_equal(raw[158:], res[key][t].astype(np.float32))
