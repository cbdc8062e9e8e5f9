"""CPU tests of the file formats either side of the path (host services of the C-ABI library)."""
import os

import numpy as np
import pytest

import helpers
from rf_inv_b200 import capi, io as rio, workloads

G = np.load(os.path.join(helpers.GOLDEN, "sample_syn.npz"))


def write_sample_problem(tmp_path, comment_style=True):
    """Re-creates sample_syn/ from the committed fixture values: params.in (own comments), SAC traces, velmod."""
    d = tmp_path
    (d / "data").mkdir(); (d / "model").mkdir(); (d / "rslt").mkdir()
    rio.write_sac(str(d / "data" / "sample_1.trc"), G["trc1"], float(G["delta"]), float(G["b"]))
    rio.write_sac(str(d / "data" / "sample_2.trc"), G["trc2"], float(G["delta"]), float(G["b"]))
    with open(d / "model" / "sample.velmod", "w") as f:
        for z, vp, vs in G["ref_velmod"]:
            f.write(f"{z:.1f}   {vp:.2f} {vs:.2f}\n")
    lines = []
    for i, v in enumerate(G["params"]):
        if comment_style:
            lines.append(f"# value {i}")
            lines.append("   #   indented comment lines are skipped too (adjustl, src/params.f90:397)")
        lines.append(("   " if i % 3 == 0 else "") + str(v))
    p = d / "params.in"
    p.write_text("\n".join(lines) + "\n")
    return str(p)


def test_load_sample_problem_matches_shipped_values(tmp_path):
    p = write_sample_problem(tmp_path)
    cfg = rio.load_problem(p)
    ref = workloads.make_config("sample")
    assert (cfg.nburn, cfg.niter, cfg.ncorr, cfg.nchains, cfg.ncool, cfg.iseed) == (3000, 8000, 10, 5, 1, 12345678)
    assert cfg.t_high == 15.0 and cfg.ntrc == 2 and cfg.nfft == 256 and cfg.deconv_mode == 0 and cfg.sdep == 2.0
    assert cfg.rayps == [0.06, 0.08] and cfg.a_gus == [4.0, 4.0] and cfg.ipha == [1, 1]
    assert cfg.nsmp == 101 and cfg.delta == float(np.float32(0.05))            # from the SAC header (float32)
    assert np.array_equal(cfg.obs[0], G["trc1"].astype(np.float64)) and np.array_equal(cfg.obs[1], G["trc2"].astype(np.float64))
    assert len(cfg.vp_ref) == 61 and cfg.z_ref_min == 0.0 and cfg.dz_ref == 0.5 and cfg.vs_ref[0] == 2.89
    for name in ("vp_mode", "k_min", "k_max", "z_min", "z_max", "h_min", "prior_mode", "dvs_prior", "dvp_prior", "dev_z",
                 "dev_dvs", "dev_dvp", "dev_sig", "nbin_z", "nbin_vs", "nbin_vp", "nbin_vpvs", "nbin_sig", "nbin_amp",
                 "amp_min", "amp_max", "vp_min", "vp_max", "vs_min", "vs_max", "vpvs_min", "vpvs_max"):
        assert getattr(cfg, name) == getattr(ref, name), name
    assert cfg.sig_min == [0.01, 0.01] and cfg.sig_max == [0.01, 0.01] and cfg.sig_mode == [0, 0]
    assert cfg.out_dir.endswith("rslt") and cfg.t_end == 5.0


def test_sac_window_and_side_copies(tmp_path):
    p = write_sample_problem(tmp_path)
    txt = open(p).read().replace("0.0 5.0", "1.0 3.0")           # T_START T_END: cut a window out of the trace
    open(p, "w").write(txt)
    cfg = rio.load_problem(p)
    assert cfg.nsmp == 41 and cfg.t_start == 1.0
    assert np.array_equal(cfg.obs[0], G["trc1"][20:61].astype(np.float64))
    rio.write_side_copies(p, str(tmp_path / "rslt"), str(tmp_path))
    copy = [l.strip() for l in open(tmp_path / "rslt" / "params.in.copy")]
    assert len(copy) == len(G["params"]) and copy[1] == "3000"
    rows = np.loadtxt(tmp_path / "input01")
    assert rows.shape == (41, 2) and rows[0, 0] == 1.0 and np.allclose(rows[:, 1], cfg.obs[0], rtol=0, atol=0)


def test_reader_errors(tmp_path):
    lib = capi.load()
    with pytest.raises(capi.RfinvError) as e:
        rio.load_problem(str(tmp_path / "missing.in"))
    assert e.value.status == capi.RFINV_ERR_IO
    p = write_sample_problem(tmp_path, comment_style=False)
    bad = open(p).read().splitlines()
    bad[19] = "2"                                                   # DECONV_MODE (src/params.f90:195-199)
    (tmp_path / "bad.in").write_text("\n".join(bad) + "\n")
    with pytest.raises(capi.RfinvError) as e:
        rio.load_problem(str(tmp_path / "bad.in"))
    assert "deconv_mode must be either 0 or 1" in str(e.value)
    ok = open(p).read().splitlines()
    ok[20] = "2.0 1.5"                                              # SEA_DEP BOREHOLE_DEP (src/params.f90:203-213, commented out there)
    (tmp_path / "bore.in").write_text("\n".join(ok) + "\n")
    cfg = rio.load_problem(str(tmp_path / "bore.in"))
    assert cfg.sdep == 2.0 and cfg.bdep == 1.5
    assert rio.load_problem(p).bdep == 0.0                          # the second number is optional
    ok[20] = "2.0 -1.0"
    (tmp_path / "bore_bad.in").write_text("\n".join(ok) + "\n")
    with pytest.raises(capi.RfinvError) as e:
        rio.load_problem(str(tmp_path / "bore_bad.in"))
    assert "BOREHOLE_DEP must be positive" in str(e.value)
    with open(tmp_path / "model" / "sample.velmod", "a") as f:      # non-constant depth increment (src/model.f90:131-138)
        f.write("31.0 5.0 2.89\n")
    with pytest.raises(capi.RfinvError) as e:
        rio.load_problem(p)
    assert "Depth increment must be constant" in str(e.value)


def test_write_outputs_formats(tmp_path):
    cfg = helpers.small_config(sig_min=[0.01, 0.005], sig_max=[0.01, 0.05], nbin_z=4, nbin_vs=3, nbin_vp=2, nbin_vpvs=2,
                               nbin_sig=5, nbin_amp=4, nsmp=3, nfft=64, k_max=5, niter=10, nburn=2, ncool=1)
    cfg.obs = np.zeros((2, 3))
    rng = np.random.default_rng(0)
    hist = dict(nmod=8, nk=rng.integers(0, 5, 5), nz=rng.integers(0, 9, 4), nsig=rng.integers(0, 9, (2, 5)),
                namp=rng.integers(0, 9, (2, 3, 4)), nvpz=rng.integers(0, 9, (2, 4)), nvsz=rng.integers(0, 9, (3, 4)),
                nvpvsz=rng.integers(0, 9, (2, 4)), vp_mean=rng.uniform(30, 50, 4), vs_mean=rng.uniform(10, 30, 4),
                vpvs_mean=rng.uniform(10, 20, 4))
    lh = rng.normal(-500, 10, 12)
    vp_model = rng.uniform(4, 6, (3, 4)); vs_model = rng.uniform(2, 4, (3, 4))
    vs_model[1, 0] = -999.9                                          # unrecorded slot: skipped (src/mcmc_out.f90:119-121)
    out = tmp_path / "rslt"
    rio.write_outputs(cfg, str(out), 2, hist, lh, vp_model, vs_model)
    names = {"all_models", "likelihood", "num_interface.ppd", "syn_trace.ppd", "interface_depth.ppd", "sigma.ppd", "vs_z.ppd",
             "vp_z.ppd", "vpvs_z.ppd", "vs_z.mean", "vp_z.mean", "vpvs_z.mean"}
    assert names <= set(os.listdir(out))
    lk = np.loadtxt(out / "likelihood")
    assert lk.shape == (12, 2) and np.allclose(lk[:, 1], lh / 2.0, rtol=1e-15) and lk[0, 0] == 1
    nk = np.loadtxt(out / "num_interface.ppd")
    assert nk.shape == (4, 2) and np.allclose(nk[:, 1], hist["nk"][:4] / 8.0)
    syn = open(out / "syn_trace.ppd").read().splitlines()
    assert len(syn) == 2 * 3 * 4 and len(syn[0]) == 36                 # '(3F10.5,I6)'
    t0, amp0, p0, itrc0 = syn[0][:10], syn[0][10:20], syn[0][20:30], syn[0][30:]
    assert float(t0) == 0.0 and abs(float(amp0) - (-0.8 + 0.5 * 0.4)) < 1e-5 and int(itrc0) == 1
    assert abs(float(p0) - hist["namp"][0, 0, 0] / 8.0) < 1e-5
    sig = np.loadtxt(out / "sigma.ppd")
    assert sig.shape == (5, 3) and (sig[:, 2] == 2).all()              # only the solved trace is written
    vs = open(out / "vs_z.ppd").read().splitlines()
    assert len(vs) == 3 * 4 and len(vs[0]) == 30                       # '(3F10.5)'
    mean = np.loadtxt(out / "vp_z.mean")
    assert mean.shape == (4, 2) and np.allclose(mean[:, 0], hist["vp_mean"] / 8.0, atol=1e-5)
    blocks = [b for b in open(out / "all_models").read().split("\n\n") if b.strip()]
    assert len(blocks) == 2
    first = np.array([[float(x) for x in l.split()] for l in blocks[0].strip().splitlines()])
    assert first.shape == (4, 3) and np.allclose(first[:, 1], vp_model[0], rtol=1e-15) and np.allclose(first[:, 0], (np.arange(4) + 0.5) * 5.0)
    # list-directed reals look like gfortran's: 17 significant digits, 26-column fields
    line = open(out / "interface_depth.ppd").readline().rstrip("\n")
    assert len(line) == 52 and line.startswith("   2.5000000000000000     ")


def test_list_directed_real_matches_gfortran_layout():
    """make_syn's test_vel is written with list-directed output (src/make_syn.f90:73): 26 columns per real(8)."""
    from rf_inv_b200.make_syn import _ld_real
    assert _ld_real(5.0) == "   5.0000000000000000     "
    assert _ld_real(2.5347508187769563) == "   2.5347508187769563     "
    assert _ld_real(999.0) == "   999.00000000000000     "
    assert _ld_real(0.125) == "  0.12500000000000000     "
    assert _ld_real(-999.0) == "  -999.00000000000000     "
    assert _ld_real(1.0e-5) == "   1.0000000000000001E-005"
    assert all(len(_ld_real(v)) == 26 for v in (0.0, 3.14, 1e20, -2.5e-7))
