"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
from rf_inv_b200 import build as rbuild
from rf_inv_b200 import capi, workloads
from rf_inv_b200.config import RFConfig, RfinvConfigC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rfinv_b200.h")


@pytest.fixture(scope="module")
def lib():
    rbuild.build()
    return capi.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rfinv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rfinv_b200.h but not exported"
        assert s in capi.SIGNATURES, f"{s} has no ctypes signature in rf_inv_b200/capi.py"
    declared = int(re.search(r"#define\s+RFINV_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert lib.rfinv_abi_version() == declared >= 3


def test_config_struct_layout_matches_header():
    # field order of the ctypes mirror == field order of struct rfinv_config
    src = open(HEADER).read()
    body = re.search(r"typedef struct rfinv_config \{(.*?)\} rfinv_config;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        decl = re.sub(r"^(const\s+)?(int32_t|double)\s*\*?", "", stmt)
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == [f[0] for f in RfinvConfigC._fields_]
    assert C.sizeof(RfinvConfigC) % 8 == 0


def test_no_gpu_means_loud_failure_not_cpu_fallback(lib):
    n = lib.rfinv_device_count()
    if n > 0:
        pytest.skip("a CUDA device is present")
    cfg = helpers.small_config()
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp))
    cfg.r_inv = np.zeros((cfg.ntrc, cfg.nsmp, cfg.nsmp))
    c = cfg.to_c()
    h = C.c_void_p()
    st = lib.rfinv_create(C.byref(c), 0, C.byref(h))
    assert st == capi.RFINV_ERR_CUDA and not h.value
    assert len(lib.rfinv_last_error()) > 0
    from rf_inv_b200.evaluator import Evaluator
    with pytest.raises(capi.RfinvError):
        Evaluator(cfg)


def test_create_rejects_bad_configs_before_touching_cuda(lib):
    cfg = helpers.small_config()
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp))
    c = cfg.to_c()
    c.nfft = 3000                                  # lengths that are not powers of two go through Bluestein up to 2048
    h = C.c_void_p()
    assert lib.rfinv_create(C.byref(c), 0, C.byref(h)) == capi.RFINV_ERR_ARG
    assert b"any other length in [64,2048]" in lib.rfinv_last_error()
    c = cfg.to_c()
    c.nfft = 8192
    assert lib.rfinv_create(C.byref(c), 0, C.byref(h)) == capi.RFINV_ERR_ARG
    c = cfg.to_c()
    c.deconv_mode = 2                              # src/params.f90:195-199
    assert lib.rfinv_create(C.byref(c), 0, C.byref(h)) == capi.RFINV_ERR_ARG
    assert b"deconv_mode must be either 0 or 1" in lib.rfinv_last_error()
    c = cfg.to_c()
    c.vp_mode = 3                                  # src/pt_mcmc.f90:329-335
    assert lib.rfinv_create(C.byref(c), 0, C.byref(h)) == capi.RFINV_ERR_ARG


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        capi.load(str(tmp_path / "nope.so"))


def test_config_validation_and_modes():
    cfg = helpers.small_config(sig_min=[0.01, 0.005], sig_max=[0.01, 0.05])
    assert cfg.sig_mode == [0, 1]
    assert not cfg.is_ray_common
    assert helpers.small_config(rayps=[0.06, 0.06]).is_ray_common
    assert not helpers.small_config(rayps=[0.06, 0.06], ipha=[1, -1]).is_ray_common
    with pytest.raises(ValueError):
        helpers.small_config(nfft=3000).to_c()
    helpers.small_config(nfft=300).to_c()          # FFTW accepts any length (src/fftw.f90:43-45); so does the library up to 2048
    with pytest.raises(ValueError):
        helpers.small_config(rayps=[0.06]).to_c()
    # sig_mode threshold is the float32 literal 1.0e-5 (src/params.f90:262)
    eps = float(np.float32(1e-5))
    assert helpers.small_config(sig_min=[0.01, 0.01], sig_max=[0.01, 0.01 + 0.5 * eps]).sig_mode == [0, 0]


def test_workloads_are_deterministic_and_valid():
    for name in ("sample", "c2", "c3", "c4", "c4_laplace", "c5", "target"):
        cfg = workloads.make_config(name)
        cfg.validate()
        m1 = workloads.draw_models(cfg, 16, seed=4)
        m2 = workloads.draw_models(cfg, 16, seed=4)
        assert all(np.array_equal(m1[k], m2[k]) for k in m1)
        assert workloads.models_valid(cfg, m1["k"], m1["z"], m1["dvp"], m1["dvs"]).all()
        assert (m1["k"] >= cfg.k_min).all() and (m1["k"] < cfg.k_max).all()
    tgt = workloads.make_config("target")
    assert (tgt.nfft, tgt.k_max, tgt.ntrc) == (1024, 30, 3)
    soa = workloads.to_soa(m1)
    assert soa["z"].shape == (cfg.k_max - 1, 16)


def test_models_valid_agrees_with_oracle_format_model():
    import oracle_c
    cfg = workloads.make_config("c3")
    rng = np.random.default_rng(0)
    m = workloads.draw_models(cfg, 40, seed=9)
    m["z"][:20] = rng.uniform(0, 6, m["z"][:20].shape)       # many invalid (above sea floor / thin)
    ok = workloads.models_valid(cfg, m["k"], m["z"], m["dvp"], m["dvs"])
    for i in range(40):
        assert ok[i] == oracle_c.format_model(cfg, int(m["k"][i]), m["z"][i], m["dvp"][i], m["dvs"][i])[5]


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference (the CPU restatement timed on the host cores) is run by the driver next to the CUDA arm:
    exactly one JSON line on stdout, with the same metric / unit / config keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "48"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "forward+likelihood evals/sec" and j["unit"] == "evals/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["config"]["workload"] == "target"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
