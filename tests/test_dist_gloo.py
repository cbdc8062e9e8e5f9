"""world_size-2 gloo test (CPU) of the multi-process exchange protocol: rank split, swap-table layout, one
all-gather per iteration, replicated swap decision.  The chain arithmetic itself is the oracle's here; the
same protocol with the CUDA kernels is tested in tests/test_gpu_pt.py::test_split_driver_matches_single_driver."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nproc_total, nchains, n_iter, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import rfinv_oracle as pyo
    from rf_inv_b200.pt import build_table, decode_swap, split_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, count = split_ranks(nproc_total, world, rank)
    n_local = count * nchains
    rng = np.random.default_rng(1234)                       # same synthetic chain state on both processes
    temps_all = np.exp(rng.uniform(0, np.log(15.0), nproc_total * nchains))
    temps_all[::nchains] = 1.0
    logl_all = rng.normal(-500.0, 50.0, nproc_total * nchains)
    streams = [pyo.MT19937(pyo.rank_seed(12345678, r)) for r in range(nproc_total)]
    decisions = []
    for it in range(n_iter):
        logl_all = logl_all + np.random.default_rng(it).normal(0, 5.0, logl_all.shape)   # "chain steps"
        pair = (-1, -1)
        if begin == 0:                                      # owner of virtual rank 0 draws the pair (pt_mcmc.f90:501-507)
            n_all = nproc_total * nchains
            i1 = int(streams[0].grnd() * n_all)
            while True:
                i2 = int(streams[0].grnd() * n_all)
                if i2 != i1:
                    break
            pair = (i1, i2)
        peeks = []
        for r in range(begin, begin + count):               # peek = next uniform without consuming
            st = streams[r]
            save = (list(st.mt), st.mti)
            peeks.append(st.grnd())
            st.mt, st.mti = save[0], save[1]
        sl = slice(begin * nchains, (begin + count) * nchains)
        table = torch.from_numpy(build_table(temps_all[sl], logl_all[sl], peeks, pair))
        out = torch.empty(world * table.numel(), dtype=torch.float64)
        dist.all_gather_into_tensor(out, table)
        i1, i2, yn, own1, l1, own2, l2 = decode_swap(out.numpy(), world, n_local, count, nchains)
        if own1 == rank:
            streams[begin + l1 // nchains].grnd()           # rank1's stream consumed the judge_pt uniform
        if yn:                                              # every process applies the same decision to its replica
            temps_all[i1], temps_all[i2] = temps_all[i2], temps_all[i1]
        decisions.append((i1, i2, int(yn)))
    q.put((rank, decisions, temps_all.copy()))
    dist.barrier()
    dist.destroy_process_group()


def _reference_sequence(nproc_total, nchains, n_iter):
    """Single-process restatement of the same loop (pt_control, src/pt_mcmc.f90:498-571)."""
    import math
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import rfinv_oracle as pyo
    rng = np.random.default_rng(1234)
    temps = np.exp(rng.uniform(0, np.log(15.0), nproc_total * nchains))
    temps[::nchains] = 1.0
    logl = rng.normal(-500.0, 50.0, nproc_total * nchains)
    streams = [pyo.MT19937(pyo.rank_seed(12345678, r)) for r in range(nproc_total)]
    out = []
    for it in range(n_iter):
        logl = logl + np.random.default_rng(it).normal(0, 5.0, logl.shape)
        n_all = nproc_total * nchains
        i1 = int(streams[0].grnd() * n_all)
        while True:
            i2 = int(streams[0].grnd() * n_all)
            if i2 != i1:
                break
        u = streams[i1 // nchains].grnd()
        del_s = (logl[i2] - logl[i1]) * (1.0 / temps[i1] - 1.0 / temps[i2])
        yn = (math.log(u) if u > 0 else -math.inf) <= del_s
        if yn:
            temps[i1], temps[i2] = temps[i2], temps[i1]
        out.append((i1, i2, int(yn)))
    return out, temps


def test_two_process_swap_protocol_matches_single_process():
    nproc_total, nchains, n_iter, world = 6, 4, 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nproc_total, nchains, n_iter, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_seq, ref_temps = _reference_sequence(nproc_total, nchains, n_iter)
    for rank, decisions, temps in results:
        assert decisions == ref_seq
        assert np.array_equal(temps, ref_temps)
    assert sum(d[2] for d in ref_seq) > 0 and any(d[0] // (3 * nchains) != d[1] // (3 * nchains) for d in ref_seq)


def test_split_ranks():
    from rf_inv_b200.pt import split_ranks
    assert split_ranks(4096, 8, 3) == (1536, 512)
    assert split_ranks(20, 1, 0) == (0, 20)
    with pytest.raises(ValueError):
        split_ranks(20, 8, 0)
