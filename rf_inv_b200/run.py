"""Driver with the reference's command line (src/rf_inv.f90:28-108):  python -m rf_inv_b200.run [params.in] --nproc N

``--nproc`` is the number of MPI ranks the reference would be started with (``mpirun -np N``): N virtual ranks of
N_CHAINS chains each, all resident on the GPU(s).  Under torchrun (one process per GPU) the virtual ranks are split
over the processes and the per-iteration swap exchange is one NCCL all-gather issued by the library itself.  Writes the reference's output files
into OUT_DIR."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import io as rio
from . import workloads
from .pt import ParallelTempering


def run(params_path: str, nproc: int, verbose: bool = True, n_iter: int = None):
    cfg = rio.load_problem(params_path)
    cfg.r_inv = workloads.lapack_r_inv(cfg)          # init_r_inv: LAPACK dgesvd like the reference
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # torch.distributed only carries the library's communicator id between the processes (gloo: no second NCCL communicator)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    if rank == 0:
        os.makedirs(cfg.out_dir, exist_ok=True)
        rio.write_side_copies(params_path, cfg.out_dir, os.path.dirname(os.path.abspath(params_path)))
    pt = ParallelTempering(cfg, nproc, device=local_rank, world=world, rank=rank)
    n_tot = cfg.nburn + cfg.niter if n_iter is None else n_iter
    done = 0
    while done < n_tot:                                   # progress line every N_CORR iterations (src/pt_mcmc.f90:489-491)
        step = min(max(cfg.ncorr, 1) * 100, n_tot - done)
        if world == 1:
            pt.run(step)
        else:
            pt.run_distributed(step, dist, torch)
        done += step
        if verbose and rank == 0:
            print(f" Iteration #: {done} / {n_tot}", flush=True)
    if world > 1:                                         # the reference's mpi_reduce / mpi_gather (src/mcmc_out.f90:52-93)
        pt.reduce_outputs()                               # ncclReduce / ncclSend+Recv inside the library: process 0 is job-wide now
    hist, cnt = pt.hist(), pt.counters()
    vp_model, vs_model = pt.models()
    lh = cnt["likelihood_hist"]
    if rank == 0:
        if verbose:                                       # summary, src/mcmc_out.f90:99-107
            labels = ["Birth proposal", "Death proposal", "Moving interface depth proposal", "Perturbing dVs proposal"]
            if cfg.vp_mode == 1:
                labels.append("Perturbing dVs proposal")  # sic: the reference labels the dVp proposal like this (pt_mcmc.f90:386)
            if any(cfg.sig_mode):
                labels.append("Perturbing sigma proposal")
            print(" --- Summary ---")
            print(f" # of sampled models: {hist['nmod']}")
            for lab, a, n in zip(labels, cnt["naccept"], cnt["nprop"]):
                print(f" # of {lab}: {int(a)} / {int(n)}")
        rio.write_outputs(cfg, cfg.out_dir, nproc, hist, lh, vp_model, vs_model)
    pt.close()
    if world > 1:
        dist.destroy_process_group()
    return hist, cnt


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("params", nargs="?", default="params.in")
    ap.add_argument("--nproc", type=int, default=20, help="number of virtual MPI ranks (mpirun -np of the reference)")
    ap.add_argument("--iters", type=int, default=None, help="override N_BURN + N_ITER (testing)")
    a = ap.parse_args(argv)
    run(a.params, a.nproc, n_iter=a.iters)


if __name__ == "__main__":
    main()
