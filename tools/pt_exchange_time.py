"""PT iterations/s of the distributed run (launch under torchrun, one process per GPU); RFINV_PT_EXCHANGE=nccl forces the
NCCL all-gather, default = peer-memory exchange where the GPUs can map each other.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/pt_exchange_time.py [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from rf_inv_b200 import workloads, capi
if os.environ.get("RFINV_LIB"):
    capi._lib = capi.load(os.environ["RFINV_LIB"])
from rf_inv_b200.pt import ParallelTempering

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cfg = workloads.make_config("target")
cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp)); cfg.r_inv = workloads.lapack_r_inv(cfg)
pt = ParallelTempering(cfg, 1024 * world, device=local_rank, world=world, rank=rank)
run = (lambda n: pt.run(n)) if world == 1 else (lambda n: pt.run_distributed(n, dist, torch))
run(10)
best = 0.0
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); run(n_it); dt = time.perf_counter() - t0
    best = max(best, n_it / dt)
st = pt.state()
print(f"[process {rank}] forward evaluations {pt.counters()['n_eval']}, mean k {st['k'].mean():.3f}, sum of (k + 1) over chains {int((st['k'] + 1).sum())}", flush=True)
if rank == 0:
    print(f"world {world} exchange {pt.exchange_mode}: {best:.1f} iterations/s ({16384 * world} chains)", flush=True)
pt.close()
if world > 1:
    dist.destroy_process_group()
