"""The distributed parallel-tempering run on the real multi-GPU path (one process per GPU under torchrun): accept / reject
flags, proposal types, swaps, final state and job-wide bookkeeping of N processes driving rfinv_pt_run_distributed equal those
of ONE process holding every virtual rank -- once with the swap tables stored into peer memory (CUDA IPC over NVLink) and once
with the NCCL all-gather.  Needs two GPUs: skipped on a single-GPU box (tools/dist_check.py and bench.py at N > 1 run the
same check; results under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from rf_inv_b200 import capi
    return int(capi.load().rfinv_device_count())


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_processes_reproduce_the_single_process_run(exchange, tmp_path):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, RFINV_PT_EXCHANGE=exchange, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531" if exchange == "peer" else "29532", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([line for line in r.stdout.splitlines() if line.startswith("{")][-1])
    for name, res in out.items():
        assert res["identical"], (name, res)
        assert res["exchange"] == ("peer memory" if exchange == "peer" else "nccl all-gather")
