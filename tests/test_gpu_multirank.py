"""Multi-rank Fock build on the GPU: the sharded build (grid slices, J/K work items, aux-shell slices,
one packed all-reduce) equals the unsharded one to <= 1e-11.  With >= 2 GPUs the ranks run one per GPU
over NCCL (torchrun); on a single-GPU box two ranks share the GPU and reduce over gloo, which still
exercises every sharded kernel launch."""
import os
import socket
import subprocess
import sys
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_match_one_rank(cuda):
    if torch.cuda.device_count() >= 2:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
               os.path.join(ROOT, "tests", "multigpu_check.py")]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "MULTIRANK_MAXDIFF" in out.stdout
    else:
        code = ("import sys; sys.path.insert(0, %r); import torch.multiprocessing as mp; "
                "from tests.multigpu_check import _spawned; "
                "mp.spawn(_spawned, args=(2, %d, 'gloo'), nprocs=2, join=True)" % (ROOT, _free_port()))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "MULTIRANK_MAXDIFF" in out.stdout
