"""z-slab decomposition on 2 GPUs (NCCL K-transposes) against the single-domain oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_rk_steps(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "DIST_ERRS" in r.stdout


def test_single_rank_transposes_are_copies(cuda):
    """P = 1: TLabMPI_Trp_ExecK_* degenerate to copies (OPR_CHECK's round trip, opr_check.f90:46-64)."""
    import ctypes
    import torch
    from tlab_b200 import lib as tl
    L = tl.load()
    tl.check(L.tlab_mpi_init(0, 1, None))
    a = torch.randn(6 * 40, dtype=torch.float64, device=cuda)
    b = torch.zeros_like(a)
    c = torch.zeros_like(a)
    torch.cuda.synchronize()
    tl.check(L.tlab_trp_exec_k_forward(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), 40, 6, 0))
    tl.check(L.tlab_trp_exec_k_backward(ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(c.data_ptr()), 40, 6, 0))
    assert torch.equal(a, c)
