"""Host-side logic of the N > 1 path on CPUs: 2 gloo ranks exercise the K-transpose layout mirror and the slab-wise
restart files."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_k_transpose_layout_gloo_world2(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", TLAB_TMP=str(tmp_path))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "DIST_CPU_OK" in r.stdout


def test_slab_partition():
    from tlab_b200 import mpi
    assert mpi.slab(1024, 3, 8) == (128, 384)
    try:
        mpi.slab(10, 0, 4)
    except ValueError:
        pass
    else:
        raise AssertionError("uneven slabs must be rejected (tlab_mpi_procs.f90:44-58)")
