"""Multi-GPU parity (needs >= 2 devices on the box; skipped otherwise): see tests/_mpi_nccl_worker.py."""
import ctypes as C
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ndev():
    from petsc_b200 import _capi
    n = C.c_int(0)
    return n.value if _capi.lib().b200DeviceCount(C.byref(n)) == 0 else 0


@pytest.mark.parametrize("nproc", [2])
def test_mpiaij_nccl(oracle, nproc):
    if ndev() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1", "--master-port", "29633",
           os.path.join(ROOT, "tests", "_mpi_nccl_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("OK") == 4, out.stdout
