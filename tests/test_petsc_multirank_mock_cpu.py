"""The multi-rank path of the PETSc plugin (mpiaijb200 / mpib200: one process per rank joined through PETSCB200_NRANKS / RANK /
NCCL_ID) in the build container: 2 and 3 driver processes, each bound (LD_PRELOAD) to the host test double of the C ABI, whose
collectives go through files in /dev/shm.  The same checks bench.py --gpus N runs on N GPUs (tools/plugin_parity.py): garray and the
diagonal / off-diagonal blocks index-exact against the restatement of MatSetUpMultiply_MPIAIJ, rank-local MatMult bit-exact, halo
MatMult and the reverse scatter of MatMultTranspose against the sequential product, fused MatMult+PCJACOBI == unfused, all-reduced
reductions, GMRES + PCBJACOBI/ILU(0) (ex2_2.out to all printed digits on 2 ranks), GMRES + Jacobi, CG + PCBJACOBI at 1e-12 * r0.
What runs is the plugin's host logic and the reference's own KSPSolve; the kernels and NCCL are the GPU tests' business."""
import binascii
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
MOCK = os.path.join(ROOT, "tests", "mock", "libb200mock.so")


@pytest.mark.parametrize("size", [2, 3])
def test_plugin_parity_block_on_several_ranks_through_the_mock_device(oracle, size):
    from petsc_b200 import _capi, petsc_driver as drv
    n = C.c_int(0)
    if _capi.lib().b200DeviceCount(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is visible: bench.py --gpus N runs this block on the real library there")
    if not drv.available():
        pytest.skip("baseline/_ref/petsc or petsc_plugin/b200_driver not built (needs the build container)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200mock_build", os.path.join(ROOT, "tests", "mock", "build.py"))
    mb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mb)
    mb.build()
    import plugin_parity as PP
    uid = binascii.hexlify(os.urandom(128)).decode()
    with tempfile.TemporaryDirectory(prefix="b200parity_") as d:
        cs = [PP.write(d, oracle, r, size) for r in range(size)]
        procs = []
        for r in range(size):
            env = dict(os.environ, LD_PRELOAD=MOCK, LD_LIBRARY_PATH=drv.BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), **drv.rank_env(r, size, uid))
            procs.append(subprocess.Popen([drv.DRIVER_EXE, "-parity", d, "-no_signal_handler", "-options_left", "0"], env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
        passed, failed = [0], []

        def ck(cond, what):
            if bool(cond):
                passed[0] += 1
            else:
                failed.append(str(what))
        for r in range(size):
            def gather(a, _r=r):
                # the same output file of every rank, concatenated in rank order
                for c in cs[_r]:
                    for fn in os.listdir(c["dir"]):
                        arr = np.fromfile(os.path.join(c["dir"], fn), dtype=np.float64) if fn.endswith(".f64") else None
                        if arr is not None and arr.shape == np.asarray(a).shape and np.array_equal(arr, a):
                            name = os.path.basename(os.path.dirname(c["dir"]))
                            return np.concatenate([np.fromfile(os.path.join(d, name, "rank%d" % q, fn), dtype=np.float64) for q in range(size)])
                raise AssertionError("gather: array not found among the outputs")
            PP.check(cs[r], oracle, r, size, ck, gather)
        assert not failed, failed[:10]
        assert passed[0] >= 100 * size
    for f in os.listdir("/dev/shm"):
        if f.startswith("b200mock_" + uid[:16]):
            import shutil
            shutil.rmtree(os.path.join("/dev/shm", f), ignore_errors=True)
