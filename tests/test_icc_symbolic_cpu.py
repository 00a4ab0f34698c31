"""Host logic of the planned device ICC(0) (csrc/host/iccsym.c): the factor layout, the reference's merge order,
the dependency levels and the column view are index work -- checked here on the CPU, index-exact, against the oracle's
restatement of the reference (itself pinned to PCApply(PCICC) fixtures) and the numpy prototype that proved the schedule
reproduces MatCholeskyFactorNumeric_SeqAIJ bit for bit."""
import ctypes as C

import numpy as np
import pytest

from harness import petsc

vp = C.c_void_p


def _p(a):
    return a.ctypes.data_as(vp)


def host_symbolic(ai, aj):
    L = petsc.lib()
    ai, aj = np.ascontiguousarray(ai, np.int32), np.ascontiguousarray(aj, np.int32)
    n = len(ai) - 1
    ui, udiag, uj = np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32), np.empty(max(len(aj), 1), np.int32)
    petsc.chk(L.PetscB200ICC0Symbolic(n, _p(ai), _p(aj), _p(ui), _p(uj), _p(udiag)))
    nzu = int(ui[n])
    m = max(nzu - n, 1)
    mptr, mrow, mpos, level, nlev = np.empty(n + 1, np.int32), np.empty(m, np.int32), np.empty(m, np.int32), np.empty(max(n, 1), np.int32), C.c_int()
    petsc.chk(L.PetscB200ICC0MergeSchedule(n, _p(ui), _p(uj), _p(mptr), _p(mrow), _p(mpos), _p(level), C.byref(nlev)))
    tptr, trow, tpos = np.empty(n + 1, np.int32), np.empty(m, np.int32), np.empty(m, np.int32)
    petsc.chk(L.PetscB200ICC0ColumnView(n, _p(ui), _p(uj), _p(tptr), _p(trow), _p(tpos)))
    k = nzu - n
    return dict(ui=ui, uj=uj[:nzu], udiag=udiag[:n], mptr=mptr, mrow=mrow[:k], mpos=mpos[:k], level=level[:n], nlevels=nlev.value,
                tptr=tptr, trow=trow[:k], tpos=tpos[:k])


@pytest.mark.parametrize("gen", ["lap5", "lap7", "lap27", "tridiag", "diag"])
def test_icc0_symbolic_schedule_matches_oracle(oracle, gen):
    from oracle import icc_schedule as S
    if gen == "lap5":
        ai, aj, aa = oracle.lap5(13, 9)
    elif gen == "lap7":
        ai, aj, aa = oracle.lap7(7, 6, 5)
    elif gen == "lap27":
        ai, aj, aa = oracle.lap27(6)
    elif gen == "tridiag":
        n = 40
        ai = np.concatenate([[0], np.cumsum([2] + [3] * (n - 2) + [2])]).astype(np.int32)
        aj = np.concatenate([[0, 1]] + [[i - 1, i, i + 1] for i in range(1, n - 1)] + [[n - 2, n - 1]]).astype(np.int32)
        aa = np.where(aj == np.repeat(np.arange(n), np.diff(ai)), 2.0, -1.0)
    else:
        n = 9
        ai, aj, aa = np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.full(n, 3.0)
    h = host_symbolic(ai, aj)
    ui, uj, udiag, ua = oracle.icc0(ai, aj, aa)
    assert np.array_equal(h["ui"], ui) and np.array_equal(h["uj"], uj) and np.array_equal(h["udiag"], udiag)   # factor layout
    ptr, rows, pos = S.merge_schedule(ui, uj)
    assert np.array_equal(h["mptr"], ptr) and np.array_equal(h["mrow"], rows) and np.array_equal(h["mpos"], pos)  # merge order
    lev = S.levels(ptr, rows)
    assert np.array_equal(h["level"], lev) and h["nlevels"] == (int(lev.max()) + 1 if len(lev) else 0)
    # column view: ascending rows inside every column, positions point at the right entries
    n = len(ui) - 1
    for c in range(n):
        seg = slice(h["tptr"][c], h["tptr"][c + 1])
        assert np.all(np.diff(h["trow"][seg]) > 0)
        assert np.all(uj[h["tpos"][seg]] == c)
        assert all(ui[i] <= t < ui[i + 1] - 1 for i, t in zip(h["trow"][seg], h["tpos"][seg]))
    assert h["tptr"][n] == len(uj) - n
    # the C schedule drives the prototype's row-wise numeric phase to the oracle's (= the reference's) factor, bit for bit
    orig, final = S.numeric_rowwise(ai, aj, aa, ui, uj, udiag, (h["mptr"].astype(np.int64), h["mrow"].astype(np.int64), h["mpos"].astype(np.int64)))
    assert np.array_equal(final, ua)


def test_icc0_symbolic_missing_diagonal_is_the_reference_error():
    L = petsc.lib()
    ai, aj = np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32)       # no diagonal entries at all
    ui, uj, ud = np.empty(3, np.int32), np.empty(4, np.int32), np.empty(2, np.int32)
    rc = L.PetscB200ICC0Symbolic(2, _p(ai), _p(aj), _p(ui), _p(uj), _p(ud))
    assert rc == 73 and b"missing diagonal" in L.PetscB200GetLastErrorMessage()  # PETSC_ERR_ARG_WRONGSTATE (aijfact.c:2071)
