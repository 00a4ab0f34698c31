"""The symbolic merge-order schedule planned for the device ICC(0) factorisation (oracle/icc_schedule.py):
a row-wise recomputation that only follows the schedule reproduces the oracle's restatement of
MatCholeskyFactorNumeric_SeqAIJ -- and through it the reference's PCApply(PCICC) fixtures -- bit for bit."""
import glob
import os

import numpy as np
import pytest

from conftest import golden_path

OPS = [p for p in sorted(glob.glob(golden_path("ops_*.npz"))) if "ref_iccsolve" in np.load(p).files]


def load_matrix(O, g):
    return getattr(O, str(g["gen"]))(*[int(a) for a in g["args"]])


@pytest.mark.parametrize("path", OPS, ids=[os.path.basename(p)[:-4] for p in OPS])
def test_rowwise_icc_with_reference_merge_order_is_bit_exact(oracle, path):
    from oracle import icc_schedule as S
    g = np.load(path)
    ai, aj, aa = load_matrix(oracle, g)
    ui, uj, udiag, ua = oracle.icc0(ai, aj, aa)
    sched = S.merge_schedule(ui, uj)
    ptr, rows, pos = sched
    n = len(ui) - 1
    # every strictly-upper entry (i, c) is merged exactly once, into row c
    assert len(rows) == len(uj) - n
    assert all(int(uj[pos[q]]) == k for k in range(n) for q in range(ptr[k], ptr[k + 1]))
    orig, final = S.numeric_rowwise(ai, aj, aa, ui, uj, udiag, sched)
    assert np.array_equal(final, ua)
    # and the factor solves to the reference's PCApply(PCICC) output
    assert np.array_equal(oracle.matsolve_icc(ui, uj, udiag, final, g["x"]), g["ref_iccsolve"])
    # the scatter-free (gather) form of the two sweeps, as a level-scheduled device kernel will run them
    assert np.array_equal(S.solve_gather(ui, uj, udiag, final, g["x"]), g["ref_iccsolve"])
    # the schedule's dependency levels: rows only wait for rows of strictly lower level
    lev = S.levels(ptr, rows)
    assert all(lev[rows[q]] < lev[k] for k in range(n) for q in range(ptr[k], ptr[k + 1]))
    assert lev.max() + 1 < n  # there is parallelism to schedule


def test_merge_order_is_not_simply_ascending(oracle):
    """The reference's lists are LIFO: a schedule that walked contributors in ascending row order would change the rounding.
    (Guards the prototype against being 'simplified'.)"""
    from oracle import icc_schedule as S
    ai, aj, aa = oracle.lap27(5)
    ui, uj, udiag, ua = oracle.icc0(ai, aj, aa)
    ptr, rows, pos = S.merge_schedule(ui, uj)
    asc = all(np.all(np.diff(rows[ptr[k]:ptr[k + 1]]) > 0) for k in range(len(ui) - 1))
    assert not asc
