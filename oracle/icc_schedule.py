"""Design prototype for the device ICC(0) numeric phase -- TEST INFRASTRUCTURE, numpy/pure Python.

MatCholeskyFactorNumeric_SeqAIJ (aijfact.c:1701-1866) merges, into row k, the finished rows i < k with U(i,k) != 0 in the
order its linked lists (c2r / il) happen to hold them; that order fixes the floating-point association of every entry
of row k, so a parallel (level-scheduled) factorisation reproduces the reference bit for bit only if each row walks its
contributors in exactly that order.  The order depends on the sparsity pattern alone:

  merge_schedule(ui, uj)      simulates the lists on indices (no arithmetic) and returns, per row k, the contributing rows and
                              the position of U(i,k) inside row i -- the host-side SYMBOLIC step of the planned device path
                              (the analogue of b200Ilu0Symbolic's level schedule).
  numeric_rowwise(...)        recomputes every row k from A's row, the ORIGINAL (unscaled) entries of its contributors and
                              their inverted pivots only -- no shared work vector, no in-place rewriting of other rows -- i.e.
                              what one warp of the device kernel will do for its row once its contributors are done.
  levels(...)                 the dependency levels of that kernel (row k waits for its contributors).

tests/test_icc_schedule_cpu.py checks numeric_rowwise against oracle.c's restatement of the reference (ora_icc0_numeric),
which itself is pinned to PCApply(PCICC) fixtures produced by the reference: equal bit for bit.
"""
import numpy as np


def merge_schedule(ui, uj):
    """Per row k: (rows i merged into k, position of U(i,k) in row i), in the reference's list order."""
    n = len(ui) - 1
    c2r = [n] * (n + 1)
    il = [0] * (n + 1)
    ptr, rows, pos = [0], [], []
    for k in range(n):
        i = c2r[k]
        while i < k:
            nexti = c2r[i]
            ili = il[i]
            rows.append(i); pos.append(ili)
            jmin, jmax = ili + 1, ui[i + 1]
            if jmin < jmax:
                il[i] = jmin
                j = int(uj[jmin])
                c2r[i] = c2r[j]
                c2r[j] = i
            i = nexti
        jmin, jmax = ui[k], ui[k + 1] - 1
        if jmin < jmax:
            il[k] = jmin
            c = int(uj[jmin])
            c2r[k] = c2r[c]
            c2r[c] = k
        ptr.append(len(rows))
    return np.array(ptr, np.int64), np.array(rows, np.int64), np.array(pos, np.int64)


def solve_gather(ui, uj, udiag, final, b):
    """MatSolve_SeqSBAIJ_1_NaturalOrdering (sbaijfact2.c:2030-2065) without scatters, the form a level-scheduled device sweep
    needs: the forward sweep's x[c] += v(i,c) * x_i over rows i in ascending order is a GATHER along column c of U (= row c
    of the transposed pattern, ascending i) of the UNSCALED y_i, followed by one multiply with 1/D(c); the backward sweep
    gathers along row i from its last off-diagonal entry to its first."""
    n = len(ui) - 1
    # transposed pattern of the strictly upper part: for column c the entries (i, c), i ascending (rows are visited in order)
    tcols = [[] for _ in range(n)]
    for i in range(n):
        for t in range(int(ui[i]), int(ui[i + 1]) - 1):
            tcols[int(uj[t])].append((i, t))
    y = np.zeros(n)       # unscaled forward values
    x = np.zeros(n)
    for c in range(n):
        s = b[c]
        for i, t in tcols[c]:
            s = s + final[t] * y[i]
        y[c] = s
        x[c] = s * final[udiag[c]]
    for i in range(n - 2, -1, -1):
        s = x[i]
        for t in range(int(ui[i + 1]) - 2, int(ui[i]) - 1, -1):
            s = s + final[t] * x[int(uj[t])]
        x[i] = s
    return x


def levels(ptr, rows):
    n = len(ptr) - 1
    lev = np.zeros(n, np.int64)
    for k in range(n):
        r = rows[ptr[k]:ptr[k + 1]]
        if len(r):
            lev[k] = lev[r].max() + 1
    return lev


def numeric_rowwise(ai, aj, aa, ui, uj, udiag, schedule):
    """Row-parallel formulation: returns (orig, final) where final is the array the reference's solve reads
    (off-diagonal slots -U(i,c)/D(i), diagonal slot 1/D(k)) and orig holds the unscaled U(k,c)."""
    ptr, rows, pos = schedule
    n = len(ui) - 1
    nz = int(ui[n])
    orig = np.zeros(nz)
    final = np.zeros(nz)
    dinv = np.zeros(n)
    for k in range(n):                       # any order that respects levels(...) gives the same result
        cols = uj[ui[k]:ui[k + 1]]           # pattern of row k (diagonal last)
        where = {int(c): t for t, c in enumerate(cols)}
        rtmp = np.zeros(len(cols))
        for p in range(ai[k], ai[k + 1]):    # upper triangle of A's row
            if aj[p] >= k:
                rtmp[where[int(aj[p])]] = aa[p]
        dk = rtmp[where[k]]
        for q in range(ptr[k], ptr[k + 1]):  # contributors, reference order
            i, e = int(rows[q]), int(pos[q])
            u = orig[e]                       # U(i,k) as row i left it
            uikdi = -u * dinv[i]
            dk = dk + uikdi * u
            final[e] = uikdi                  # the slot's final content (used by the solves)
            for t in range(e + 1, int(ui[i + 1]) - 1):   # later off-diagonal entries of row i (the reference's loop also
                c = int(uj[t])                            # touches the diagonal slot and non-pattern columns: dead stores)
                if c in where:
                    rtmp[where[c]] = rtmp[where[c]] + uikdi * orig[t]
        for t, c in enumerate(cols[:-1]):
            orig[ui[k] + t] = rtmp[t]
            final[ui[k] + t] = rtmp[t]        # stays as is until the row of column c finalises it
        dinv[k] = 1.0 / dk
        final[udiag[k]] = dinv[k]
    return orig, final
