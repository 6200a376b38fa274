"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

Restatement of NumPy's blocked pairwise summation -- the arithmetic behind ``np.nanmean`` in
``MRP.init_pred_grid`` (VPint/MRP.py:73-77) -- in the decomposition the device mean init uses
(vpint_b200/csrc/vpint_ingest.cu: ``leaf_of``, ``fill_leaf_kernel``, ``subtree_sum``, ``fill_top_kernel``).

NumPy is a third-party dependency of the reference (``numpy~=1.26.4``, requirements.txt:2; 2.3 in this image), not
vendored under /root/reference.  Its published algorithm (numpy/_core/src/umath/loops_utils.h.src, ``pairwise_sum``):

* n < 8: plain left-to-right sum;
* n <= 128 (a LEAF): eight running accumulators ``r[j] += a[i + j]``, combined as
  ``((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7))``, then the ``n % 8`` tail left to right;
* n > 128: split at ``n2 = n // 2 - (n // 2) % 8`` and add the two halves.

Pinned by tests/test_host_api.py against NumPy itself, bit for bit, in float32 and float64.
"""
import numpy as np

BLOCK = 128


def leaf_of(n, idx):
    """(start, length) of the leaf that holds element ``idx`` of a range of ``n`` elements -- the table-free walk
    down the tree every warp of ``fill_leaf_kernel`` performs."""
    a, length = 0, n
    while length > BLOCK:
        h = length // 2
        h -= h % 8
        if idx < a + h:
            length = h
        else:
            a += h
            length -= h
    return a, length


def leaf_sum(a):
    """Sum of one leaf (<= 128 elements) in the array's own dtype."""
    dt = a.dtype.type
    n = a.size
    if n < 8:
        res = dt(0)
        for v in a:
            res = dt(res + v)
        return res
    r = [dt(a[j]) for j in range(8)]
    m8 = n - n % 8
    for i in range(8, m8, 8):
        for j in range(8):
            r[j] = dt(r[j] + a[i + j])
    res = dt(dt(dt(r[0] + r[1]) + dt(r[2] + r[3])) + dt(dt(r[4] + r[5]) + dt(r[6] + r[7])))
    for i in range(m8, n):
        res = dt(res + a[i])
    return res


def leaves(n):
    """All leaves of a range of n elements in order, found the way the kernel finds them: chunk c (64 elements) owns
    the leaf that holds element min(64 c + 63, n - 1) if that leaf starts inside the chunk."""
    out = []
    for c in range((n + 63) // 64):
        lo, m = leaf_of(n, min(64 * c + 63, n - 1))
        if lo >= 64 * c:
            out.append((lo, m))
    return out


def tree_sum(leaf_sums, a, length, dt):
    """Adds the leaf sums (indexed by start // 64) in recursion order: ``subtree_sum`` of the device code."""
    if length <= BLOCK:
        return leaf_sums[a >> 6]
    h = length // 2
    h -= h % 8
    return dt(tree_sum(leaf_sums, a, h, dt) + tree_sum(leaf_sums, a + h, length - h, dt))


def pairwise_sum(a):
    """``np.add.reduce`` of a contiguous 1-D array, via leaves + tree."""
    a = np.ascontiguousarray(a).reshape(-1)
    dt = a.dtype.type
    n = a.size
    if n == 0:
        return dt(0)
    sums = {}
    for lo, m in leaves(n):
        sums[lo >> 6] = leaf_sum(a[lo:lo + m])
    return dt(dt(0) + tree_sum(sums, 0, n, dt))


def nanmean(a):
    """``np.nanmean`` of an array in ITS dtype: NaN cells count as 0 in the sum and are left out of the count; the
    quotient is formed as NumPy forms it (true_divide of the dtype's sum by the integer count)."""
    a = np.ascontiguousarray(a).reshape(-1)
    known = ~np.isnan(a)
    total = pairwise_sum(np.where(known, a, a.dtype.type(0)))
    return np.true_divide(total, np.intp(known.sum()), dtype=a.dtype, casting="unsafe")
