"""CPU oracle for the VPint value-propagation hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference algorithm (LaurensArp/VPint
0.2.4, pure Python + NumPy) for the path behind ``SD_SMRP.run``,
``WP_SMRP.run``, ``SD_STMRP.run``, ``WP_STMRP.run`` and
``VPint2_interpolator.run / run_serial``.  It exists to CHECK the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``vpint_b200/`` imports it and the product has no CPU fallback.

Pinning: the reference publishes no golden vectors (its 8 tests only assert
shape / no-NaN on unseeded inputs, ``test/test_SD_SMRP.py:23-50``,
``test/test_WP_SMRP.py:30-45``), so this oracle is pinned against outputs of
the reference itself, generated in the build container by importing
``/root/reference`` (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``)
and compared bit-for-bit in ``tests/test_oracle_golden.py``.

The arithmetic engine of the reference is NumPy (``numpy~=1.26.4`` pinned in
``requirements.txt:2``; not vendored).  The restatement issues the same NumPy
primitives on arrays with the same memory layout (``einsum('ij,ji->i')`` on a
C-contiguous value matrix and a transposed weight view, ``nanmean`` for the
stopping rule), so it reproduces the reference bit-for-bit, not just to 1e-9.

Every function cites the reference lines (relative to ``/root/reference``) it
follows.
"""

from __future__ import annotations

import numpy as np

# Iteration caps when ``iterations`` is left at -1.
WP_SPATIAL_CAP = 5000      # VPint/WP_MRP.py:112-115
SD_SPATIAL_CAP = 10000     # VPint/SD_MRP.py:63-66
TEMPORAL_CAP = 10000       # VPint/SD_MRP.py:296-299, VPint/WP_MRP.py:1300-1303


class OracleError(Exception):
    """Mirrors VPintError (VPint/MRP.py:489-490) for invalid arguments."""


# --------------------------------------------------------------------------
# A1  initial state                                      VPint/MRP.py:47-52,64-91
# --------------------------------------------------------------------------
def init_pred(original, init_strategy="mean"):
    """Copy of ``original`` with NaN cells filled (MRP.init_pred_grid).

    'mean' fills with ``nanmean`` taken in the INPUT dtype (float32 grids get a
    float32 mean, which is what VPint2 feeds), 'zero' with 0, 'random' with
    draws from N(nanmean, nanstd) using the global NumPy RNG.
    """
    pred = np.array(original, copy=True)
    flat = pred.reshape(pred.size)
    holes = np.isnan(flat)
    if init_strategy == "mean":
        flat[holes] = np.nanmean(pred)
    elif init_strategy == "zero":
        flat[holes] = 0
    elif init_strategy == "random":
        flat[holes] = np.random.normal(np.nanmean(flat), np.nanstd(flat), size=int(holes.sum()))
    else:
        raise OracleError("Invalid initialisation strategy: " + str(init_strategy))
    return flat.reshape(pred.shape)


# --------------------------------------------------------------------------
# A3  neighbour counts          VPint/WP_MRP.py:201-209, VPint/SD_MRP.py:327-343
# --------------------------------------------------------------------------
def neighbour_count(shape):
    """Number of in-grid axis neighbours of every cell (2-D: 4/3/2, 3-D: 6..3).

    Each axis subtracts one at index 0 and one at index n-1, so a length-1 axis
    subtracts two, exactly as the reference's successive row/column updates do.
    """
    nc = np.ones(shape) * (2 * len(shape))
    for axis, n in enumerate(shape):
        edge = np.zeros(n)
        edge[0] += 1
        edge[n - 1] += 1
        view = [1] * len(shape)
        view[axis] = n
        nc = nc - edge.reshape(view)
    return nc


# --------------------------------------------------------------------------
# A2  weights                                           VPint/WP_MRP.py:138-195
# --------------------------------------------------------------------------
def weights_exact_2d(feature_grid):
    """'exact' weights of WP_SMRP.run (VPint/WP_MRP.py:138-170).

    ``feature_grid`` is (H, W, F) float64 and is MUTATED: zeros become 0.01
    (:143-145).  Returns (H, W, 4) ordered up, right, down, left where
    w_dir[i, j] = mean_k f[i, j, k] / f[neighbour_dir, k] and the rows/columns
    without that neighbour are 0 (:161-168).
    """
    f = feature_grid
    flat = f.reshape(f.size)
    flat[flat == 0] = 0.01
    H, W, F = f.shape
    nb = [np.ones((H, W, F)) for _ in range(4)]
    nb[0][1:, :, :] = f[0:-1, :, :]     # neighbour above
    nb[1][:, 0:-1, :] = f[:, 1:, :]     # neighbour to the right
    nb[2][0:-1, :, :] = f[1:, :, :]     # neighbour below
    nb[3][:, 1:, :] = f[:, 0:-1, :]     # neighbour to the left
    planes = [np.mean(f / g, axis=2) for g in nb]
    planes[0][0, :] = 0
    planes[1][:, -1] = 0
    planes[2][-1, :] = 0
    planes[3][:, 0] = 0
    return np.stack(planes, axis=-1)


def pair_weight(f_nb, f_own, method, min_gamma=0, max_gamma=np.inf, clamp=True):
    """One weight from neighbour features to own features.

    WP_SMRP.predict_weight (VPint/WP_MRP.py:396-427; clamps to
    [min_gamma, max_gamma]) and WP_STMRP.predict_weight
    (VPint/WP_MRP.py:1443-1468; no clamp for these methods -> clamp=False).
    """
    if method == "cosine_similarity":
        g = np.dot(f_nb, f_own) / max(np.sum(f_nb) * np.sum(f_own), 0.01)
    elif method == "exact":
        d = f_nb.copy()
        d[d == 0] = 0.01
        g = np.mean(f_own / d)
    elif method == "exact_inverse":
        d = f_own.copy()
        d[d == 0] = 0.01
        g = np.mean(f_nb / d)
    else:
        raise OracleError("Invalid weight prediction method: " + str(method))
    if clamp:
        g = max(min_gamma, min(g, max_gamma))
    return g


def weights_pairwise_2d(feature_grid, method, min_gamma=0, max_gamma=np.inf):
    """Per-pixel weight loop used for every method except the vectorised
    'exact' branch (VPint/WP_MRP.py:173-195).  O(HW) interpreter work: small
    grids only."""
    H, W, _ = feature_grid.shape
    out = np.zeros((H, W, 4))
    for i in range(H):
        for j in range(W):
            own = feature_grid[i, j, :]
            if i > 0:
                out[i, j, 0] = pair_weight(feature_grid[i - 1, j, :], own, method, min_gamma, max_gamma)
            if j < W - 1:
                out[i, j, 1] = pair_weight(feature_grid[i, j + 1, :], own, method, min_gamma, max_gamma)
            if i < H - 1:
                out[i, j, 2] = pair_weight(feature_grid[i + 1, j, :], own, method, min_gamma, max_gamma)
            if j > 0:
                out[i, j, 3] = pair_weight(feature_grid[i, j - 1, :], own, method, min_gamma, max_gamma)
    return out


def weights_3d(feature_grid, depth, method):
    """WP_STMRP weights (VPint/WP_MRP.py:1319-1355) as (H, W, T, 6).

    Features are (H, W, F) and static in time (:1323 indexes [i, j, :]).
    Slots 0-3 are up/right/down/left; slot 4 is set where t > 0 and slot 5
    where t < T-1, both from the cell's own features (:1344-1353).
    """
    H, W, _ = feature_grid.shape
    spatial = np.zeros((H, W, 4))
    temporal = np.zeros((H, W))
    for i in range(H):
        for j in range(W):
            own = feature_grid[i, j, :]
            if i > 0:
                spatial[i, j, 0] = pair_weight(feature_grid[i - 1, j, :], own, method, clamp=False)
            if j < W - 1:
                spatial[i, j, 1] = pair_weight(feature_grid[i, j + 1, :], own, method, clamp=False)
            if i < H - 1:
                spatial[i, j, 2] = pair_weight(feature_grid[i + 1, j, :], own, method, clamp=False)
            if j > 0:
                spatial[i, j, 3] = pair_weight(feature_grid[i, j - 1, :], own, method, clamp=False)
            temporal[i, j] = pair_weight(own, own, method, clamp=False)
    w = np.zeros((H, W, depth, 6))
    w[:, :, :, 0:4] = spatial[:, :, None, :]
    if depth > 1:
        w[:, :, 1:, 4] = temporal[:, :, None]
        w[:, :, :-1, 5] = temporal[:, :, None]
    return w


# --------------------------------------------------------------------------
# A6  identity priority                                 VPint/WP_MRP.py:243-276
# --------------------------------------------------------------------------
def priority_planes(weight_grid, priority_intensity, known_value_bias=0, original=None):
    """Static per-direction priority weights.

    beta == 0: ones with the missing-neighbour entries zeroed (:247-255).
    Otherwise two SEQUENTIAL masked rewrites of a copy of the weights
    (:257-259): entries > 1 become (1/p) * ((1/p) / ((1/p) * beta)); then every
    entry < 1 -- including ones just rewritten -- becomes
    p * (p / ((p + 0.001) * beta)).  ``known_value_bias`` > 0 multiplies by an
    SD_SMRP solve of the known-cell indicator (:263-276).
    """
    if priority_intensity == 0:
        p = np.ones(weight_grid.shape)
        p[0, :, 0] = 0.0
        p[:, -1, 1] = 0.0
        p[-1, :, 2] = 0.0
        p[:, 0, 3] = 0.0
    else:
        p = weight_grid.copy()
        big = p > 1
        p[big] = 1 / p[big] * (1 / p[big] / (1 / p[big] * priority_intensity))
        small = p < 1
        p[small] = p[small] * (p[small] / ((p[small] + 0.001) * priority_intensity))
    if known_value_bias > 0:
        ind = np.zeros(p.shape)
        for d in range(p.shape[2]):
            ind[:, :, d] = original.copy()
        ind[~np.isnan(ind)] = 1
        solved = ind.copy()
        for d in range(p.shape[2]):
            sub_pred = init_pred(ind[:, :, d], "zero")
            solved[:, :, d], _, _ = sd_run_2d(ind[:, :, d], sub_pred, gamma=(1 - known_value_bias))
        p = np.multiply(p, solved)
        flat = p.reshape(p.size)
        flat[flat == 0] = 0.01
        p = flat.reshape(p.shape)
    return p


# --------------------------------------------------------------------------
# A4  one Jacobi sweep
# --------------------------------------------------------------------------
def _shifted_values(pred, val):
    """Fill val[..., d] with the neighbour of every cell in direction d, 0
    outside the grid (VPint/WP_MRP.py:291-304, VPint/SD_MRP.py:112-126,
    357-378).  2-D order: up, right, down, left; 3-D adds t+1 then t-1."""
    nd = pred.ndim
    val[1:, ..., 0] = pred[:-1]
    val[0, ..., 0] = 0
    if nd == 2:
        val[:, :-1, 1] = pred[:, 1:]
        val[:, -1, 1] = 0
        val[:-1, :, 2] = pred[1:, :]
        val[-1, :, 2] = 0
        val[:, 1:, 3] = pred[:, :-1]
        val[:, 0, 3] = 0
    else:
        val[:, :-1, :, 1] = pred[:, 1:, :]
        val[:, -1, :, 1] = 0
        val[:-1, :, :, 2] = pred[1:, :, :]
        val[-1, :, :, 2] = 0
        val[:, 1:, :, 3] = pred[:, :-1, :]
        val[:, 0, :, 3] = 0
        val[:, :, :-1, 4] = pred[:, :, 1:]
        val[:, :, -1, 4] = 0
        val[:, :, 1:, 5] = pred[:, :, :-1]
        val[:, :, 0, 5] = 0


def _restore_known(new_flat, original):
    """Known cells keep their original value (VPint/SD_MRP.py:133-134,
    VPint/WP_MRP.py:321,323)."""
    flat_orig = original.reshape(original.size)
    known = ~np.isnan(flat_orig)
    new_flat[known] = flat_orig[known]


def relative_change(new, old):
    """Stopping statistic: nanmean(|new - old|) / nanmean(old) over the WHOLE
    grid (VPint/SD_MRP.py:139, VPint/WP_MRP.py:352)."""
    return np.nanmean(np.absolute(new - old)) / np.nanmean(old)


def _loop(original, pred, weight_grid, cap, iterations, auto_terminate, threshold,
          clip_val=np.inf, priority=None, resistance=None):
    """Shared Jacobi driver (VPint/SD_MRP.py:108-149, VPint/WP_MRP.py:287-367,
    VPint/SD_MRP.py:353-401, VPint/WP_MRP.py:1387-1435).

    Returns (pred, delta_vec, n_sweeps).  ``delta_vec`` always has one entry per
    executed sweep (the reference only keeps it under track_delta; recording it
    is free and gives the tests the iteration count).
    """
    if iterations > -1:
        auto_terminate = False
    else:
        iterations = cap
    nslots = weight_grid.shape[-1]
    ncell = pred.size
    nc_vec = neighbour_count(pred.shape).reshape(ncell)
    w_t = weight_grid.reshape((ncell, nslots)).transpose()
    val = np.zeros(pred.shape + (nslots,))
    deltas = []
    for _ in range(iterations):
        _shifted_values(pred, val)
        val_m = val.reshape((ncell, nslots))
        if priority is not None:
            each = np.multiply(w_t.transpose(), val_m)
            new = np.einsum('ij,ji->i', each, priority.reshape(ncell, nslots).transpose())
            new = new / np.sum(priority, axis=2).reshape(ncell)
        else:
            new = np.einsum('ij,ji->i', val_m, w_t)
            new = new / nc_vec
        new[new > clip_val] = clip_val          # VPint/WP_MRP.py:322, before the restore
        _restore_known(new, original)
        new = new.reshape(pred.shape)
        if resistance is not None:              # VPint/WP_MRP.py:327-348
            epsilon, mu = resistance
            nv = new.reshape(ncell)
            ov = pred.reshape(ncell)
            dy = nv - ov
            low = np.where(nv <= mu)
            nv[low] = ov[low] + dy[low]
            high = np.where(nv > mu)
            nv[high] = ov[high] + dy[high] - (epsilon * ov[high])
        delta = relative_change(new, pred)
        deltas.append(delta)
        pred = new
        if auto_terminate and delta <= threshold:
            break
    return pred, np.array(deltas, dtype=np.float64), len(deltas)


# --------------------------------------------------------------------------
# run() restatements
# --------------------------------------------------------------------------
def sd_run_2d(original, pred, gamma=0.9, iterations=-1, auto_terminate=True, threshold=1e-4):
    """SD_SMRP.run (VPint/SD_MRP.py:51-158): all four weights equal gamma.

    The reference only works on square grids: its column update subtracts a
    length-W vector from a length-H column (:95-96) and NumPy raises ValueError.
    """
    H, W = pred.shape
    if H != W:
        raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) " % (H, W))
    w = np.ones((H, W, 4)) * gamma
    return _loop(original, pred, w, SD_SPATIAL_CAP, iterations, auto_terminate, threshold)


def wp_run_2d(original, pred, feature_grid, iterations=-1, method="exact", auto_terminate=True,
              threshold=1e-4, clip_val=np.inf, prioritise_identity=False, priority_intensity=1,
              known_value_bias=0, resistance=False, epsilon=0.01, mu=1.0,
              min_gamma=0, max_gamma=np.inf):
    """WP_SMRP.run (VPint/WP_MRP.py:82-393) without the host-side search
    (auto_adapt) and GIF recording.  ``feature_grid`` is (H, W, F) float64 and is
    mutated by the 'exact' branch like the reference's ``self.feature_grid``."""
    if method == "exact":
        w = weights_exact_2d(feature_grid)
    else:
        w = weights_pairwise_2d(feature_grid, method, min_gamma, max_gamma)
    prio = None
    if prioritise_identity:
        prio = priority_planes(w, priority_intensity, known_value_bias, original)
    res = (epsilon, mu) if resistance else None
    return _loop(original, pred, w, WP_SPATIAL_CAP, iterations, auto_terminate, threshold,
                 clip_val=clip_val, priority=prio, resistance=res)


def sd_run_3d(original, pred, gamma=0.9, tau=0.9, iterations=-1, auto_terminate=True, threshold=1e-4):
    """SD_STMRP.run (VPint/SD_MRP.py:286-406): gamma on the four spatial slots,
    tau on the two temporal ones (:315-321), on an (H, W, T) cube."""
    w = np.ones(pred.shape + (6,))
    w[..., 0:4] = w[..., 0:4] * gamma
    w[..., 4:6] = w[..., 4:6] * tau
    return _loop(original, pred, w, TEMPORAL_CAP, iterations, auto_terminate, threshold)


def wp_run_3d(original, pred, feature_grid, iterations=-1, method="exact", auto_terminate=True,
              threshold=1e-4):
    """WP_STMRP.run (VPint/WP_MRP.py:1289-1440) for the feature-derived methods.

    Keeps the reference's slot pairing: weight slot 4 is nonzero for t > 0 but
    multiplies the t+1 value, slot 5 is nonzero for t < T-1 but multiplies the
    t-1 value (:1344-1353 vs :1406-1412)."""
    w = weights_3d(feature_grid, pred.shape[2], method)
    return _loop(original, pred, w, TEMPORAL_CAP, iterations, auto_terminate, threshold)


def vpint2_run(target, features, **kwargs):
    """VPint2_interpolator.run (VPint/VPint2.py:58-106): every band b is an
    independent WP_SMRP(target[:, :, b], features[:, :, b]).run(**kwargs); the
    result is float64 (H, W, B).  ``target``/``features`` are the interpolator's
    stored arrays (float32 by default, :18-24).  Returns (pred, sweeps_per_band)."""
    H, W, B = target.shape
    out = np.empty((B, H, W), dtype=np.float64)
    sweeps = np.zeros(B, dtype=np.int64)
    for b in range(B):
        grid = target[:, :, b]
        original = grid.copy()
        pred = init_pred(original, "mean")
        feat = features[:, :, b].copy().astype(float).reshape((H, W, 1))
        res, _, n = wp_run_2d(original, pred, feat, **kwargs)
        out[b] = res
        sweeps[b] = n
    return out.swapaxes(0, 2).swapaxes(0, 1), sweeps
