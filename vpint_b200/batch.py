"""Batched entry points: many independent problems of one shape in one device solve.

The reference has no batch API (callers loop in Python; VPint2 forks one process
per band, VPint/VPint2.py:63-99).  These helpers are what the drop-in classes use
underneath, and what a caller with many patches / trials should call directly.
"""

from __future__ import annotations

import math

import numpy as np

from . import solve as _solve
from .engine import Solver, require_cuda, to_device
from .errors import VPintError


def sd_trials(grid, gammas, iterations, taus=None, device=None, k_block=None, nan_start_after_first=False):
    """Run SD_SMRP / SD_STMRP (mean init, fixed ``iterations``) on the SAME grid for
    every discount in ``gammas`` (and ``taus`` for cubes) as one batched solve.
    This is the inner loop of find_gamma / find_discounts (VPint/SD_MRP.py:186-195,
    441-452).  ``nan_start_after_first``: trials after the first start from the grid itself,
    NaNs included -- what find_discounts does by calling reset() on its one scratch object
    between epochs (SD_MRP.py:472).  Returns a float64 array (len(gammas),) + grid.shape."""
    dev = require_cuda(device)
    torch = __import__("torch")
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float64))
    n = len(gammas)
    ndim = grid.ndim
    if ndim == 2 and grid.shape[0] != grid.shape[1]:
        raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) " % grid.shape)
    H, W = grid.shape[0], grid.shape[1]
    T = grid.shape[2] if ndim == 3 else 1
    if iterations <= 0:
        pred = grid.copy()
        pred[np.isnan(pred)] = np.nanmean(grid)
        return np.broadcast_to(pred, (n,) + grid.shape).copy()
    solver = Solver(ndim, n, H, W, T, planes=False, k_block=(0 if ndim == 3 else (_solve.DEFAULT_K_BLOCK if k_block is None else k_block)), device=dev)
    try:
        g_dev = to_device(grid, dev)
        original = g_dev.unsqueeze(0).expand((n,) + tuple(g_dev.shape))      # stride-0 problem axis: no copies
        if nan_start_after_first and n > 1:
            first = grid.copy()
            first[np.isnan(first)] = np.nanmean(grid)
            start = torch.empty((n,) + tuple(g_dev.shape), dtype=g_dev.dtype, device=dev)
            start[0] = to_device(first, dev)
            start[1:] = g_dev
            solver.load(original.contiguous(), pred0=start)
        else:
            fill = np.full(n, np.nanmean(grid))
            solver.load(original, fill=fill)
        solver.discounts(np.asarray(gammas, dtype=np.float64), None if taus is None else np.asarray(taus, dtype=np.float64))
        solver.run(int(iterations), auto_terminate=False)
        out, _ = solver.result()
        return out.cpu().numpy()
    finally:
        solver.close()


def wp_trials(sub_grid, features, original, trials, max_sub_iter=-1, device=None):
    """The independent trials of ``WP_SMRP.auto_adapt(search_strategy='random')`` (VPint/WP_MRP.py:505-571) as ONE
    batched solve: the same sub-sampled grid and features, one problem per candidate.

    trials : list of dicts with 'beta' and / or 'epsilon' (+ 'mu'), all with the same keys -- exactly the ``vals`` the
             reference draws per epoch.  Every problem is what ``WP_SMRP(sub_grid, features).run(...)`` computes for
             its candidate (prioritise_identity with priority_intensity = beta, iterations = max_sub_iter; resistance
             with epsilon / mu), so the mean absolute errors -- and with them the parameters picked -- are the
             reference's.  Returns the errors, one per trial (np.nanmean(|pred - original|))."""
    dev = require_cuda(device)
    n = len(trials)
    if n == 0:
        return []
    keys = set(trials[0])
    if any(set(t) != keys for t in trials):
        raise VPintError("wp_trials needs trials with the same parameters")
    use_beta, use_res = 'beta' in keys, 'epsilon' in keys
    if use_res and 'mu' not in keys:
        raise VPintError("wp_trials needs mu with epsilon")
    sub = np.ascontiguousarray(np.asarray(sub_grid, dtype=np.float64))
    feat = np.ascontiguousarray(np.asarray(features, dtype=np.float64))
    H, W = sub.shape
    # iterations: trials with beta run `iterations=max_sub_iter`, resistance-only trials run()'s default (-1)
    iters = max_sub_iter if use_beta else -1
    auto = iters <= -1
    max_iter = 5000 if auto else int(iters)
    if max_iter == 0:
        pred = sub.copy()
        pred[np.isnan(pred)] = np.nanmean(sub)
        err = np.nanmean(np.absolute(pred.reshape(pred.size) - np.asarray(original).reshape(pred.size)))
        return [err] * n
    kb = 1 if use_res else _solve.DEFAULT_K_BLOCK
    solver = Solver(2, n, H, W, planes=True, k_block=kb, device=dev, resistance=use_res)
    try:
        g_dev = to_device(sub, dev)
        f_dev = to_device(feat, dev)
        solver.load(g_dev.unsqueeze(0).expand((n, H, W)), fill=np.full(n, np.nanmean(sub)))
        solver.weights(f_dev.unsqueeze(0).expand((n,) + tuple(f_dev.shape)), method="exact", clamp=False)
        solver.trial_params(beta=[t['beta'] for t in trials] if use_beta else None,
                            epsilon=[t['epsilon'] for t in trials] if use_res else None,
                            mu=[t['mu'] for t in trials] if use_res else None)
        if use_beta:
            solver.priority(0.0)
        solver.run(max_iter, auto_terminate=auto, threshold=1e-4, resistance=(0.0, 0.0) if use_res else None)
        out, _ = solver.result()
        pred = out.cpu().numpy()
    finally:
        solver.close()
    orig = np.asarray(original).reshape(-1)
    return [np.nanmean(np.absolute(pred[i].reshape(-1) - orig)) for i in range(n)]
