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
