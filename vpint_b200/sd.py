"""Static-discount interpolators: drop-in ``SD_SMRP`` / ``SD_STMRP``.

Mirrors the public surface of VPint/SD_MRP.py (class names, constructor and
``run`` signatures, side effects); the loop itself runs on the B200 through
libvpint_b200.so.
"""

from __future__ import annotations

import numpy as np

from .base import SMRP, STMRP
from .errors import VPintError  # noqa: F401  (re-exported like the reference module)
from .solve import propagate

SPATIAL_CAP = 10000    # VPint/SD_MRP.py:63-66
TEMPORAL_CAP = 10000   # VPint/SD_MRP.py:296-299


class SD_SMRP(SMRP):
    """2-D value propagation with one discount ``gamma`` on every edge
    (reference: VPint/SD_MRP.py:9-158)."""

    def __init__(self, grid, gamma=0.9, init_strategy='mean', mask=None):
        super().__init__(grid, init_strategy=init_strategy, mask=mask)
        self.gamma = gamma

    def set_gamma(self, gamma):
        self.gamma = gamma

    def run(self, iterations=-1, auto_terminate=True, auto_terminate_threshold=1e-4, track_delta=False,
            confidence=False):
        """Same contract as VPint/SD_MRP.py:51-158.  ``confidence`` is accepted and
        unused, as in the reference.  Non-square grids raise ValueError exactly where
        NumPy does in the reference's neighbour-count setup (:95-96)."""
        if iterations > -1:
            auto_terminate = False
        else:
            iterations = SPATIAL_CAP
        height, width = self.pred_grid.shape[0], self.pred_grid.shape[1]
        if height != width:
            raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) " % (height, width))
        delta_vec = np.zeros(iterations) if track_delta else None
        if iterations > 0:
            new_grid, deltas, sweeps = propagate(
                self.original_grid, self.pred_grid, planes=False, gamma=self.gamma, max_iter=iterations,
                auto_terminate=auto_terminate, threshold=auto_terminate_threshold, track_delta=track_delta)
            self.pred_grid = new_grid
            if track_delta:
                delta_vec[:sweeps] = deltas
                if auto_terminate:
                    delta_vec = delta_vec[0:sweeps]
        self.run_state = True
        if track_delta:
            return self.pred_grid, delta_vec
        return self.pred_grid

    def find_gamma(self, search_epochs, subsample_proportion, sub_iterations=100, ext=None, max_gamma=1, min_gamma=0):
        """Random search for gamma on a sub-sampled copy of the grid
        (reference: VPint/SD_MRP.py:162-214).  The NumPy global RNG is consumed in
        the reference's order -- one ``rand()`` per known cell in row-major order,
        then one ``uniform`` per epoch -- so a seeded caller gets the same gammas.
        All epochs are independent fixed-length solves of the same grid, so they run
        as ONE batched device launch (problem p = epoch p)."""
        from .batch import sd_trials
        if ext is not None:
            # the reference's training-set branch is an empty stub that then fails on an
            # unbound name (VPint/SD_MRP.py:172-174,212)
            raise NotImplementedError("find_gamma(ext=...) is not implemented by the reference")
        sub_grid = self.original_grid.copy()
        known = ~np.isnan(sub_grid)
        draws = np.random.rand(int(known.sum()))
        hide = np.zeros(sub_grid.shape, dtype=bool)
        hide[known] = draws < subsample_proportion
        sub_grid[hide] = np.nan
        gammas = np.array([np.random.uniform(low=min_gamma, high=max_gamma) for _ in range(search_epochs)])
        best_gamma = 0.9
        if search_epochs > 0:
            preds = sd_trials(sub_grid, gammas, sub_iterations)
            best_loss = np.inf
            for ep in range(search_epochs):
                # the reference accumulates |error| cell by cell in row-major order
                mae = np.add.accumulate(np.abs(preds[ep][hide] - self.original_grid[hide]))[-1] / int(hide.sum())
                if mae < best_loss:
                    best_gamma = gammas[ep]
                    best_loss = mae
        self.gamma = best_gamma
        return best_gamma


class SD_STMRP(STMRP):
    """(H, W, T) value propagation with spatial discount ``gamma`` and temporal
    discount ``tau`` (reference: VPint/SD_MRP.py:234-406)."""

    def __init__(self, grid, auto_timesteps=False, gamma=0.9, tau=0.9):
        super(SD_STMRP, self).__init__(grid, auto_timesteps)
        self.gamma = gamma
        self.tau = tau

    def set_gamma(self, gamma):
        self.gamma = gamma

    def set_tau(self, tau):
        self.tau = tau

    def find_discounts(self, search_epochs, subsample_proportion, sub_iterations=100, ext=None):
        """Random search for (gamma, tau) on a sub-sampled copy of the cube (reference:
        VPint/SD_MRP.py:409-476).  The NumPy global RNG is consumed in the reference's order (one
        ``rand()`` per known cell in (i, j, t) order, then ``rand(), rand()`` per epoch).  The reference
        reuses ONE scratch interpolator and calls ``reset()`` between epochs, so every epoch after
        the first starts from the sub-sampled cube with its NaNs; that is reproduced.  All epochs run
        as one batched device solve."""
        from .batch import sd_trials
        if ext is not None:
            raise NotImplementedError("find_discounts(ext=...) is not implemented by the reference")
        sub_grid = self.original_grid.copy()
        known = ~np.isnan(sub_grid)
        draws = np.random.rand(int(known.sum()))
        hide = np.zeros(sub_grid.shape, dtype=bool)
        hide[known] = draws < subsample_proportion
        sub_grid[hide] = np.nan
        pairs = [(np.random.rand(), np.random.rand()) for _ in range(search_epochs)]
        best_gamma, best_tau, best_loss = self.gamma, self.tau, np.inf
        if search_epochs > 0:
            preds = sd_trials(sub_grid, [g for g, _ in pairs], sub_iterations, taus=[t for _, t in pairs],
                              nan_start_after_first=True)
            n_err = int(hide.sum())
            for ep, (gamma, tau) in enumerate(pairs):
                mae = np.add.accumulate(np.abs(preds[ep][hide] - self.original_grid[hide]))[-1] / n_err
                if mae < best_loss:
                    best_gamma, best_tau, best_loss = gamma, tau, mae
        self.gamma = best_gamma
        self.tau = best_tau
        return (best_gamma, best_tau)

    def run(self, iterations=-1, method='predict', auto_terminate=True, auto_terminate_threshold=1e-4,
            track_delta=False):
        """Same contract as VPint/SD_MRP.py:286-406 (``method`` is accepted and unused)."""
        if iterations > -1:
            auto_terminate = False
        else:
            iterations = TEMPORAL_CAP
        delta_vec = np.zeros(iterations) if track_delta else None
        if iterations > 0:
            new_grid, deltas, sweeps = propagate(
                self.original_grid, self.pred_grid, planes=False, gamma=self.gamma, tau=self.tau,
                max_iter=iterations, auto_terminate=auto_terminate, threshold=auto_terminate_threshold,
                track_delta=track_delta)
            self.pred_grid = new_grid
            if track_delta:
                delta_vec[:sweeps] = deltas
                if auto_terminate:
                    delta_vec = delta_vec[0:sweeps]
        if track_delta:
            return self.pred_grid, delta_vec
        return self.pred_grid
