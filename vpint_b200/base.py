"""Grid storage and initial-state semantics shared by the interpolators.

Host-side mirror of the reference base classes (VPint/MRP.py): ``MRP`` (:8-113),
``SMRP`` (:262-282), ``STMRP`` (:328-438).  This stays NumPy on the host: its
output (``original_grid`` with NaN for unknown cells, ``pred_grid`` as the start
state) is exactly what the CUDA path is fed, so parity of the initial state is
by construction.
"""

from __future__ import annotations

import datetime

import numpy as np

from .errors import VPintError


class MRP:
    """Common state of all interpolators (reference: VPint/MRP.py:8-113).

    Attributes: ``original_grid`` (copy of the input, NaN = unknown),
    ``pred_grid`` (current state; ``run()`` starts from it and stores its result
    there, so ``run(a); run(b)`` continues), ``dims``.
    """

    def __init__(self, grid, init_strategy='mean', mask=None):
        if mask is not None:
            grid = self.apply_mask(grid, mask)
        self.original_grid = grid.copy()
        self.dims = len(grid.shape)
        self.init_pred_grid(init_strategy=init_strategy)

    def __str__(self):
        return str(self.pred_grid)

    def reset(self):
        """pred_grid becomes original_grid itself, NaNs included (VPint/MRP.py:59-61)."""
        self.pred_grid = self.original_grid

    def init_pred_grid(self, init_strategy='mean'):
        """Fill unknown cells of a copy of original_grid (VPint/MRP.py:64-91):
        'mean' -> nanmean in the grid's own dtype, 'zero' -> 0, 'random' ->
        N(nanmean, nanstd) draws from the global NumPy RNG."""
        pred = self.original_grid.copy()
        flat = pred.reshape(pred.size)
        unknown = np.isnan(flat)
        if init_strategy == 'mean':
            flat[unknown] = np.nanmean(pred)
        elif init_strategy == 'zero':
            flat[unknown] = 0
        elif init_strategy == 'random':
            flat[unknown] = np.random.normal(np.nanmean(flat), np.nanstd(flat), size=int(unknown.sum()))
        else:
            raise VPintError("Invalid initialisation strategy: " + str(init_strategy))
        self.pred_grid = flat.reshape(pred.shape)

    def get_pred_grid(self):
        return self.pred_grid

    def apply_mask(self, grid, mask):
        """Reference behaviour kept as is (VPint/MRP.py:100-113): the "mask vector" is
        built from ``grid`` itself, so cells whose VALUE equals 1 become NaN, and the
        caller's array is modified through the reshape view.  ``mask`` only has to
        have the right shape.  Not "fixed" here: a drop-in must not change results."""
        shp = grid.shape
        grid_vec = grid.reshape(grid.size)
        mask_vec = grid.reshape(mask.size)
        if grid.shape != mask.shape:
            raise VPintError("Target and mask grids have different shapes: " + str(grid.shape) + " and " + str(mask.shape))
        grid_vec[mask_vec == 1] = np.nan
        return grid_vec.reshape(shp)

    # -- evaluation helpers on unknown cells (VPint/MRP.py:116-259) -------------------
    # Host NumPy like the reference (not on the device path); the formulas are the reference's,
    # its per-cell loops written as array expressions: NaNs of the ground truth are first replaced by
    # its nanmean, the errors are taken over the cells that were unknown in original_grid.
    def _eval_pairs(self, true_grid):
        true = np.asarray(true_grid)
        true = np.nan_to_num(true, nan=np.nanmean(true))
        sel = np.isnan(self.original_grid)
        return true, true[sel], np.asarray(self.pred_grid)[sel]

    def r_squared(self, true_grid):
        """1 - SS_res / SS_tot over the unknown cells, SS_tot against the mean of the WHOLE ground truth
        (VPint/MRP.py:123-157)."""
        true, t, p = self._eval_pairs(true_grid)
        m = np.mean(true)
        res = np.sum((t - p) ** 2)
        tot = np.sum((t - m) ** 2)
        return 1 - (res / tot)

    def MAE(self, true):
        """VPint/MRP.py:160-180 (nanmean of |true - pred| over the unknown cells)."""
        _, t, p = self._eval_pairs(true)
        return np.nanmean(np.absolute(t - p))

    def RMSE(self, true):
        """The reference returns the MEAN SQUARED error under this name (VPint/MRP.py:183-204); kept."""
        _, t, p = self._eval_pairs(true)
        return np.mean(np.square(t - p))

    def PSNR(self, true):
        """VPint/MRP.py:207-233: 20 log10(255 / sqrt(mse + 0.001)) / 100."""
        from math import log10, sqrt
        _, t, p = self._eval_pairs(true)
        mse = np.nanmean((t - p) ** 2) + 0.001
        if mse == 0:
            return 1
        return 20 * log10(255.0 / sqrt(mse)) / 100

    def SSIM(self, true):
        """VPint/MRP.py:235-259 needs scikit-image's ``compare_ssim`` and returns NaN without it; the same here."""
        _, t, p = self._eval_pairs(true)
        try:
            from skimage.measure import compare_ssim
            (s, d) = compare_ssim(t, p, full=True)
        except Exception:
            print("Error running compare_ssim. For this functionality, please ensure that scikit-image is installed.")
            s = np.nan
        return s


class SMRP(MRP):
    """2-D grids (VPint/MRP.py:262-282)."""

    def __init__(self, grid, init_strategy='mean', mask=None):
        super().__init__(grid, init_strategy=init_strategy, mask=mask)

    def mean_absolute_error(self, true_grid, scale_factor=1, gridded=False):
        """Block-summed MAE over the unknown cells ("old code" of the reference, VPint/MRP.py:284-325; the
        ground truth is only block-summed along the rows there -- kept)."""
        height, width = self.pred_grid.shape[0], self.pred_grid.shape[1]
        nh, nw = int(height / scale_factor), int(width / scale_factor)
        blk = (nh, int(height / nh), nw, int(width / nw))
        pred = self.pred_grid.copy().reshape(blk).sum(3).sum(1)
        orig = self.original_grid.copy().reshape(blk).sum(3).sum(1)
        # (nh, width): summed over row blocks only; the reference then indexes its first nw columns
        true = true_grid.copy().reshape((nh, int(height / nh), width, int(nw / nw))).sum(3).sum(1)[:, :nw]
        sel = np.isnan(orig)
        error_grid = np.zeros((nh, nw))
        error_grid[sel] = np.abs(pred - true)[sel]
        mae = error_grid[sel].sum() / int(sel.sum())
        return (mae, error_grid) if gridded else mae


class STMRP(MRP):
    """(H, W, T) cubes (VPint/MRP.py:328-438)."""

    def __init__(self, data, auto_timesteps, init_strategy='mean'):
        if auto_timesteps:
            new_grid = self.set_timesteps(data.copy())
        else:
            new_grid = self.dim_check(data.copy())
        super().__init__(new_grid, init_strategy=init_strategy)

    def dim_check(self, grid):
        """1-D / 2-D inputs become single-column / single-time-step cubes (VPint/MRP.py:355-372)."""
        if grid.ndim == 1:
            out = np.zeros((grid.shape[0], 1, 1))
            out[:, 0, 0] = grid
            return out
        if grid.ndim == 2:
            out = np.zeros((grid.shape[0], grid.shape[1], 1))
            out[:, :, 0] = grid
            return out
        return grid

    def set_timesteps(self, data):
        """dict {"YYYY-MM-DD HH:MM:SS": 2-D grid} (chronological) -> NaN-padded cube with one
        slice per smallest gap between consecutive stamps (VPint/MRP.py:384-438)."""
        fmt = "%Y-%m-%d %H:%M:%S"
        stamps = [datetime.datetime.strptime(k, fmt) for k in data.keys()]
        gaps = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
        min_gap = min(gaps)
        first, last = stamps[0], stamps[-1]
        sample = data[first.strftime(fmt)]
        depth = int((last - first) / min_gap) + 1
        cube = np.empty((sample.shape[0], sample.shape[1], depth))
        cube[:] = np.nan
        self.true_time_indices = []
        for k, v in data.items():
            ind = int((datetime.datetime.strptime(k, fmt) - first) / min_gap)
            cube[:, :, ind] = v
            self.true_time_indices.append(ind)
        return cube
