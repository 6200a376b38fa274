"""Host glue between the NumPy-facing classes and the device solver.

One call = one ``run()`` of the reference: upload the start state, build the
weights on the device, iterate to the reference's stopping sweep, download
``pred_grid`` (float64) and, if asked, the per-sweep delta vector.
"""

from __future__ import annotations

import math

import numpy as np

from .engine import Solver, require_cuda, to_device
from .errors import VPintError

# k_block used when the caller does not choose one (0 lets the library decide).
DEFAULT_K_BLOCK = 0
# Several feature layers: the weights are four planes per cell (no ratio form, no cluster-resident solve), the sweep
# streams 48 B per unknown cell and a single grid is launch-latency bound at one sweep per launch -- the TMA-staged
# kernel with 4 sweeps per launch is the better default there (BASELINE config 2, 1024 x 1024, F = 3: 0.38 ms against
# 0.67 ms per run()).
MULTI_FEATURE_K_BLOCK = 4


def _as_float(a):
    a = np.asarray(a)
    if a.dtype == np.float32 or a.dtype == np.float64:
        return a
    return a.astype(np.float64)


def propagate(original, pred, *, planes, features=None, method="exact", clamp=False, min_gamma=0.0,
              max_gamma=math.inf, gamma=None, tau=None, max_iter, auto_terminate, threshold, clip_val=math.inf,
              track_delta=False, storage="float64", k_block=None, device=None,
              priority=None, resistance=None, info=None):
    """Run the propagation loop for ONE grid (2-D) or cube (3-D).

    original / pred : numpy (H, W) or (H, W, T); NaN in ``original`` = unknown.
    features        : numpy (H, W, F) when ``planes``.
    priority        : None or dict(beta=..., bias=(H, W) ndarray or None, want_planes=bool): prioritise_identity
                      (VPint/WP_MRP.py:243-276); the planes come back as ``info["priority_planes"]``.
    info            : optional dict that receives by-products of the call (no process-global state: the library
                      is driven from several host threads).
    resistance      : None or (epsilon, mu) (VPint/WP_MRP.py:327-348).
    Returns (pred_grid float64 ndarray, delta_vec float64 ndarray or None, sweeps).
    """
    dev = require_cuda(device)
    original = _as_float(original)
    pred = _as_float(pred)
    if original.shape != pred.shape:
        raise VPintError("pred_grid and original_grid have different shapes")
    if original.dtype != pred.dtype:
        original = original.astype(np.float64)
        pred = pred.astype(np.float64)
    ndim = original.ndim
    if ndim not in (2, 3):
        raise VPintError("expected a 2-D grid or a 3-D cube, got shape " + str(original.shape))
    H, W = original.shape[0], original.shape[1]
    T = original.shape[2] if ndim == 3 else 1
    kb = DEFAULT_K_BLOCK if k_block is None else k_block
    if ndim == 3:
        kb = 0
    if (priority is not None or resistance is not None) and (ndim != 2 or not planes):
        raise VPintError("prioritise_identity / resistance belong to WP_SMRP (2-D, weight planes)")
    if (kb == 0 and ndim == 2 and planes and resistance is None and features is not None
            and np.ndim(features) == 3 and np.shape(features)[2] > 1):
        kb = MULTI_FEATURE_K_BLOCK
    if resistance is not None and kb not in (0, 1):
        kb = 1          # the resistance sweep rewrites every cell: one sweep per launch
    solver = Solver(ndim, 1, H, W, T, storage=storage, planes=planes, k_block=kb, device=dev,
                    resistance=resistance is not None)
    try:
        o_dev = to_device(np.ascontiguousarray(original), dev).unsqueeze(0)
        p_dev = to_device(np.ascontiguousarray(pred), dev).unsqueeze(0)
        solver.load(o_dev, pred0=p_dev)
        if planes:
            f = _as_float(features)
            f_dev = to_device(np.ascontiguousarray(f), dev).unsqueeze(0)
            solver.weights(f_dev, method=method, clamp=clamp, min_gamma=min_gamma, max_gamma=max_gamma)
            if priority is not None:
                bias = priority.get("bias")
                b_dev = None if bias is None else to_device(np.ascontiguousarray(bias, dtype=np.float64), dev).unsqueeze(0)
                planes_out = solver.priority(priority["beta"], bias=b_dev, want_planes=priority.get("want_planes", False))
                if planes_out is not None and info is not None:
                    info["priority_planes"] = planes_out[0].cpu().numpy()
        else:
            solver.discounts(gamma, tau if ndim == 3 else None)
        solver.run(max_iter, auto_terminate=auto_terminate, threshold=threshold, clip_val=clip_val,
                   track_delta=track_delta, resistance=resistance)
        out, niter = solver.result()
        sweeps = int(niter.cpu().numpy()[0])
        result = out[0].cpu().numpy()
        delta_vec = None
        if track_delta:
            if solver.delta is None:
                delta_vec = np.zeros(0)
            else:
                delta_vec = solver.delta[0, :sweeps].cpu().numpy()
        return result, delta_vec, sweeps
    finally:
        solver.close()

