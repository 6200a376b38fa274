"""Device-side driver: owns the workspace (a torch CUDA byte tensor), marshals
tensors to the C ABI and returns torch tensors.  torch is plumbing only (device
memory, streams); every kernel lives in libvpint_b200.so.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native as N
from .errors import VPintError

_DTYPE_CODE = {"float64": N.F64, "float32": N.F32}

# Loop driver of the tiled path when the caller does not choose one: N.LOOP_CHUNKED (host enqueues chunks of
# launches and polls one pinned counter per chunk) or N.LOOP_GRAPH (one CUDA-graph launch with a device-side
# WHILE node; no host involvement until every problem has stopped).
import os as _os
DEFAULT_LOOP_MODE = N.LOOP_GRAPH if _os.environ.get("VPINT_LOOP_MODE", "").lower() == "graph" else N.LOOP_CHUNKED


def _torch():
    import torch
    return torch


def require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise VPintError("vpint_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _code(t):
    torch = _torch()
    if t.dtype == torch.float64:
        return N.F64
    if t.dtype == torch.float32:
        return N.F32
    raise VPintError("unsupported tensor dtype %s (float32 / float64 only)" % t.dtype)


def _raster_code(t):
    torch = _torch()
    code = {torch.float64: N.F64, torch.float32: N.F32, torch.uint16: N.U16, torch.uint8: N.U8}.get(t.dtype)
    if code is None:
        raise VPintError("unsupported raster dtype %s (uint16 / float32 / float64; uint8 for masks)" % t.dtype)
    return code


def _out_code(t):
    """dtype code of a result tensor: float64 / float32, or uint16 for the write-back of a product."""
    torch = _torch()
    if t.dtype == torch.uint16:
        return N.U16
    return _code(t)


def _clip_pair(clip):
    """clip_target / clip_features of the VPint2 constructor (VPint/VPint2.py:30-41): a number is an upper bound,
    anything else is (min, max).  -> (on, lo, hi)."""
    if clip is None:
        return 0, 0.0, 0.0
    if type(clip) == type(2) or type(clip) == type(2.1):
        return 1, -math.inf, float(clip)
    return 1, float(clip[0]), float(clip[1])


def cloud_mask(mask, threshold, out=None):
    """``mask > threshold`` as a uint8 CUDA tensor of the same shape (VPint/VPint2.py:122-132).  mask: CUDA tensor,
    float64 / float32 / uint16 / uint8."""
    torch = _torch()
    lib = N.load_library()
    mask = mask.contiguous()
    if out is None:
        out = torch.empty(mask.shape, dtype=torch.uint8, device=mask.device)
    st = C.c_void_p(torch.cuda.current_stream(mask.device).cuda_stream)
    N.check(lib, lib.vpint_cloud_mask(C.c_void_p(mask.data_ptr()), _raster_code(mask), float(threshold), int(mask.numel()),
                                      C.c_void_p(out.data_ptr()), st), "vpint_cloud_mask")
    return out


def buffer_clouds(target, cloud=None, buffer_size=5, passes=1, scratch=None, out=None):
    """Cloud buffering of the VPint2 constructor on the device (VPint/VPint2.py:135-150, column bound of :146 kept).
    target: CUDA (N, H, W, B) raster (its NaNs join the footprint); cloud: uint8 CUDA (N, H, W) all-band mask, rewritten
    in place, or None -- then the footprint is the target's NaNs alone and the mask is written to ``out``.
    Returns the dilated all-band mask."""
    torch = _torch()
    lib = N.load_library()
    n, H, W, B = target.shape
    have = cloud is not None
    if not have:
        cloud = out if out is not None else torch.empty((n, H, W), dtype=torch.uint8, device=target.device)
    if scratch is None:
        scratch = torch.empty(3 * n * H * W, dtype=torch.uint8, device=target.device)
    st = C.c_void_p(torch.cuda.current_stream(target.device).cuda_stream)
    N.check(lib, lib.vpint_buffer_clouds(C.c_void_p(target.data_ptr()), _raster_code(target), C.c_void_p(cloud.data_ptr()),
                                         int(have), int(n), int(H), int(W), int(B), int(buffer_size), int(passes),
                                         C.c_void_p(scratch.data_ptr()), st), "vpint_buffer_clouds")
    return cloud


def _i64(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def to_device(x, device, dtype=None, non_blocking=False):
    """numpy array / torch tensor -> CUDA tensor (no copy if already there)."""
    torch = _torch()
    if isinstance(x, np.ndarray):
        if not x.flags.writeable:
            x = x.copy()
        x = torch.from_numpy(x)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x.to(device, non_blocking=non_blocking)


class Solver:
    """One batched solve: ``nprob`` independent problems of identical shape.

    Arrays are addressed through a two-level problem index (outer, inner) so
    that e.g. a ``(N, H, W, B)`` VPint2 batch is read in place with
    outer = patch and inner = band.  Tensors handed to :meth:`load`,
    :meth:`weights` and :meth:`result` are viewed as
    ``(outer, inner, H, W[, T | F])``.
    """

    def __init__(self, ndim, nprob, H, W, T=1, *, nprob_inner=1, storage="float64", planes=True,
                 k_block=0, shard=None, device=None, resistance=False, blocking_wait=False):
        self.lib = N.load_library()
        self.torch = _torch()
        self.device = require_cuda(device)
        self.ndim, self.nprob, self.nprob_inner = int(ndim), int(nprob), int(nprob_inner)
        self.H, self.W, self.T = int(H), int(W), int(T)
        self.storage = storage
        self.planes = bool(planes)
        d = N.Desc()
        d.ndim, d.nprob, d.nprob_inner = self.ndim, self.nprob, self.nprob_inner
        d.H, d.W, d.T = self.H, self.W, self.T
        d.storage = _DTYPE_CODE[storage]
        d.weight_mode = N.W_PLANES if planes else N.W_SCALAR
        d.k_block = int(k_block)
        self._global_H = self.H
        if shard is not None:
            d.global_H, d.row_offset, d.own_row0, d.own_rows = [int(v) for v in shard]
            self._global_H = int(shard[0]) or self.H
        d.features = (N.FEAT_RESISTANCE if resistance else 0) | (N.FEAT_BLOCKING_WAIT if blocking_wait else 0)
        self._h = C.c_void_p()
        N.check(self.lib, self.lib.vpint_solver_create(C.byref(d), C.byref(self._h)), "vpint_solver_create")
        nbytes = self.lib.vpint_solver_workspace_bytes(self._h)
        with self.torch.cuda.device(self.device):
            self.workspace = self.torch.empty(nbytes + 256, dtype=self.torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        self._ws_ptr = (base + 255) // 256 * 256
        N.check(self.lib, self.lib.vpint_solver_bind(self._h, C.c_void_p(self._ws_ptr), C.c_size_t(nbytes)),
                "vpint_solver_bind")
        self.workspace_bytes = int(nbytes)
        self._keep = []
        self.delta = None
        self.max_iter = 0

    # -- helpers -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.vpint_solver_destroy(self._h)
            self._h = None
        self.workspace = None
        self._shard_result_buf = None
        self._keep = []

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _strides(self, t, tail):
        """Element strides {outer, inner, row, column, last} of ``t`` whose shape is
        ``(nprob,) + tail`` or ``(outer, inner) + tail``; ``tail`` is (H, W) or (H, W, n)."""
        outer = self.nprob // self.nprob_inner
        shape = tuple(t.shape)
        if shape == (self.nprob,) + tail:
            ps = (t.stride(0) * self.nprob_inner, t.stride(0))
            rest = [t.stride(i) for i in range(1, t.dim())]
        elif shape == (outer, self.nprob_inner) + tail:
            ps = (t.stride(0), t.stride(1))
            rest = [t.stride(i) for i in range(2, t.dim())]
        else:
            raise VPintError("tensor shape %s does not match the solver layout %s" % (shape, (self.nprob,) + tail))
        if len(rest) == 2:
            rest.append(0)
        return [ps[0], ps[1]] + rest

    def _cell_tail(self):
        return (self.H, self.W) if self.ndim == 2 else (self.H, self.W, self.T)

    # -- C ABI calls ---------------------------------------------------------
    def load(self, original, pred0=None, fill=None):
        """Stage the start state.  ``original``: NaN = unknown.  Either ``pred0``
        (same layout) or ``fill`` (one value per problem for the unknown cells)."""
        ov = original
        st = self._strides(ov, self._cell_tail())
        pptr = None
        if pred0 is not None:
            pv = pred0 if pred0.dtype == ov.dtype else pred0.to(ov.dtype)
            if tuple(pv.shape) != tuple(ov.shape):
                raise VPintError("pred0 and original must have the same shape")
            if self._strides(pv, self._cell_tail()) != st:
                pv = pv.contiguous()
                ov = ov.contiguous()
                st = self._strides(ov, self._cell_tail())
            self._keep.append(pv)
            pptr = C.c_void_p(pv.data_ptr())
            fptr = None
        elif fill is None:
            fptr = None       # mean init on the device (np.nanmean per problem, NumPy's summation order)
        else:
            f = np.ascontiguousarray(np.asarray(fill, dtype=np.float64).reshape(-1))
            if f.size != self.nprob:
                raise VPintError("fill needs one value per problem")
            fptr = f.ctypes.data_as(C.POINTER(C.c_double))
        del self._keep[:-8]            # references only have to outlive the enqueue of the kernels that read them
        self._keep.append(ov)
        N.check(self.lib, self.lib.vpint_solver_load(self._h, C.c_void_p(ov.data_ptr()), pptr, fptr, _code(ov),
                                                      _i64(st), self._stream()), "vpint_solver_load")
        return self

    def weights(self, feat, method="exact", clamp=False, min_gamma=0.0, max_gamma=math.inf):
        """feat viewed as (outer, inner, H, W, F)."""
        if method not in N.METHODS:
            raise VPintError("Invalid weight prediction method: " + str(method))
        F = int(feat.shape[-1])
        fv = feat
        st = self._strides(fv, (self.H, self.W, F))
        self._keep.append(fv)
        N.check(self.lib, self.lib.vpint_solver_weights(self._h, C.c_void_p(fv.data_ptr()), _code(fv), _i64(st), int(F),
                                                         N.METHODS[method], int(bool(clamp)), float(min_gamma),
                                                         float(max_gamma), self._stream()), "vpint_solver_weights")
        return self

    def discounts(self, gamma, tau=None):
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gamma, dtype=np.float64), (self.nprob,)))
        gp = g.ctypes.data_as(C.POINTER(C.c_double))
        tp = None
        if tau is not None:
            t = np.ascontiguousarray(np.broadcast_to(np.asarray(tau, dtype=np.float64), (self.nprob,)))
            tp = t.ctypes.data_as(C.POINTER(C.c_double))
        N.check(self.lib, self.lib.vpint_solver_set_discounts(self._h, gp, tp, self._stream()), "vpint_solver_set_discounts")
        return self

    def trial_params(self, beta=None, epsilon=None, mu=None):
        """Per-problem beta and / or (epsilon, mu) for batched trials of a parameter search (arrays of nprob values)."""
        def arr(v):
            if v is None:
                return None, None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.nprob,)))
            return a, a.ctypes.data_as(C.POINTER(C.c_double))
        (b, bp), (e, ep), (m, mp) = arr(beta), arr(epsilon), arr(mu)
        N.check(self.lib, self.lib.vpint_solver_set_trial_params(self._h, bp, ep, mp, self._stream()),
                "vpint_solver_set_trial_params")
        return self

    def priority(self, beta, bias=None, want_planes=False):
        """prioritise_identity: fold the priority planes into the weights (after :meth:`weights`).
        ``bias``: CUDA float64 tensor (nprob, H, W) of the known_value_bias sub-solve, or None.
        Returns the (nprob, H, W, 4) priority planes when ``want_planes``."""
        p_out = None
        if want_planes:
            p_out = self.torch.empty((self.nprob, self.H, self.W, 4), dtype=self.torch.float64, device=self.device)
        if bias is not None:
            bias = bias.to(self.torch.float64).contiguous()
            if tuple(bias.shape) != (self.nprob, self.H, self.W):
                raise VPintError("bias must have shape (nprob, H, W)")
            self._keep.append(bias)
        N.check(self.lib, self.lib.vpint_solver_priority(
            self._h, float(beta), C.c_void_p(bias.data_ptr()) if bias is not None else None,
            C.c_void_p(p_out.data_ptr()) if p_out is not None else None, self._stream()), "vpint_solver_priority")
        return p_out

    def _params(self, max_iter, auto_terminate, threshold, clip_val, track_delta, loop_mode, chunk,
                resistance=None):
        prm = N.RunParams()
        if resistance is not None:
            prm.resistance, prm.epsilon, prm.mu = 1, float(resistance[0]), float(resistance[1])
        prm.max_iter = int(max_iter)
        prm.auto_terminate = int(bool(auto_terminate))
        prm.threshold = float(threshold)
        prm.clip_val = float(clip_val)
        prm.loop_mode = int(loop_mode)
        prm.chunk = int(chunk)
        self.max_iter = int(max_iter)
        if track_delta and max_iter > 0:
            self.delta = self.torch.zeros((self.nprob, int(max_iter)), dtype=self.torch.float64, device=self.device)
            prm.delta_out = self.delta.data_ptr()
        else:
            self.delta = None
            prm.delta_out = None
        return prm

    def run(self, max_iter, auto_terminate=True, threshold=1e-4, clip_val=math.inf, track_delta=False,
            loop_mode=None, chunk=0, resistance=None):
        """``resistance``: None or (epsilon, mu) (solver must have been created with ``resistance=True``)."""
        if loop_mode is None:
            loop_mode = DEFAULT_LOOP_MODE
        prm = self._params(max_iter, auto_terminate, threshold, clip_val, track_delta, loop_mode, chunk, resistance)
        N.check(self.lib, self.lib.vpint_solver_run(self._h, C.byref(prm), self._stream()), "vpint_solver_run")
        return self

    def run_async(self, max_iter, auto_terminate=True, threshold=1e-4, clip_val=math.inf, track_delta=False,
                  loop_mode=None, chunk=0, resistance=None):
        """Non-blocking :meth:`run` where the cluster-resident kernel serves the solve; :meth:`run_finish` completes it."""
        if loop_mode is None:
            loop_mode = DEFAULT_LOOP_MODE
        prm = self._params(max_iter, auto_terminate, threshold, clip_val, track_delta, loop_mode, chunk, resistance)
        N.check(self.lib, self.lib.vpint_solver_run_async(self._h, C.byref(prm), self._stream()), "vpint_solver_run_async")
        return self

    def run_finish(self):
        """Waits for :meth:`run_async`.  True when the run had to be completed on the tiled path (results read
        before must be read again)."""
        reran = C.c_int(0)
        N.check(self.lib, self.lib.vpint_solver_run_finish(self._h, self._stream(), C.byref(reran)), "vpint_solver_run_finish")
        return bool(reran.value)

    # -- fused VPint2 load ---------------------------------------------------------------------------
    def fill_scratch(self, value_dtype="float32"):
        """Scratch tensor for the on-device np.nanmean (leaf sums of NumPy's summation tree + counts)."""
        cells = self._global_H * self.W
        nbytes = self.lib.vpint_fill_scratch_bytes(cells, self.nprob, _DTYPE_CODE[value_dtype])
        return self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)

    def fill_views(self, scratch, value_dtype="float32"):
        """(leaf sums, counts) views of ``scratch`` for the all-reduce of a row-sharded mean."""
        torch = self.torch
        cells = self._global_H * self.W
        n = int(self.lib.vpint_fill_leaf_count(cells, self.nprob))
        off = int(self.lib.vpint_fill_count_offset(cells, self.nprob, _DTYPE_CODE[value_dtype]))
        esz, dt = (8, torch.float64) if value_dtype == "float64" else (4, torch.float32)
        return scratch[:n * esz].view(dt), scratch[off:off + 8 * self.nprob].view(torch.int64)

    def _ingest_params(self, target, features, cloud, value_dtype, clip_target, clip_features, fill, fill_ready, out, scratch):
        p = N.IngestParams()
        lead = self.nprob // self.nprob_inner
        shape = (lead, self.H, self.W, self.nprob_inner)
        for name, t in (("target", target), ("features", features)):
            if t is None:
                continue
            if tuple(t.shape) != shape or not t.is_contiguous():
                raise VPintError("%s must be a contiguous (N, H, W, B) = %s raster, got %s" % (name, shape, tuple(t.shape)))
        if features is not None and features.dtype != target.dtype:
            raise VPintError("target and features must have the same raster dtype")
        p.target = target.data_ptr()
        p.features = features.data_ptr() if features is not None else None
        p.raster_dtype = _raster_code(target)
        if p.raster_dtype == N.U8:
            raise VPintError("uint8 rasters are not supported (uint16 / float32 / float64)")
        p.value_dtype = _DTYPE_CODE[value_dtype]
        if cloud is not None:
            if tuple(cloud.shape) != shape[:3] or cloud.dtype != self.torch.uint8 or not cloud.is_contiguous():
                raise VPintError("cloud must be a contiguous uint8 (N, H, W) tensor")
            p.cloud = cloud.data_ptr()
        p.clip_target, p.target_min, p.target_max = _clip_pair(clip_target)
        p.clip_features, p.features_min, p.features_max = _clip_pair(clip_features)
        keep = [target, features, cloud]
        if fill_ready:
            p.fill_mode = N.FILL_READY
        elif fill is not None:
            f = np.ascontiguousarray(np.asarray(fill, dtype=np.float64).reshape(-1))
            if f.size != self.nprob:
                raise VPintError("fill needs one value per problem")
            p.fill_mode = N.FILL_HOST
            p.fill_host = f.ctypes.data_as(C.POINTER(C.c_double))
            keep.append(f)
        else:
            p.fill_mode = N.FILL_DEVICE
        if scratch is not None:
            p.scratch = scratch.data_ptr()
            keep.append(scratch)
        if out is not None:
            st = self._strides(out, (self.H, self.W))[:4]
            p.out = out.data_ptr()
            p.out_dtype = _out_code(out)
            for i in range(4):
                p.out_stride[i] = int(st[i])
            keep.append(out)
        return p, keep

    def ingest_bands(self, target, features, cloud=None, *, value_dtype="float32", clip_target=None, clip_features=None,
                     fill=None, fill_ready=False, out=None, scratch=None):
        """Fused load of a band-interleaved stack straight from the raw rasters (uint16 / float32 / float64 CUDA
        tensors (N, H, W, B)): VPint2_interpolator's constructor steps, the per-band mean init and the feature plane
        in one pass, no host synchronisation.  ``cloud``: uint8 (N, H, W), nonzero = unknown on every band.
        ``out``: tensor in any layout :meth:`result` accepts; receives the KNOWN cells of the result."""
        if fill is None and not fill_ready and scratch is None:
            scratch = self.fill_scratch(value_dtype)
        p, keep = self._ingest_params(target, features, cloud, value_dtype, clip_target, clip_features, fill, fill_ready,
                                      out, scratch)
        self._keep = keep
        N.check(self.lib, self.lib.vpint_solver_ingest_bands(self._h, C.byref(p), self._stream()), "vpint_solver_ingest_bands")
        return self

    def fill_leaves(self, target, cloud, scratch, *, value_dtype="float32", clip_target=None):
        p, keep = self._ingest_params(target, None, cloud, value_dtype, clip_target, None, None, False, None, scratch)
        self._keep += keep
        N.check(self.lib, self.lib.vpint_solver_fill_leaves(self._h, C.byref(p), self._stream()), "vpint_solver_fill_leaves")

    def fill_combine(self, scratch, value_dtype="float32"):
        N.check(self.lib, self.lib.vpint_solver_fill_combine(self._h, _DTYPE_CODE[value_dtype], C.c_void_p(scratch.data_ptr()),
                                                            self._stream()), "vpint_solver_fill_combine")

    def fill_tensor(self):
        """Device float64 [nprob]: the start value of every problem's unknown cells."""
        ptr = int(self.lib.vpint_solver_fill_ptr(self._h))
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off:off + 8 * self.nprob].view(self.torch.float64)

    def result_unknown(self, out, niter=None):
        """Writes ONLY the unknown cells of the current state into ``out`` (layouts as :meth:`result`); ``out`` may be a
        CUDA tensor or a pinned host tensor (written through its device mapping).  Returns sweeps[int32, nprob]."""
        torch = self.torch
        st = self._strides(out, (self.H, self.W))[:4]
        if niter is None:
            niter = torch.zeros(self.nprob, dtype=torch.int32, device=self.device)
        if out.device.type == "cpu" and not out.is_pinned():
            raise VPintError("result_unknown writes through a device mapping: a host `out` must be pinned memory")
        N.check(self.lib, self.lib.vpint_solver_result_unknown(self._h, C.c_void_p(out.data_ptr()), _out_code(out), _i64(st),
                                                                C.c_void_p(niter.data_ptr()), self._stream()),
                "vpint_solver_result_unknown")
        return niter

    def begin_run(self, max_iter, auto_terminate=True, threshold=1e-4, clip_val=math.inf, track_delta=False):
        prm = self._params(max_iter, auto_terminate, threshold, clip_val, track_delta, N.LOOP_CHUNKED, 0)
        N.check(self.lib, self.lib.vpint_solver_begin_run(self._h, C.byref(prm), self._stream()), "vpint_solver_begin_run")

    def step_begin(self):
        N.check(self.lib, self.lib.vpint_solver_step_begin(self._h, self._stream()), "vpint_solver_step_begin")

    def step_profile(self):
        """step_begin that returns the stencil kernel's duration in ms (CUDA events on this stream)."""
        ms = C.c_float(0.0)
        N.check(self.lib, self.lib.vpint_solver_step_profile(self._h, self._stream(), C.byref(ms)),
                "vpint_solver_step_profile")
        return float(ms.value)

    def step_end(self):
        N.check(self.lib, self.lib.vpint_solver_step_end(self._h, self._stream()), "vpint_solver_step_end")

    def step_commit_lagged(self, sums_prev):
        """Lagged stop decision (sums of the PREVIOUS sweep, all-reduced, or None in the first launch of a run) + commit
        of the sweep the last step_begin computed."""
        N.check(self.lib, self.lib.vpint_solver_step_commit_lagged(
            self._h, C.c_void_p(sums_prev.data_ptr()) if sums_prev is not None else None, self._stream()),
            "vpint_solver_step_commit_lagged")

    def halo_push(self, rows, peer_up, flag_up, peer_down, flag_down, seq):
        """Boundary rows of the current state into the neighbours' receive slots (raw device pointers as ints, 0 = none)."""
        N.check(self.lib, self.lib.vpint_solver_halo_push(
            self._h, int(rows), C.c_void_p(peer_up or None), C.c_void_p(flag_up or None), C.c_void_p(peer_down or None),
            C.c_void_p(flag_down or None), int(seq), self._stream()), "vpint_solver_halo_push")

    def halo_pull(self, rows, from_up, flag_up, from_down, flag_down, seq):
        N.check(self.lib, self.lib.vpint_solver_halo_pull(
            self._h, int(rows), C.c_void_p(from_up or None), C.c_void_p(flag_up or None), C.c_void_p(from_down or None),
            C.c_void_p(flag_down or None), int(seq), self._stream()), "vpint_solver_halo_pull")

    def peer_bytes(self, world):
        return int(self.lib.vpint_solver_peer_bytes(self._h, int(world)))

    def peer_bind(self, world, rank, base_ptrs):
        """base_ptrs: every rank's peer block (device pointers valid in THIS process), in rank = row order."""
        arr = (C.c_void_p * int(world))(*[C.c_void_p(int(p)) for p in base_ptrs])
        N.check(self.lib, self.lib.vpint_solver_peer_bind(self._h, int(world), int(rank), arr), "vpint_solver_peer_bind")

    def run_sharded(self, max_iter, auto_terminate=True, threshold=1e-4, clip_val=math.inf):
        """The row-sharded loop inside the library (after :meth:`peer_bind`; known-cell sums already all-reduced).
        Blocks until every problem has stopped on every rank; returns the launch groups issued."""
        prm = self._params(max_iter, auto_terminate, threshold, clip_val, False, N.LOOP_CHUNKED, 0)
        n = C.c_int64(0)
        N.check(self.lib, self.lib.vpint_solver_run_sharded(self._h, C.byref(prm), self._stream(), C.byref(n)),
                "vpint_solver_run_sharded")
        return int(n.value)

    def result(self, out=None, dtype=None):
        """Current state of every problem gathered into ``out`` (same layouts as
        :meth:`load` accepts).  Returns (out, sweeps[int32, nprob])."""
        torch = self.torch
        if out is None:
            out = torch.empty((self.nprob,) + self._cell_tail(), dtype=dtype or torch.float64, device=self.device)
        st = self._strides(out, self._cell_tail())
        niter = torch.zeros(self.nprob, dtype=torch.int32, device=self.device)
        N.check(self.lib, self.lib.vpint_solver_result(self._h, C.c_void_p(out.data_ptr()), _code(out), _i64(st),
                                                        C.c_void_p(niter.data_ptr()), self._stream()), "vpint_solver_result")
        return out, niter

    # -- introspection -------------------------------------------------------
    @property
    def active_tiles(self):
        return int(self.lib.vpint_solver_active_tiles(self._h))

    @property
    def total_tiles(self):
        return int(self.lib.vpint_solver_total_tiles(self._h))

    @property
    def launches(self):
        return int(self.lib.vpint_solver_launches(self._h))

    @property
    def k_block(self):
        return int(self.lib.vpint_solver_k_block(self._h))

    def sums_tensor(self):
        """[nprob][k_block][4] fp64 per-sweep local sums of the last step_begin (a view
        of workspace memory, for the all-reduce of a row-sharded solve)."""
        n = int(self.lib.vpint_solver_sums_count(self._h))
        ptr = int(self.lib.vpint_solver_sums_ptr(self._h))
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off:off + n * 8].view(self.torch.float64).view(self.nprob, -1, 4)

    def known_tensor(self):
        """[2][nprob][4] fp64 known-cell sums of the owned rows (workspace view; all-reduced once by a
        row-sharded run)."""
        n = int(self.lib.vpint_solver_known_count(self._h))
        ptr = int(self.lib.vpint_solver_known_ptr(self._h))
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off:off + n * 8].view(self.torch.float64)

    @property
    def any_dirty(self):
        return int(self.lib.vpint_solver_any_dirty(self._h))

    def halo_pack(self, rows, to_up, to_down):
        N.check(self.lib, self.lib.vpint_solver_halo_pack(
            self._h, int(rows), C.c_void_p(to_up.data_ptr()) if to_up is not None else None,
            C.c_void_p(to_down.data_ptr()) if to_down is not None else None, self._stream()), "vpint_solver_halo_pack")

    def halo_unpack(self, rows, from_up, from_down):
        N.check(self.lib, self.lib.vpint_solver_halo_unpack(
            self._h, int(rows), C.c_void_p(from_up.data_ptr()) if from_up is not None else None,
            C.c_void_p(from_down.data_ptr()) if from_down is not None else None, self._stream()), "vpint_solver_halo_unpack")

    def state_tensor(self, which):
        """Ping-pong buffer ``which`` as a [slices][H][pitch] tensor (workspace view); slices = nprob in 2-D,
        nprob * T (time-major) in 3-D."""
        ptr = int(self.lib.vpint_solver_state_ptr(self._h, int(which)))
        pitch = int(self.lib.vpint_row_pitch(self.W))
        dt = self.torch.float64 if self.storage == "float64" else self.torch.float32
        esz = 8 if self.storage == "float64" else 4
        off = ptr - self.workspace.data_ptr()
        slices = self.nprob * (self.T if self.ndim == 3 else 1)
        n = slices * self.H * pitch
        return self.workspace[off:off + n * esz].view(dt).view(slices, self.H, pitch)

    def active_tensor(self):
        """Device int32[1]: problems still running (a view of workspace memory)."""
        ptr = int(self.lib.vpint_solver_active_ptr(self._h))
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off:off + 4].view(self.torch.int32)

    def active_count(self):
        ptr = int(self.lib.vpint_solver_active_ptr(self._h))
        off = ptr - self.workspace.data_ptr()
        return int(self.workspace[off:off + 4].view(self.torch.int32).item())
