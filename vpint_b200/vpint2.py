"""VPint2 cloud-removal wrapper: drop-in ``VPint2_interpolator``.

Mirrors VPint/VPint2.py.  The reference splits an (H, W, B) scene into B
independent ``WP_SMRP`` problems and forks one process per band
(VPint/VPint2.py:58-106); here all bands -- and, through :func:`run_patches`, all
bands of many patches -- are ONE batched device solve with per-problem stopping.
"""

from __future__ import annotations

import os

import numpy as np

from . import solve as _solve
from .engine import Solver, require_cuda, to_device
from .errors import VPintError
from .wp import wp_run_options


# Largest grid whose init_strategy='mean' value the library computes on the device (NumPy-exact pairwise
# nanmean, csrc/vpint_setup.cu fill_mean_kernel); larger scenes take it on the host.
DEVICE_FILL_MAX_CELLS = 131072


def _band_fill_values(target):
    """Per-band start value of the unknown cells: ``WP_SMRP(target[:, :, b], ...)``
    initialises them with ``np.nanmean`` of the band IN THE BAND'S dtype
    (VPint/VPint2.py:105 -> VPint/MRP.py:73-77).  target: (..., H, W, B)."""
    lead = target.shape[:-3]
    B = target.shape[-1]
    flat = target.reshape((-1,) + target.shape[-3:])
    out = np.empty((flat.shape[0], B), dtype=np.float64)
    for n in range(flat.shape[0]):
        for b in range(B):
            out[n, b] = np.nanmean(flat[n, :, :, b])
    return out.reshape(lead + (B,))


# Patches per pipeline stage when a host-resident stack is solved chunk by chunk (H2D of chunk i+1 and D2H
# of chunk i-1 overlap the solve of chunk i).
PIPELINE_CHUNK_PATCHES = 16
# Host threads (each with its own stream and solver) that take alternate chunks.
PIPELINE_WORKERS = 6


def _pipeline_blocking_waits(n_workers):
    """Should the pipeline threads sleep in their device waits instead of spinning?  Spinning answers a few
    microseconds sooner (one GPU, 16 cores: 87 ms per step against 93 ms), but only while the waiting threads leave
    half of the cores to everything else that runs (NCCL and torch helper threads, the copies): with several ranks
    per node (LOCAL_WORLD_SIZE) the spinning threads of all ranks fight for the cores the process may run on
    (4 ranks x 7 threads on 32 cores: 230 ms spinning, 197 ms sleeping).  VPINT_PIPELINE_BLOCKING=0/1 overrides."""
    env = os.environ.get("VPINT_PIPELINE_BLOCKING")
    if env in ("0", "1"):
        return env == "1"
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    ranks = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return 2 * (n_workers + 1) * ranks > cores


def _known_value_bias(t_dev, kvb, dev):
    """known_value_bias sub-solve (VPint/WP_MRP.py:263-272) for a stack (N, H, W, B) of patches: SD_SMRP on the
    known-cell indicator of every band (init 'zero', gamma = 1 - kvb, default run()); (N * B, H, W) float64."""
    import torch
    N, H, W, B = t_dev.shape
    if H != W:      # the reference's SD_SMRP only accepts square grids (VPint/SD_MRP.py:95-96)
        raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) " % (H, W))
    ind = torch.where(torch.isnan(t_dev), t_dev, torch.ones_like(t_dev)).to(torch.float64)
    sub = Solver(2, N * B, H, W, nprob_inner=B, planes=False, k_block=_solve.DEFAULT_K_BLOCK, device=dev)
    try:
        sub.load(ind.permute(0, 3, 1, 2), fill=np.zeros(N * B))
        sub.discounts(1 - kvb)
        sub.run(10000, auto_terminate=True, threshold=1e-4)
        out, _ = sub.result()
    finally:
        sub.close()
    return out


def _make_solver(nprob, H, W, B, storage, kb, dev, opts, blocking_wait=False):
    if opts.get("resistance") and kb not in (0, 1):
        kb = 1
    return Solver(2, nprob, H, W, nprob_inner=B, storage=storage, planes=True, k_block=kb, device=dev,
                  resistance=bool(opts.get("resistance")), blocking_wait=blocking_wait)


def _solve_chunk(solver, t_dev, f_dev, opts, res_dev, out_layout, fill):
    solver.load(t_dev.permute(0, 3, 1, 2), fill=fill)
    solver.weights(f_dev.permute(0, 3, 1, 2).unsqueeze(-1), method=opts["method"],
                   clamp=(opts["method"] != "exact"), min_gamma=0.0, max_gamma=float("inf"))
    if opts.get("prioritise_identity"):
        bias = None
        if opts.get("known_value_bias", 0) > 0:
            bias = _known_value_bias(t_dev, opts["known_value_bias"], solver.device)
        solver.priority(opts["priority_intensity"], bias=bias)
    solver.run(opts["iterations"], auto_terminate=opts["auto_terminate"], threshold=opts["threshold"],
               clip_val=opts["clip_val"],
               resistance=(opts["epsilon"], opts["mu"]) if opts.get("resistance") else None)
    # run() blocks the host until every problem of the chunk has stopped
    if res_dev is None:
        return None
    _, niter = solver.result(out=res_dev if out_layout == "bhw" else res_dev.permute(0, 3, 1, 2))
    return niter


def _solve_chunk_noresult(solver, t_dev, f_dev, opts, fill):
    return _solve_chunk(solver, t_dev, f_dev, opts, None, None, fill)


def _solve_bands_torch(target, features, opts, out_layout, out_dtype, storage, k_block, dev, out, fill,
                       chunk_patches=None):
    """solve_bands for torch inputs of shape (N, H, W, B).

    CUDA inputs (or a small stack): one batched solve.  Host inputs (ideally pinned): the stack is cut into
    chunks of ``chunk_patches`` patches and pipelined over three streams -- host->device copy of chunk
    i+1, solve of chunk i, device->host copy of chunk i-1 -- so the PCIe transfers hide behind the solves.
    ``out`` (preallocated, ideally pinned) receives the result."""
    import torch
    N, H, W, B = target.shape
    kb = _solve.DEFAULT_K_BLOCK if k_block is None else k_block
    if fill is None and H * W > DEVICE_FILL_MAX_CELLS:
        fill = _band_fill_values(target.cpu().numpy()).reshape(-1)
    fill = None if fill is None else np.asarray(fill, dtype=np.float64).reshape(N, B)
    chunk = int(chunk_patches or PIPELINE_CHUNK_PATCHES)
    res_shape = (lambda m: (m, B, H, W)) if out_layout == "bhw" else (lambda m: (m, H, W, B))
    res_dtype = torch.float64 if out_layout == "bhw" else out_dtype

    if target.device.type != "cpu" or N <= chunk:
        solver = _make_solver(N * B, H, W, B, storage, kb, dev, opts)
        try:
            t_dev = target.to(dev, non_blocking=True)
            f_dev = features.to(dev, non_blocking=True)
            res_dev = torch.empty(res_shape(N), dtype=res_dtype, device=dev)
            niter = _solve_chunk(solver, t_dev, f_dev, opts, res_dev, out_layout, None if fill is None else fill.reshape(-1))
            if out is not None:
                out.copy_(res_dev, non_blocking=True)
                sweeps = niter.cpu()          # synchronises the stream: `out` is complete after this
                res = out
            else:
                res = res_dev.cpu()
                sweeps = niter.cpu()
        finally:
            solver.close()
        return res, sweeps.to(torch.int64).view(N, B)

    # ---- pipelined path ---------------------------------------------------------------------------
    # PIPELINE_WORKERS host threads, each with its own stream and solver, take alternate chunks.  A thread blocks in
    # run() until its chunk has stopped; meanwhile the other thread has already copied in and queued the
    # next chunk, whose cluster-resident kernel moves onto the SMs as the first one drains (the long-tailed
    # problems of a chunk no longer leave the GPU idle), and the device->host copy of a finished chunk
    # overlaps the next solve.
    import threading
    if out is None:
        out = torch.empty(res_shape(N), dtype=res_dtype)
    sweeps_host = torch.empty(N * B, dtype=torch.int32)
    if out.is_pinned():
        sweeps_host = sweeps_host.pin_memory()
    spans = [(a, min(a + chunk, N)) for a in range(0, N, chunk)]
    main = torch.cuda.current_stream(dev)
    n_workers = PIPELINE_WORKERS
    blocking = _pipeline_blocking_waits(n_workers)
    errors = []

    def worker(w):
        try:
            with torch.cuda.device(dev):
                st = torch.cuda.Stream(device=dev)
                st.wait_stream(main)
                solvers = {}
                t_buf = torch.empty((chunk, H, W, B), dtype=target.dtype, device=dev)
                f_buf = torch.empty((chunk, H, W, B), dtype=features.dtype, device=dev)
                r_buf = torch.empty(res_shape(chunk), dtype=res_dtype, device=dev)
                done = torch.cuda.Event(blocking=True)
                try:
                    with torch.cuda.stream(st):
                        for i in range(w, len(spans), n_workers):
                            a, b = spans[i]
                            m = b - a
                            t_buf[:m].copy_(target[a:b], non_blocking=True)
                            f_buf[:m].copy_(features[a:b], non_blocking=True)
                            if m not in solvers:
                                solvers[m] = _make_solver(m * B, H, W, B, storage, kb, dev, opts, blocking_wait=blocking)
                            niter = _solve_chunk(solvers[m], t_buf[:m], f_buf[:m], opts, r_buf[:m], out_layout,
                                                 None if fill is None else fill[a:b].reshape(-1))
                            out[a:b].copy_(r_buf[:m], non_blocking=True)
                            sweeps_host[a * B:b * B].copy_(niter, non_blocking=True)
                            if blocking:              # r_buf / niter are reused by this thread's next chunk
                                done.record(st)
                                done.synchronize()
                            else:
                                st.synchronize()
                finally:
                    for sv in solvers.values():
                        sv.close()
        except BaseException as exc:      # surfaced to the caller below
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(min(n_workers, len(spans)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out, sweeps_host.to(torch.int64).view(N, B)


def solve_bands(target, features, opts, *, out_dtype=np.float64, out_layout="bhw", storage="float64",
                k_block=None, device=None, return_sweeps=False, out=None, fill=None):
    """All bands of one scene (H, W, B) or of a stack of patches (N, H, W, B) as one
    batched solve.  ``target`` / ``features`` are the interpolator's stored arrays
    (DTYPE, NaN = cloud).  Returns float64 (..., B, H, W) for out_layout 'bhw' or
    ``out_dtype`` (..., H, W, B) for 'hwb'.

    torch tensors (e.g. pinned host buffers) of shape (N, H, W, B) are accepted too:
    then copies are asynchronous, ``out`` (a preallocated, ideally pinned, tensor)
    receives the result and torch tensors are returned."""
    dev = require_cuda(device)
    import torch
    if isinstance(target, torch.Tensor):
        if opts["iterations"] == 0:
            raise VPintError("iterations=0 is handled by the NumPy entry points")
        tdt = out_dtype if isinstance(out_dtype, torch.dtype) else {np.dtype(np.float32): torch.float32}.get(
            np.dtype(out_dtype), torch.float64)
        res, sweeps = _solve_bands_torch(target, features, opts, out_layout, tdt, storage, k_block, dev, out, fill)
        return (res, sweeps) if return_sweeps else res
    single = (target.ndim == 3)
    tgt = target[None] if single else target
    fts = features[None] if single else features
    # WP_SMRP casts the features with astype(float) (VPint/WP_MRP.py:67-71): integer rasters are fine on either side
    if tgt.dtype not in (np.float32, np.float64):
        tgt = tgt.astype(np.float64)
    if fts.dtype not in (np.float32, np.float64):
        fts = fts.astype(np.float64)
    N, H, W, B = tgt.shape
    max_iter = opts["iterations"]
    if max_iter == 0:
        # the reference returns each band's initial pred_grid (NaN -> band mean)
        fill = _band_fill_values(tgt)
        res = np.where(np.isnan(tgt), fill[:, None, None, :].astype(tgt.dtype), tgt)
        if out_layout == "bhw":
            res = np.ascontiguousarray(np.moveaxis(res, -1, 1)).astype(tgt.dtype)
        else:
            res = res.astype(out_dtype)
        res = res[0] if single else res
        return (res, np.zeros((N, B), dtype=np.int64)[0 if single else slice(None)]) if return_sweeps else res
    fill = None if H * W <= DEVICE_FILL_MAX_CELLS else _band_fill_values(tgt).reshape(-1)
    solver = _make_solver(N * B, H, W, B, storage, _solve.DEFAULT_K_BLOCK if k_block is None else k_block, dev, opts)
    try:
        t_dev = to_device(np.ascontiguousarray(tgt), dev)                 # (N, H, W, B) as stored
        f_dev = to_device(np.ascontiguousarray(fts), dev)
        niter = _solve_chunk_noresult(solver, t_dev, f_dev, opts, fill)   # views (N, B, H, W): no transpose copy
        if out_layout == "bhw":
            out = torch.empty((N, B, H, W), dtype=torch.float64, device=dev)
            _, niter = solver.result(out=out)
        else:
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}.get(np.dtype(out_dtype))
            out = torch.empty((N, H, W, B), dtype=tdt or torch.float64, device=dev)
            _, niter = solver.result(out=out.permute(0, 3, 1, 2))
        res = out.cpu().numpy()
        if out_layout == "hwb" and res.dtype != np.dtype(out_dtype):
            res = res.astype(out_dtype)
        sweeps = niter.cpu().numpy().astype(np.int64).reshape(N, B)
    finally:
        solver.close()
    if single:
        res, sweeps = res[0], sweeps[0]
    return (res, sweeps) if return_sweeps else res


class VPint2_interpolator():
    """Reference: VPint/VPint2.py:12-150.  Constructor semantics are unchanged
    (dtype cast, optional axis move, clipping, cloud mask -> NaN on every band,
    optional cloud buffering); only the per-pixel Python loops are vectorised."""

    def __init__(self, target, features, mask=None, buffer_mask=False, mask_buffer_size=5, mask_buffer_passes=1,
                 bands_first=False, clip_target=None, clip_features=None, dtype=None, **kwargs):
        assert target.shape == features.shape
        assert len(target.shape) == 3
        assert len(features.shape) == 3

        self.DTYPE = dtype if dtype is not None else np.float32
        target = target.astype(self.DTYPE)
        features = features.astype(self.DTYPE)
        if bands_first:
            target = np.moveaxis(target, 0, -1)
            features = np.moveaxis(features, 0, -1)

        # clip_* : int / float = upper bound only, otherwise (min, max)   (VPint/VPint2.py:30-41)
        if clip_target is not None:
            if type(clip_target) == type(2) or type(clip_target) == type(2.1):
                target = np.clip(target, a_min=-np.inf, a_max=clip_target)
            else:
                target = np.clip(target, a_min=clip_target[0], a_max=clip_target[1])
        if clip_features is not None:
            if type(clip_features) == type(2) or type(clip_features) == type(2.1):
                features = np.clip(features, a_min=-np.inf, a_max=clip_features)
            else:
                features = np.clip(features, a_min=clip_features[0], a_max=clip_features[1])

        if mask is not None:
            assert len(mask.shape) == 2
            assert mask.shape[0] == target.shape[0]
            assert mask.shape[1] == target.shape[1]
            self.target = self.apply_mask(target, mask, **kwargs).astype(self.DTYPE)
        else:
            self.target = target.astype(self.DTYPE)

        if buffer_mask:
            self.target = self.buffer_clouds(self.target, buffer_size=mask_buffer_size, passes=mask_buffer_passes)

        self.features = features.astype(self.DTYPE)

    def run(self, **kwargs):
        """All bands in one batched device solve (the reference forks one process per
        band, VPint/VPint2.py:58-101).  Returns float64 (H, W, B), a strided view of a
        (B, H, W) array exactly like the reference's swapaxes result."""
        opts = wp_run_options(**kwargs)
        if opts["auto_adapt"]:
            return self._run_bands_with_search(kwargs, forked=True).swapaxes(0, 2).swapaxes(0, 1)
        pred_bhw = solve_bands(self.target, self.features, opts, out_layout="bhw")
        return pred_bhw.swapaxes(0, 2).swapaxes(0, 1)

    def _run_bands_with_search(self, kwargs, forked):
        """auto_adapt=True: every band runs its own host-side parameter search (VPint/WP_MRP.py:212-239), so
        the bands go through ``WP_SMRP.run`` one after the other (each trial is a device solve).  The
        reference's ``run`` forks one process per band: every child starts from the SAME global RNG state
        and the parent's state does not advance; ``run_serial`` shares one advancing stream.  Both are kept."""
        from .wp import WP_SMRP
        H, W, B = self.target.shape
        out = np.empty((B, H, W), dtype=np.float64)
        state = np.random.get_state()
        for b in range(B):
            if forked:
                np.random.set_state(state)
            out[b] = WP_SMRP(self.target[:, :, b], self.features[:, :, b]).run(**kwargs)
        if forked:
            np.random.set_state(state)
        return out

    def VPint2_single(self, pred_dict, grids, band, **kwargs):
        """Kept for API compatibility (VPint/VPint2.py:104-106)."""
        from .wp import WP_SMRP
        interp = WP_SMRP(grids[0], grids[1])
        pred_dict[band] = interp.run(**kwargs)

    def run_serial(self, **kwargs):
        """Same results as :meth:`run`, returned as a DTYPE (H, W, B) array
        (VPint/VPint2.py:109-119)."""
        opts = wp_run_options(**kwargs)
        if opts["auto_adapt"]:
            pred = self._run_bands_with_search(kwargs, forked=False)
            return np.ascontiguousarray(np.moveaxis(pred, 0, -1)).astype(self.DTYPE)
        return solve_bands(self.target, self.features, opts, out_dtype=self.DTYPE, out_layout="hwb")

    def apply_mask(self, target, mask, threshold=0.2):
        """mask > threshold -> NaN on every band (VPint/VPint2.py:122-132), vectorised."""
        target_cloudy = target.copy()
        target_cloudy[np.asarray(mask) > threshold, :] = np.nan
        return target_cloudy

    def buffer_clouds(self, target, buffer_size=5, passes=1):
        """Dilate the NaN footprint (VPint/VPint2.py:135-150), vectorised.  Each pixel
        with a NaN in any band blanks rows [i-b, i+b) and columns [j-b, min(H, j+b)) --
        the reference bounds the COLUMN range by shape[0] (:146); kept."""
        img = target.copy()
        H, W = target.shape[0], target.shape[1]
        b = int(buffer_size)
        for _ in range(0, passes):
            new_img = target.copy()
            cloudy = np.isnan(img).any(axis=2)
            grown = np.zeros((H, W), dtype=bool)
            if b > 0:
                # output (y, x) is covered by source (i, j) iff i-b <= y < i+b and j-b <= x < j+b: a rectangle, so the
                # dilation separates into a pass along the rows and a pass along the columns (4b shifts, not 4b^2)
                pad = np.zeros((H + 2 * b, W), dtype=bool)
                pad[b:b + H, :] = cloudy
                tall = np.zeros((H, W), dtype=bool)
                for dy in range(-b + 1, b + 1):
                    tall |= pad[b + dy:b + dy + H, :]
                pad = np.zeros((H, W + 2 * b), dtype=bool)
                pad[:, b:b + W] = tall
                for dx in range(-b + 1, b + 1):
                    grown |= pad[:, b + dx:b + dx + W]
                if H < W:
                    grown[:, H:] = False
            new_img[grown, :] = np.nan
            img = new_img
        return img


def run_patches(interpolators, serial=False, **kwargs):
    """Run many ``VPint2_interpolator`` objects of one shape as ONE batched solve
    (problem = (patch, band)).  Equivalent to ``[vp.run(**kwargs) for vp in
    interpolators]`` (or ``run_serial`` with ``serial=True``)."""
    if not interpolators:
        return []
    opts = wp_run_options(**kwargs)
    if opts["auto_adapt"]:          # per-band searches: no batching across patches
        return [vp.run_serial(**kwargs) if serial else vp.run(**kwargs) for vp in interpolators]
    shape, dt = interpolators[0].target.shape, interpolators[0].target.dtype
    for vp in interpolators:
        if vp.target.shape != shape or vp.target.dtype != dt:
            raise VPintError("run_patches needs interpolators of one shape and dtype")
    tgt = np.stack([vp.target for vp in interpolators])
    fts = np.stack([vp.features for vp in interpolators])
    if serial:
        res = solve_bands(tgt, fts, opts, out_dtype=interpolators[0].DTYPE, out_layout="hwb")
        return [res[i] for i in range(len(interpolators))]
    res = solve_bands(tgt, fts, opts, out_layout="bhw")
    return [res[i].swapaxes(0, 2).swapaxes(0, 1) for i in range(len(interpolators))]
