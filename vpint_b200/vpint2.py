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
from .engine import Solver, buffer_clouds, cloud_mask, require_cuda, to_device
from .errors import VPintError
from .wp import wp_run_options


# Largest grid whose init_strategy='mean' value vpint_solver_load computes itself (one kernel, NumPy-exact pairwise
# nanmean, csrc/vpint_setup.cu fill_mean_kernel); larger grids take the same value from the leaf / subtree kernels of
# csrc/vpint_ingest.cu (_device_band_means) -- on the device either way.
DEVICE_FILL_MAX_CELLS = 131072


def _device_band_means(solver, t_dev):
    """np.nanmean of every band of a large (N, H, W, B) float stack in the band's dtype (VPint/MRP.py:73-77), computed
    on the device and left in the solver for the next ``load`` (vpint_solver_fill_leaves / _fill_combine)."""
    import torch
    vd = {torch.float32: "float32", torch.float64: "float64"}.get(t_dev.dtype)
    if vd is None:
        raise VPintError("band stacks must be float32 or float64, got %s" % (t_dev.dtype,))
    t_dev = t_dev.contiguous()
    scratch = solver.fill_scratch(vd)
    solver.fill_leaves(t_dev, None, scratch, value_dtype=vd)
    solver.fill_combine(scratch, vd)


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
    if fill is None and t_dev.shape[1] * t_dev.shape[2] > DEVICE_FILL_MAX_CELLS:
        _device_band_means(solver, t_dev)
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
    fill = None
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


# Streams (each with its own solver and staging buffers) that take alternate chunks in :func:`solve_raw`.
RAW_STREAMS = 4


def _fused_ok(opts):
    """Can :func:`solve_raw` serve these run() options?  (exact weights, no priority / bias / resistance / search)"""
    return (opts["method"] == "exact" and not opts.get("prioritise_identity") and not opts.get("resistance")
            and not opts.get("auto_adapt") and opts["iterations"] != 0)


class _RawSlot:
    """One stream of the :func:`solve_raw` pipeline: staging buffers, solver and the chunk in flight."""

    def __init__(self, dev, chunk, H, W, B, raster_dtype, mask_dtype, res_shape, res_dtype, need_res, need_cloud, buffer,
                 stage_in=True):
        import torch
        self.dev = dev
        # high priority: the copies, load and result kernels of this chunk should take the next free SM ahead of the
        # (low-priority, persistent) solve kernels of the chunks in flight on the other streams
        self.stream = torch.cuda.Stream(device=dev, priority=-8)     # mapped to the highest the device has
        self.solvers = {}
        # two sets of input staging buffers: the rasters of this stream's NEXT chunk are copied in (on a copy stream of
        # its own) while the current chunk computes, so the host link never makes a chunk wait
        self.copy_stream = torch.cuda.Stream(device=dev, priority=-8) if stage_in else None
        self.t_bufs = [torch.empty((chunk, H, W, B), dtype=raster_dtype, device=dev) if stage_in else None for _ in range(2)]
        self.f_bufs = [torch.empty((chunk, H, W, B), dtype=raster_dtype, device=dev) if stage_in else None for _ in range(2)]
        self.m_bufs = [None if mask_dtype is None else torch.empty((chunk, H, W), dtype=mask_dtype, device=dev) for _ in range(2)]
        self.copied = [None, None]          # event: staging set j holds its chunk
        self.consumed = [None, None]        # event: the load of the chunk that used staging set j has read it
        self.uses = 0
        self.cloud = torch.empty((chunk, H, W), dtype=torch.uint8, device=dev) if need_cloud else None
        self.buf_scratch = torch.empty(3 * chunk * H * W, dtype=torch.uint8, device=dev) if buffer else None
        self.res = torch.empty(res_shape(chunk), dtype=res_dtype, device=dev) if need_res else None
        self.niter = torch.zeros(chunk * B, dtype=torch.int32, device=dev)
        self.fill_scratch = {}
        self.pending = None

    def close(self):
        for sv in self.solvers.values():
            sv.close()
        self.solvers = {}


class RawPipeline:
    """The streams, staging buffers and solvers of :func:`solve_raw`, kept across calls: a caller that cleans many
    stacks of one shape (a tile server, the benchmark) pays for workspace allocation and solver creation once."""

    def __init__(self, device=None, chunk_patches=None, streams=None, k_block=None):
        self.dev = require_cuda(device)
        self.k_block = _solve.DEFAULT_K_BLOCK if k_block is None else int(k_block)
        self.chunk = int(chunk_patches or PIPELINE_CHUNK_PATCHES)
        self.K = int(streams or RAW_STREAMS)
        self.slots = []
        self.key = None
        self.launches = 0

    def close(self):
        for sl in self.slots:
            self.launches += sum(sv.launches for sv in sl.solvers.values())
            sl.close()
        self.slots = []
        self.key = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def total_launches(self):
        """Kernels launched by the library through this pipeline so far."""
        return self.launches + sum(sv.launches for sl in self.slots for sv in sl.solvers.values())

    def _slots(self, key, make):
        if self.key != key:
            self.close()
            import torch
            with torch.cuda.device(self.dev):
                self.slots = [make() for _ in range(self.K)]
            self.key = key
        return self.slots

    def solve(self, target, features, mask=None, opts=None, *, mask_threshold=0.2, buffer_mask=False, mask_buffer_size=5,
              mask_buffer_passes=1, clip_target=None, clip_features=None, dtype=np.float32, out=None, out_layout="bhw",
              out_dtype=None, result="full", return_sweeps=False):
        """See :func:`solve_raw`."""
        import torch
        dev = self.dev
        opts = wp_run_options() if opts is None else opts
        if not _fused_ok(opts):
            raise VPintError("solve_raw serves run() with exact weights and without prioritise_identity / resistance / "
                             "auto_adapt; use VPint2_interpolator for those")
        if result not in ("full", "unknown"):
            raise VPintError("result must be 'full' or 'unknown'")
        as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        target, features = as_t(target), as_t(features)
        mask = None if mask is None else as_t(mask)
        if target.dim() != 4 or target.shape != features.shape:
            raise VPintError("target and features must be (N, H, W, B) stacks of one shape")
        if target.dtype != features.dtype:
            features = features.to(target.dtype)
        N_, H, W, B = target.shape
        if mask is not None and tuple(mask.shape) != (N_, H, W):
            raise VPintError("mask must have shape (N, H, W)")
        if mask is not None and mask.dtype not in (torch.uint8, torch.uint16, torch.float32, torch.float64):
            mask = mask.to(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.float64)
        if H * W < 128:
            raise VPintError("solve_raw needs grids of at least 128 cells (smaller ones: VPint2_interpolator / solve_bands)")
        value_dtype = "float64" if np.dtype(dtype) == np.float64 else "float32"
        chunk = self.chunk
        if out_layout == "bhw":
            res_shape, res_dtype = (lambda m: (m, B, H, W)), torch.float64
        else:
            tdt = out_dtype if isinstance(out_dtype, torch.dtype) else {
                np.dtype(np.float32): torch.float32, np.dtype(np.uint16): torch.uint16}.get(
                np.dtype(out_dtype if out_dtype is not None else dtype), torch.float64)
            res_shape, res_dtype = (lambda m: (m, H, W, B)), tdt
        in_on_dev = target.device.type != "cpu"
        if out is None:
            if result == "unknown":
                raise VPintError("result='unknown' writes into the caller's array: pass out")
            out = torch.empty(res_shape(N_), dtype=res_dtype, device=dev) if in_on_dev else \
                torch.empty(res_shape(N_), dtype=res_dtype, pin_memory=True)
        if tuple(out.shape) != res_shape(N_) or out.dtype != res_dtype:
            raise VPintError("out must have shape %s and dtype %s" % (res_shape(N_), res_dtype))
        out_on_dev = out.device.type != "cpu"
        if result == "unknown" and not out_on_dev and not out.is_pinned():
            raise VPintError("result='unknown' needs a CUDA or pinned host `out`")
        need_res = (result == "full") and not out_on_dev
        # Chunks of `chunk` patches -- after a ramp: with the rasters on the host the first chunk of every stream waits
        # for its copies (one copy engine serves all streams), so the first two rounds use quarter / half chunks and the
        # SMs get work a few milliseconds sooner.
        sizes = []
        if not in_on_dev and chunk >= 4 and N_ >= 8 * chunk and os.environ.get("VPINT_PIPELINE_RAMP", "1") != "0":
            sizes = [chunk // 4] * self.K + [chunk // 2] * self.K
        spans, a0 = [], 0
        for m0 in sizes:
            spans.append((a0, a0 + m0))
            a0 += m0
        spans += [(a, min(a + chunk, N_)) for a in range(a0, N_, chunk)]
        sweeps_host = torch.empty(N_ * B, dtype=torch.int32, pin_memory=True)
        need_cloud = mask is not None or buffer_mask
        mask_stage = None if (mask is None or mask.device.type != "cpu") else mask.dtype
        main = torch.cuda.current_stream(dev)
        run_kw = dict(auto_terminate=opts["auto_terminate"], threshold=opts["threshold"], clip_val=opts["clip_val"])
        as_prob = (lambda t: t) if out_layout == "bhw" else (lambda t: t.permute(0, 3, 1, 2))
        key = (H, W, B, target.dtype, mask_stage, out_layout, res_dtype, need_res, need_cloud, bool(buffer_mask), in_on_dev)
        slots = self._slots(key, lambda: _RawSlot(dev, chunk, H, W, B, target.dtype, mask_stage, res_shape, res_dtype,
                                                  need_res, need_cloud, buffer_mask, stage_in=not in_on_dev))
        K = max(1, min(self.K, len(spans)))

        def emit_result(sl, a, b, solver):
            m = b - a
            if need_res:
                solver.result_unknown(as_prob(sl.res[:m]), niter=sl.niter[:m * B])
                out[a:b].copy_(sl.res[:m], non_blocking=True)
            else:
                solver.result_unknown(as_prob(out[a:b]), niter=sl.niter[:m * B])
            sweeps_host[a * B:b * B].copy_(sl.niter[:m * B], non_blocking=True)

        def finish(sl):
            if sl.pending is None:
                return
            a, b, solver = sl.pending
            sl.pending = None
            with torch.cuda.stream(sl.stream):
                if solver.run_finish():            # the resident kernel declined: the tiled path finished the run
                    if result == "full":           # every cell again (the known ones were written by the load)
                        solver.result(out=as_prob(sl.res[:b - a] if need_res else out[a:b]))
                    emit_result(sl, a, b, solver)

        try:
            with torch.cuda.device(dev):
                for sl in slots[:K]:
                    sl.stream.wait_stream(main)
                    if sl.copy_stream is not None:
                        sl.copy_stream.wait_stream(main)
                for i, (a, b) in enumerate(spans):
                    sl = slots[i % K]
                    finish(sl)                             # the chunk this slot carried K chunks ago
                    m = b - a
                    j = sl.uses & 1
                    sl.uses += 1

                    def stage(span, jj, on):
                        """Enqueue the host -> device copies of chunk `span` into staging set jj on stream `on`."""
                        a2, b2 = span
                        m2 = b2 - a2
                        with torch.cuda.stream(on):
                            if sl.consumed[jj] is not None:
                                on.wait_event(sl.consumed[jj])
                            sl.t_bufs[jj][:m2].copy_(target[a2:b2], non_blocking=True)
                            sl.f_bufs[jj][:m2].copy_(features[a2:b2], non_blocking=True)
                            if mask_stage is not None:
                                sl.m_bufs[jj][:m2].copy_(mask[a2:b2], non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(on)
                        sl.copied[jj] = ev

                    if not in_on_dev and sl.copied[j] is None:
                        stage((a, b), j, sl.stream)                # the first chunk of this stream: nothing was prefetched
                    with torch.cuda.stream(sl.stream):
                        if in_on_dev:
                            t_c, f_c = target[a:b], features[a:b]
                        else:
                            sl.stream.wait_event(sl.copied[j])
                            sl.copied[j] = None
                            t_c, f_c = sl.t_bufs[j][:m], sl.f_bufs[j][:m]
                        cloud = None
                        if mask is not None:
                            m_c = sl.m_bufs[j][:m] if mask_stage is not None else mask[a:b]
                            cloud = cloud_mask(m_c, mask_threshold, out=sl.cloud[:m])
                        if buffer_mask:
                            cloud = buffer_clouds(t_c, cloud, mask_buffer_size, mask_buffer_passes, scratch=sl.buf_scratch,
                                                  out=sl.cloud[:m])
                        if m not in sl.solvers:
                            sl.solvers[m] = Solver(2, m * B, H, W, nprob_inner=B, planes=True, k_block=self.k_block, device=dev,
                                                   blocking_wait=True)
                            sl.fill_scratch[m] = sl.solvers[m].fill_scratch(value_dtype)
                        solver = sl.solvers[m]
                        known_dst = None
                        if result == "full":
                            known_dst = as_prob(sl.res[:m] if need_res else out[a:b])
                        solver.ingest_bands(t_c, f_c, cloud, value_dtype=value_dtype, clip_target=clip_target,
                                            clip_features=clip_features, out=known_dst, scratch=sl.fill_scratch[m])
                        if not in_on_dev:
                            sl.consumed[j] = torch.cuda.Event()
                            sl.consumed[j].record(sl.stream)       # the staged rasters have been read (buffer_clouds reads them too)
                        solver.run_async(opts["iterations"], **run_kw)
                        emit_result(sl, a, b, solver)
                    sl.pending = (a, b, solver)
                    if not in_on_dev and i + K < len(spans):
                        stage(spans[i + K], j ^ 1, sl.copy_stream)     # this stream's next chunk, while this one computes
                for sl in slots[:K]:
                    finish(sl)
                for sl in slots[:K]:
                    sl.stream.synchronize()
                    main.wait_stream(sl.stream)
        finally:
            # whatever happened, the next call starts from a clean pipeline (an error leaves chunks in flight: drain them)
            for sl in slots[:K]:
                if sl.pending is not None:
                    try:
                        sl.stream.synchronize()
                        if sl.copy_stream is not None:
                            sl.copy_stream.synchronize()
                    except Exception:       # noqa: BLE001 -- a sticky CUDA error: nothing more to drain
                        pass
                    sl.pending = None
                sl.copied = [None, None]
                sl.consumed = [None, None]
                sl.uses = 0
        sweeps = sweeps_host.to(torch.int64).view(N_, B)
        return (out, sweeps) if return_sweeps else out


def solve_raw(target, features, mask=None, opts=None, *, device=None, chunk_patches=None, streams=None, k_block=None, **kw):
    """VPint2 cloud removal of a stack of patches straight from the RAW rasters: the interpolator's constructor
    (VPint/VPint2.py:13-55), every band's mean init (VPint/MRP.py:73-77) and run() (VPint/VPint2.py:58-119) on the
    device, chunk by chunk on ``streams`` CUDA streams driven by ONE host thread -- no step blocks the host, so the
    host->device copy of a chunk, the solve of another and the result of a third overlap.

    target, features : (N, H, W, B) uint16 / float32 / float64, NumPy or torch, host (pinned for asynchronous copies)
                       or CUDA.  NaN in a float target = unknown.
    mask             : (N, H, W) cloud mask (any real dtype) -> unknown on every band where ``mask > mask_threshold``.
    buffer_mask, mask_buffer_size, mask_buffer_passes, clip_target, clip_features : as the VPint2 constructor.
    dtype            : the interpolator's DTYPE (np.float32 = VPint2's default): values, clip and mean init live in it.
    result           : 'full'    -> every cell of the result (float64 (N, B, H, W) for out_layout 'bhw' as run()
                                    assembles it, ``out_dtype`` (N, H, W, B) for 'hwb' as run_serial() returns it);
                                    ``out_dtype=np.uint16`` ('hwb' only) is the write-back of a product: the result
                                    clipped to [0, 65535] and truncated like ``astype(np.uint16)``;
                       'unknown' -> ONLY the cloud cells are written, into ``out`` (required; CUDA or pinned host
                                    memory, written in place through its device mapping): the known cells are the
                                    caller's own data and never cross the host link.
    Returns out (and the sweep counts (N, B) with ``return_sweeps``) as torch tensors.  :class:`RawPipeline` is the
    same thing with its buffers kept across calls."""
    with RawPipeline(device, chunk_patches, streams, k_block) as pipe:
        return pipe.solve(target, features, mask, opts, **kw)


class VPint2_interpolator():
    """Reference: VPint/VPint2.py:12-150.  Constructor semantics are unchanged (dtype cast, optional axis move,
    clipping, cloud mask -> NaN on every band, optional cloud buffering).

    The constructor does not run them on the host, though: it keeps (copies of) the raw rasters, and ``run()`` /
    ``run_serial()`` hand those to the device, where the same steps are kernels (:func:`solve_raw`: a uint16 scene
    crosses the host link as uint16, not as two float32 copies).  ``target`` / ``features`` stay available as
    attributes -- the first access builds them on the host exactly as the reference does, assigning to them switches
    the object to the host-built arrays for good."""

    _LAZY_DTYPES = (np.dtype(np.uint16), np.dtype(np.float32), np.dtype(np.float64))

    def __init__(self, target, features, mask=None, buffer_mask=False, mask_buffer_size=5, mask_buffer_passes=1,
                 bands_first=False, clip_target=None, clip_features=None, dtype=None, **kwargs):
        assert target.shape == features.shape
        assert len(target.shape) == 3
        assert len(features.shape) == 3

        self.DTYPE = dtype if dtype is not None else np.float32
        if mask is not None:
            shp = np.moveaxis(target, 0, -1).shape if bands_first else target.shape
            assert len(mask.shape) == 2
            assert mask.shape[0] == shp[0]
            assert mask.shape[1] == shp[1]
        self._ctor = dict(mask=None if mask is None else np.array(mask, copy=True), buffer_mask=buffer_mask,
                          mask_buffer_size=mask_buffer_size, mask_buffer_passes=mask_buffer_passes,
                          clip_target=clip_target, clip_features=clip_features, kwargs=dict(kwargs))
        self._target = self._features = None
        lazy = (np.dtype(self.DTYPE) in (np.dtype(np.float32), np.dtype(np.float64))
                and isinstance(target, np.ndarray) and isinstance(features, np.ndarray)
                and target.dtype in self._LAZY_DTYPES and features.dtype == target.dtype
                and set(kwargs) <= {"threshold"})
        if bands_first:
            target = np.moveaxis(target, 0, -1)
            features = np.moveaxis(features, 0, -1)
        if lazy:
            self._raw = (np.array(target, copy=True, order="C"), np.array(features, copy=True, order="C"))
        else:
            self._raw = None
            self._build_host(target, features)

    # -- host construction (the reference's own steps, vectorised) -------------------------------------------------
    def _build_host(self, target, features):
        c = self._ctor
        target = target.astype(self.DTYPE)
        features = features.astype(self.DTYPE)
        # clip_* : int / float = upper bound only, otherwise (min, max)   (VPint/VPint2.py:30-41)
        clip_target, clip_features = c["clip_target"], c["clip_features"]
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
        if c["mask"] is not None:
            tgt = self.apply_mask(target, c["mask"], **c["kwargs"]).astype(self.DTYPE)
        else:
            tgt = target.astype(self.DTYPE)
        if c["buffer_mask"]:
            tgt = self.buffer_clouds(tgt, buffer_size=c["mask_buffer_size"], passes=c["mask_buffer_passes"])
        self._target = tgt
        self._features = features.astype(self.DTYPE)

    def _materialise(self):
        if self._target is None:
            self._build_host(*self._raw)

    @property
    def target(self):
        self._materialise()
        return self._target

    @target.setter
    def target(self, value):
        self._materialise()
        self._raw = None
        self._target = value

    @property
    def features(self):
        self._materialise()
        return self._features

    @features.setter
    def features(self, value):
        self._materialise()
        self._raw = None
        self._features = value

    def _run_raw(self, opts, out_layout):
        """run() / run_serial() from the raw rasters on the device (None when the options need the host-built arrays)."""
        if self._raw is None or not _fused_ok(opts) or self._raw[0].shape[0] * self._raw[0].shape[1] < 128:
            return None
        c = self._ctor
        t, f = self._raw
        m = c["mask"]
        if m is not None and m.dtype not in (np.uint8, np.uint16, np.float32, np.float64):
            m = m.astype(np.float64)
        res = solve_raw(t[None], f[None], None if m is None else m[None], opts,
                        mask_threshold=c["kwargs"].get("threshold", 0.2), buffer_mask=c["buffer_mask"],
                        mask_buffer_size=c["mask_buffer_size"], mask_buffer_passes=c["mask_buffer_passes"],
                        clip_target=c["clip_target"], clip_features=c["clip_features"], dtype=self.DTYPE,
                        out_layout=out_layout, out_dtype=self.DTYPE, chunk_patches=1, streams=1)
        return res[0].numpy()

    def run(self, **kwargs):
        """All bands in one batched device solve (the reference forks one process per
        band, VPint/VPint2.py:58-101).  Returns float64 (H, W, B), a strided view of a
        (B, H, W) array exactly like the reference's swapaxes result."""
        opts = wp_run_options(**kwargs)
        if opts["auto_adapt"]:
            return self._run_bands_with_search(kwargs, forked=True).swapaxes(0, 2).swapaxes(0, 1)
        pred_bhw = self._run_raw(opts, "bhw")
        if pred_bhw is None:
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
        res = self._run_raw(opts, "hwb")
        if res is not None:
            return res
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
    res = _run_patches_raw(interpolators, opts, serial)
    if res is not None:
        return res
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


def _same(a, b):
    if isinstance(a, (tuple, list)) or isinstance(b, (tuple, list)):
        return isinstance(a, (tuple, list)) and isinstance(b, (tuple, list)) and list(a) == list(b)
    return a == b


def _run_patches_raw(interpolators, opts, serial):
    """run_patches straight from the raw rasters when every interpolator still holds them and was built with the same
    constructor options (None otherwise: the caller then goes through the host-built arrays)."""
    first = interpolators[0]
    if _solve.DEFAULT_K_BLOCK != 0 or not _fused_ok(opts):
        return None
    for vp in interpolators:
        if getattr(vp, "_raw", None) is None or vp._raw[0].shape != first._raw[0].shape or vp._raw[0].dtype != first._raw[0].dtype:
            return None
        if np.dtype(vp.DTYPE) != np.dtype(first.DTYPE) or (vp._ctor["mask"] is None) != (first._ctor["mask"] is None):
            return None
        for k in ("buffer_mask", "mask_buffer_size", "mask_buffer_passes", "clip_target", "clip_features", "kwargs"):
            if not _same(vp._ctor[k], first._ctor[k]):
                return None
    H, W, _ = first._raw[0].shape
    if H * W < 128:
        return None
    c = first._ctor
    tgt = np.stack([vp._raw[0] for vp in interpolators])
    fts = np.stack([vp._raw[1] for vp in interpolators])
    mask = None
    if c["mask"] is not None:
        mask = np.stack([vp._ctor["mask"] for vp in interpolators])
        if mask.dtype not in (np.uint8, np.uint16, np.float32, np.float64):
            mask = mask.astype(np.float64)
    res = solve_raw(tgt, fts, mask, opts, mask_threshold=c["kwargs"].get("threshold", 0.2), buffer_mask=c["buffer_mask"],
                    mask_buffer_size=c["mask_buffer_size"], mask_buffer_passes=c["mask_buffer_passes"],
                    clip_target=c["clip_target"], clip_features=c["clip_features"], dtype=first.DTYPE,
                    out_layout="hwb" if serial else "bhw", out_dtype=first.DTYPE).numpy()
    if serial:
        return [res[i] for i in range(len(interpolators))]
    return [res[i].swapaxes(0, 2).swapaxes(0, 1) for i in range(len(interpolators))]
