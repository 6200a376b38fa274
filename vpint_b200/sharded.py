"""Row-sharded solve of ONE large scene across ranks (BASELINE config 4).

A single grid is a nearest-neighbour Jacobi stencil, so it shards by contiguous row
bands: rank r owns rows [r0, r1) of EVERY band (problems stay aligned across ranks,
so the per-band stopping rule stays global) and keeps ``halo`` extra rows on each
interior side.  One launch runs ``k <= halo`` sweeps on the local rows (halo cells are
recomputed redundantly and are valid ``halo - j`` rows deep after sweep j), then

    all-reduce(SUM) of the per-sweep sums  [P][k][4] fp64   -> identical stop decision on every rank
    send/recv of the k boundary rows of the new state        -> fresh halos for the next launch

The reference has no counterpart (its only parallelism is one forked process per band,
VPint/VPint2.py:63-99).  The driver below is written against a small engine protocol so
that the same loop runs over NCCL with the CUDA solver and over gloo with a CPU test
double (tests/test_sharded_gloo.py).
"""

from __future__ import annotations

import numpy as np

from .errors import VPintError


class RowShard:
    """Row-band partition of a grid of ``H`` rows over ``world`` ranks.

    Attributes (all in rows): ``r0, r1`` owned global rows; ``lo, hi`` local slice
    incl. halos (clipped at the physical border); ``own_row0`` / ``own_rows`` the owned
    band inside the local slice; ``row_offset`` (= ``lo``) global row of local row 0.
    """

    def __init__(self, H, world, rank, halo):
        if world < 1 or not (0 <= rank < world):
            raise VPintError("bad rank / world size")
        if H < world:
            raise VPintError("fewer rows than ranks")
        base, rem = divmod(int(H), int(world))
        self.H, self.world, self.rank, self.halo = int(H), int(world), int(rank), int(halo)
        self.r0 = rank * base + min(rank, rem)
        self.r1 = self.r0 + base + (1 if rank < rem else 0)
        self.lo = max(0, self.r0 - halo)
        self.hi = min(self.H, self.r1 + halo)
        self.own_row0 = self.r0 - self.lo
        self.own_rows = self.r1 - self.r0
        self.row_offset = self.lo
        self.local_rows = self.hi - self.lo
        self.up = rank - 1 if rank > 0 else None
        self.down = rank + 1 if rank < world - 1 else None
        if self.own_rows < halo and world > 1:
            raise VPintError("owned band (%d rows) thinner than the halo (%d)" % (self.own_rows, halo))

    def desc(self):
        """(global_H, row_offset, own_row0, own_rows) for ``vpint_desc``."""
        return (self.H, self.row_offset, self.own_row0, self.own_rows)

    def local(self, full):
        """Rows of a full (H, ...) array this rank must hold (owned + halos)."""
        return full[self.lo:self.hi]

    def owned(self, local):
        return local[self.own_row0:self.own_row0 + self.own_rows]


class Comm:
    """The two collectives the path needs, on a torch.distributed group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:                       # single process: nothing to exchange
            self.rank, self.world = 0, 1

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def exchange(self, shard, send_up, send_down, recv_up, recv_down):
        """Boundary rows to both row-neighbours, their boundary rows back (one grouped send/recv)."""
        ops = []
        P2P = self.dist.P2POp
        g = self.group
        peer = (lambda r: r) if g is None else (lambda r: self.dist.get_global_rank(g, r))
        if shard.up is not None:
            ops.append(P2P(self.dist.isend, send_up, peer(shard.up), group=g))
            ops.append(P2P(self.dist.irecv, recv_up, peer(shard.up), group=g))
        if shard.down is not None:
            ops.append(P2P(self.dist.isend, send_down, peer(shard.down), group=g))
            ops.append(P2P(self.dist.irecv, recv_down, peer(shard.down), group=g))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()


def run_sharded(engine, shard, comm, max_iter, poll_every=8, use_graph=None):
    """Drive one run() of a row-sharded batch to the reference's stopping sweeps.

    ``engine`` protocol (CUDA: :class:`SolverEngine`; tests: a NumPy double):
      known() -> tensor             known-cell sums of the owned rows (all-reduced once here)
      begin(max_iter)               start of run()
      step_begin()                  k sweeps on the local rows + local per-sweep sums
      sums() -> tensor              [P][k][4] local sums (all-reduced in place here)
      step_end()                    stop decision from the reduced sums
      halo_send() -> (up, down)     boundary rows of the current state, packed
      halo_recv() -> (up, down)     buffers to receive into
      halo_apply()                  write the received rows into the halo
      active() -> int               problems still running (host sync)
    ``use_graph`` (CUDA engines only; default from the environment variable VPINT_SHARD_GRAPH): capture one
    launch group -- kernels, the NCCL all-reduce and the halo send/recv -- into a CUDA graph after the
    first eager launch and replay it, so that the host cost per launch is one graph launch instead of a
    dozen enqueues (at 8 GPUs a launch group is only ~0.2 ms of stencil work).
    Returns the number of launches issued.
    """
    comm.all_reduce_sum(engine.known())
    engine.begin(max_iter)
    if max_iter <= 0:
        return 0

    def launch_group():
        engine.step_begin()
        comm.all_reduce_sum(engine.sums())
        engine.step_end()
        if shard.world > 1:
            su, sd = engine.halo_send()
            ru, rd = engine.halo_recv()
            comm.exchange(shard, su, sd, ru, rd)
            engine.halo_apply()

    if use_graph is None:
        import os
        use_graph = os.environ.get("VPINT_SHARD_GRAPH", "0") == "1"
    use_graph = bool(use_graph) and getattr(engine, "graph_capable", False)
    launches = 0
    budget = 2 * int(max_iter) + 2
    graph = None
    while launches < budget:
        for _ in range(poll_every):
            if graph is not None:
                graph.replay()
            else:
                launch_group()
                if use_graph and launches == 0:      # the first launch ran eagerly (one-off repairs, NCCL warm-up)
                    graph = engine.capture(launch_group)
            launches += 1
        if engine.active() <= 0:
            return launches
    raise VPintError("internal error: sharded loop did not terminate within its launch budget")


def run_lockstep(engines, shards, max_iter, poll_every=8):
    """The same protocol with every shard in THIS process (one device): the all-reduce and the halo
    exchange become tensor copies.  Used to check the row-sharded path on a single GPU and by hosts
    that split a scene without owning several devices."""
    n = len(engines)
    total = engines[0].known().clone()
    for e in engines[1:]:
        total += e.known()
    for e in engines:
        e.known().copy_(total)
        e.begin(max_iter)
    if max_iter <= 0:
        return 0
    launches = 0
    budget = 2 * int(max_iter) + 2
    while launches < budget:
        for _ in range(poll_every):
            for e in engines:
                e.step_begin()
            total = engines[0].sums().clone()
            for e in engines[1:]:
                total += e.sums()
            for e in engines:
                e.sums().copy_(total)
                e.step_end()
            if n > 1:
                sent = [e.halo_send() for e in engines]
                for i, e in enumerate(engines):
                    ru, rd = e.halo_recv()
                    if shards[i].up is not None:
                        ru.copy_(sent[i - 1][1])
                    if shards[i].down is not None:
                        rd.copy_(sent[i + 1][0])
                    e.halo_apply()
            launches += 1
        if max(e.active() for e in engines) <= 0:
            return launches
    raise VPintError("internal error: sharded loop did not terminate within its launch budget")


class SolverEngine:
    """Engine protocol over the CUDA :class:`~vpint_b200.engine.Solver` (NCCL path)."""

    def __init__(self, solver, shard, run_kwargs):
        self.s = solver
        self.shard = shard
        self.kw = dict(run_kwargs)
        torch = solver.torch
        k = solver.k_block
        pitch = solver.state_tensor(0).shape[-1]
        dt = torch.float64 if solver.storage == "float64" else torch.float32
        mk = lambda: torch.empty((solver.nprob, k, pitch), dtype=dt, device=solver.device)
        self.rows = k
        self.send_up = mk() if shard.up is not None else None
        self.recv_up = mk() if shard.up is not None else None
        self.send_down = mk() if shard.down is not None else None
        self.recv_down = mk() if shard.down is not None else None

    graph_capable = True

    def capture(self, fn):
        """Capture ``fn`` (one launch group) into a CUDA graph on a side stream (per run: replaying a graph
        with NCCL nodes across runs of a re-loaded solver hung on 2 x B200 and is not done)."""
        torch = self.s.torch
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.s.device)
        side.wait_stream(torch.cuda.current_stream(self.s.device))
        with torch.cuda.graph(g, stream=side):
            fn()
        torch.cuda.current_stream(self.s.device).wait_stream(side)
        return g

    def known(self):
        return self.s.known_tensor()

    def begin(self, max_iter):
        if self.s.any_dirty:
            raise VPintError("a row-sharded run needs a start state that equals the original on known cells")
        self.s.begin_run(max_iter, **self.kw)

    def step_begin(self):
        self.s.step_begin()

    def sums(self):
        return self.s.sums_tensor()

    def step_end(self):
        self.s.step_end()

    def halo_send(self):
        self.s.halo_pack(self.rows, self.send_up, self.send_down)
        return self.send_up, self.send_down

    def halo_recv(self):
        return self.recv_up, self.recv_down

    def halo_apply(self):
        self.s.halo_unpack(self.rows, self.recv_up, self.recv_down)

    def active(self):
        return self.s.active_count()


def open_shard(target_local, features_local, shard, fill, opts, *, k_block=4, storage="float64", device=None):
    """Stage one rank's rows of a VPint2 scene on the device: returns (solver, engine) ready for
    :func:`run_sharded` / :func:`run_lockstep`.

    target_local / features_local : torch tensors or NumPy arrays (local_rows, W, B) = ``shard.local(full)``;
                                    NaN in the target = cloud (``VPint2_interpolator.target``).
    fill : per-band start value of the unknown cells, ``np.nanmean`` of the WHOLE band in the band's
           dtype exactly as ``MRP.init_pred_grid`` takes it (VPint/MRP.py:73-77) -- a global statistic,
           so the caller that holds (or streams) the whole scene supplies it.
    """
    from .engine import Solver, require_cuda, to_device
    import torch
    dev = require_cuda(device)
    t_dev = target_local.to(dev) if isinstance(target_local, torch.Tensor) else to_device(np.ascontiguousarray(target_local), dev)
    f_dev = features_local.to(dev) if isinstance(features_local, torch.Tensor) else to_device(np.ascontiguousarray(features_local), dev)
    Hl, W, B = t_dev.shape
    if Hl != shard.local_rows:
        raise VPintError("target_local has %d rows, the shard expects %d" % (Hl, shard.local_rows))
    kb = max(1, int(k_block))
    if shard.world > 1 and shard.halo < kb:
        raise VPintError("halo (%d) must be at least k_block (%d)" % (shard.halo, kb))
    solver = Solver(2, B, Hl, W, nprob_inner=B, storage=storage, planes=True, k_block=kb, shard=shard.desc(), device=dev)
    try:
        stage_shard(solver, t_dev, f_dev, fill, opts)
        eng = SolverEngine(solver, shard, dict(auto_terminate=opts["auto_terminate"], threshold=opts["threshold"],
                                               clip_val=opts["clip_val"]))
    except Exception:
        solver.close()
        raise
    return solver, eng


def stage_shard(solver, t_dev, f_dev, fill, opts):
    """(Re)load one rank's rows into an existing shard solver: start state + weights."""
    solver.load(t_dev.permute(2, 0, 1).unsqueeze(0), fill=np.asarray(fill, dtype=np.float64).reshape(-1))
    solver.weights(f_dev.permute(2, 0, 1).unsqueeze(0).unsqueeze(-1), method=opts["method"],
                   clamp=(opts["method"] != "exact"), min_gamma=0.0, max_gamma=float("inf"))


def open_shard_raw(target_local, features_local, cloud_local, shard, comm, opts, *, value_dtype="float32", k_block=1,
                   device=None):
    """:func:`open_shard` from the RAW rasters of this rank's rows (uint16 / float32 / float64 CUDA tensors
    (local_rows, W, B) = ``shard.local(full)``) and the all-band cloud mask (uint8 (local_rows, W), nonzero = cloud,
    or None): no host pass at all -- the per-band mean init is formed on the device, every rank summing the leaves of
    NumPy's summation tree that start in its rows and all ranks all-reducing the leaf sums (VPint/MRP.py:73-77 bit
    for bit for the WHOLE band).  Needs >= 1 halo row (W >= 127) on interior sides, which ``halo >= k_block`` gives."""
    from .engine import Solver, require_cuda
    dev = require_cuda(device)
    Hl, W, B = target_local.shape
    if Hl != shard.local_rows:
        raise VPintError("target_local has %d rows, the shard expects %d" % (Hl, shard.local_rows))
    kb = max(1, int(k_block))
    if shard.world > 1 and shard.halo < kb:
        raise VPintError("halo (%d) must be at least k_block (%d)" % (shard.halo, kb))
    solver = Solver(2, B, Hl, W, nprob_inner=B, planes=True, k_block=kb, shard=shard.desc(), device=dev)
    try:
        scratch = solver.fill_scratch(value_dtype)
        stage_shard_raw(solver, target_local, features_local, cloud_local, comm, scratch, value_dtype)
        eng = SolverEngine(solver, shard, dict(auto_terminate=opts["auto_terminate"], threshold=opts["threshold"],
                                               clip_val=opts["clip_val"]))
        eng.fill_scratch = scratch
    except Exception:
        solver.close()
        raise
    return solver, eng


def stage_shard_raw(solver, t_dev, f_dev, cloud_dev, comm, scratch, value_dtype="float32"):
    """(Re)load one rank's rows from the raw rasters: device mean init (leaf sums all-reduced over the ranks), then the
    fused load (state, unknown-cell mask, feature plane, known-cell sums)."""
    t4, f4 = t_dev.unsqueeze(0), f_dev.unsqueeze(0)
    c3 = None if cloud_dev is None else cloud_dev.unsqueeze(0)
    solver.fill_leaves(t4, c3, scratch, value_dtype=value_dtype)
    leaves, counts = solver.fill_views(scratch, value_dtype)
    comm.all_reduce_sum(leaves)
    comm.all_reduce_sum(counts)
    solver.fill_combine(scratch, value_dtype)
    solver.ingest_bands(t4, f4, c3, value_dtype=value_dtype, fill_ready=True)


def shard_result(solver, shard):
    """(pred float64 torch tensor (B, own_rows, W) on the device, sweeps int64 ndarray (B,))."""
    out, niter = solver.result()
    pred = out[:, shard.own_row0:shard.own_row0 + shard.own_rows, :].clone()
    return pred, niter.cpu().numpy().astype(np.int64)


def solve_scene_sharded(target_local, features_local, shard, fill, opts, *, group=None, k_block=4,
                        storage="float64", device=None, poll_every=8):
    """VPint2 cloud removal of one scene, row-sharded over the ranks of ``group`` (every band a problem,
    all bands on every rank).  Arguments as :func:`open_shard`; returns :func:`shard_result`."""
    solver, eng = open_shard(target_local, features_local, shard, fill, opts, k_block=k_block, storage=storage,
                             device=device)
    try:
        run_sharded(eng, shard, Comm(group), opts["iterations"], poll_every=poll_every)
        return shard_result(solver, shard)
    finally:
        solver.close()
