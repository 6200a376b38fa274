#!/usr/bin/env python
"""BASELINE config 4: VPint2 cloud removal of ONE full Sentinel-2-shaped scene (default 10980 x 10980 x 13),
row-sharded over the ranks with halo exchange + all-reduce of the per-sweep sums (vpint_b200/sharded.py).

  python bench_scene.py --size 10980                                   # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_scene.py

`bench.py` calls :func:`scene_record` under its own torchrun and puts the result into its JSON line as `scene`:

  parity : a reduced scene solved row-sharded over the N ranks (N = 1: two shards in lockstep on the one device) AND
           unsharded on rank 0, compared per pixel (<= 1e-9 relative) and per band (equal stopping sweeps)
  c4     : the full scene -- every rank generates only its own rows (+ halos) of the synthetic product on its device
           as uint16 rasters + a cloud mask; a step = the NumPy-exact per-band mean init on the device (leaf sums
           all-reduced), fused load, Jacobi sweeps to every band's own stopping sweep, result of the owned rows.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def scene_rows(lo, hi, Hs, Ws, B, seed, dev, n_discs, rad_scale=1.0):
    """Rows [lo, hi) of the synthetic scene as a product holds it: (target, features) uint16 (rows, Ws, B) -- cloudy
    pixels keep a measured value -- and the cloud mask uint8 (rows, Ws), 1 = cloud."""
    import torch
    rng = np.random.default_rng(seed)
    modes = rng.uniform(0, 1, (B, 6, 5))
    discs = rng.uniform(0, 1, (n_discs, 3))
    yy = (torch.arange(lo, hi, device=dev, dtype=torch.float64) / Hs).view(-1, 1)
    xx = (torch.arange(Ws, device=dev, dtype=torch.float64) / Ws).view(1, -1)
    rows = hi - lo
    target = torch.empty((rows, Ws, B), dtype=torch.uint16, device=dev)
    feats = torch.empty((rows, Ws, B), dtype=torch.uint16, device=dev)
    cloud = torch.zeros((rows, Ws), dtype=torch.bool, device=dev)
    for cy, cx, r in discs:
        rad = rad_scale * (100.0 + 500.0 * r) / 10980.0      # relative to the scene: the cloud fraction does not depend on its size
        cloud |= (yy - cy) ** 2 + (xx - cx) ** 2 < rad * rad
    BLK = 256            # noise is generated per fixed block of rows so that it does not depend on the sharding
    for b in range(B):
        acc = torch.zeros((rows, Ws), dtype=torch.float64, device=dev)
        for k in range(6):
            fy, fx, py, px, amp = modes[b, k]
            acc += (0.5 + 0.5 * amp) * torch.sin(2 * np.pi * ((0.3 + 4.0 * fy) * yy + py)) * torch.cos(2 * np.pi * ((0.3 + 4.0 * fx) * xx + px))
        base = 1500.0 + 1200.0 * acc / 4.5
        for blk in range(lo // BLK, (hi - 1) // BLK + 1):
            g = torch.Generator(device=dev)
            g.manual_seed(seed * 1000003 + b * 100003 + blk)
            nz = torch.randn((2, BLK, Ws), generator=g, device=dev, dtype=torch.float32)
            a, e = max(lo, blk * BLK), min(hi, (blk + 1) * BLK)
            sl = slice(a - lo, e - lo)
            nb = slice(a - blk * BLK, e - blk * BLK)
            target[sl, :, b] = torch.round(base[sl] + 20.0 * nz[0, nb]).clamp(1, 65535).to(torch.int32).to(torch.uint16)
            feats[sl, :, b] = torch.round(1.1 * base[sl] + 10.0 * nz[1, nb]).clamp(1, 65535).to(torch.int32).to(torch.uint16)
        del acc, base
    return target, feats, cloud.to(torch.uint8)


def scene_cloud_rows(Hs, Ws, seed, B, n_discs, rad_scale=1.0):
    """Cloudy pixels per row of the synthetic scene (what a caller reads off the cloud mask before it deals the rows
    out): the discs of :func:`scene_rows`, rasterised row by row on the host -- only their row extents are needed."""
    rng = np.random.default_rng(seed)
    rng.uniform(0, 1, (B, 6, 5))
    discs = rng.uniform(0, 1, (n_discs, 3))
    y = np.arange(Hs) / Hs
    x = np.arange(Ws) / Ws
    counts = np.zeros(Hs, dtype=np.int64)
    for y0 in range(0, Hs, 512):
        yy = y[y0:y0 + 512, None]
        cloud = np.zeros((yy.shape[0], Ws), dtype=bool)
        for cy, cx, r in discs:
            rad = rad_scale * (100.0 + 500.0 * r) / 10980.0
            if abs(cy - y[y0]) > rad + 512.0 / Hs and abs(cy - y[min(y0 + 511, Hs - 1)]) > rad + 512.0 / Hs:
                continue
            cloud |= (yy - cy) ** 2 + (x[None, :] - cx) ** 2 < rad * rad
        counts[y0:y0 + 512] = cloud.sum(axis=1)
    return counts


class _LocalComm:
    """Comm of a single process (nothing to exchange)."""
    world, rank = 1, 0

    def all_reduce_sum(self, t):
        pass


def _solve_rows(shard, comm, Hs, Ws, B, seed, dev, discs, opts, kb, steps=1, warmup=0, poll=8, rad_scale=1.0):
    """Generate this shard's rows, solve, return (pred (B, own_rows, W) float64, sweeps, launches, ms per step, cloud cells)."""
    import torch
    from vpint_b200.sharded import open_shard_raw, run_sharded, shard_result, stage_shard_raw
    target, feats, cloud = scene_rows(shard.lo, shard.hi, Hs, Ws, B, seed, dev, discs, rad_scale)
    solver, eng = open_shard_raw(target, feats, cloud, shard, comm, opts, k_block=kb)
    try:
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def step():
            marks[0].record()
            stage_shard_raw(solver, target, feats, cloud, comm, eng.fill_scratch)
            marks[1].record()
            launches = run_sharded(eng, shard, comm, opts["iterations"], poll_every=poll)
            marks[2].record()
            pred, sweeps = shard_result(solver, shard)
            marks[3].record()
            return pred, sweeps, launches
        for _ in range(warmup):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if comm.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(steps, 1)):
            pred, sweeps, launches = step()
        e1.record()
        if comm.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / max(steps, 1)
        own = slice(shard.own_row0, shard.own_row0 + shard.own_rows)
        cloud_cells = float(cloud[own].sum().item())
        phases = {"mean_init_and_load": marks[0].elapsed_time(marks[1]), "sweeps": marks[1].elapsed_time(marks[2]),
                  "result": marks[2].elapsed_time(marks[3])}
        tiles = (solver.active_tiles, solver.total_tiles, "peer stores + flag (symmetric memory)" if eng.peer is not None else
                 ("NCCL send/recv" if comm.world > 1 else "none"), getattr(eng, "peer_error", None), phases)
    finally:
        solver.close()
    return pred, sweeps, launches, ms, cloud_cells, tiles


def scene_record(world, rank, dev, size=10980, parity_size=2048, steps=2, bands=13, k_block=1, discs=60, seed=1004):
    """The `scene` record of bench.py's JSON line (rank 0 returns it, the others None)."""
    import torch
    import torch.distributed as dist
    from vpint_b200.sharded import Comm, RowShard, run_lockstep_lagged, open_shard_raw, shard_result
    from vpint_b200.wp import wp_run_options
    opts = wp_run_options()
    kb = max(1, k_block)
    comm = Comm() if world > 1 else _LocalComm()
    torch.cuda.empty_cache()
    rec = {"k_block": kb}

    # ---- parity: sharded == unsharded on a reduced scene --------------------------------------------------------
    Hp = Wp = parity_size
    pdiscs, pscale = 14, 3.0          # fewer, larger clouds than the full scene: tens of sweeps per band
    if world > 1:
        shard = RowShard(Hp, world, rank, halo=kb)
        pred, sweeps, _, _, _, ptiles = _solve_rows(shard, comm, Hp, Wp, bands, seed + 1, dev, pdiscs, opts, kb, rad_scale=pscale)
        if rank == 0:
            parts = [pred]
            for r in range(1, world):
                sh = RowShard(Hp, world, r, halo=kb)
                buf = torch.empty((bands, sh.own_rows, Wp), dtype=torch.float64, device=dev)
                dist.recv(buf, src=r)
                parts.append(buf)
            got = torch.cat(parts, dim=1)
        else:
            dist.send(pred.contiguous(), dst=0)
        how = "row-sharded over %d ranks (halo rows: %s) vs unsharded on rank 0" % (world, ptiles[2])
    else:
        # one device: two shards driven in lockstep (the all-reduce and the halo exchange become tensor copies)
        shards = [RowShard(Hp, 2, r, halo=kb) for r in range(2)]
        solvers, engines = [], []
        try:
            # the mean init of the whole band: both shards' leaf sums added before either combines
            from vpint_b200.engine import Solver
            raws = [scene_rows(sh.lo, sh.hi, Hp, Wp, bands, seed + 1, dev, pdiscs, pscale) for sh in shards]
            scr = []
            for sh, (t, f, c) in zip(shards, raws):
                s = Solver(2, bands, sh.local_rows, Wp, nprob_inner=bands, planes=True, k_block=kb, shard=sh.desc(), device=dev)
                solvers.append(s)
                sc = s.fill_scratch("float32")
                s.fill_leaves(t.unsqueeze(0), c.unsqueeze(0), sc)
                scr.append(sc)
            views = [s.fill_views(sc) for s, sc in zip(solvers, scr)]
            tot_l, tot_c = views[0][0] + views[1][0], views[0][1] + views[1][1]
            from vpint_b200.sharded import SolverEngine
            for s, sh, sc, (l, c), (t, f, cl) in zip(solvers, shards, scr, views, raws):
                l.copy_(tot_l); c.copy_(tot_c)
                s.fill_combine(sc)
                s.ingest_bands(t.unsqueeze(0), f.unsqueeze(0), cl.unsqueeze(0), fill_ready=True)
                engines.append(SolverEngine(s, sh, dict(auto_terminate=opts["auto_terminate"], threshold=opts["threshold"],
                                                        clip_val=opts["clip_val"])))
            run_lockstep_lagged(engines, shards, opts["iterations"])
            res = [shard_result(s, sh) for s, sh in zip(solvers, shards)]
            got = torch.cat([r[0] for r in res], dim=1)
            sweeps = res[0][1]
        finally:
            for s in solvers:
                s.close()
        del raws
        how = "two row shards in lockstep on the one device (lagged stop decision, halo rows pushed with a flag) vs unsharded"
    if rank == 0:
        full = RowShard(Hp, 1, 0, halo=0)
        ref, ref_sweeps, _, _, _, _ = _solve_rows(full, _LocalComm(), Hp, Wp, bands, seed + 1, dev, pdiscs, opts, kb, rad_scale=pscale)
        rel = float(((got - ref).abs() / ref.abs()).max().item())
        rec["parity"] = {"scene": "%dx%dx%d" % (Hp, Wp, bands), "how": how, "max_rel": rel,
                         "sweeps_sharded": [int(v) for v in sweeps], "sweeps_unsharded": [int(v) for v in ref_sweeps],
                         "tolerance": 1e-9}
        rec["parity_ok"] = bool(rel <= 1e-9 and list(sweeps) == list(ref_sweeps))
        del got, ref
    torch.cuda.empty_cache()

    # ---- the full scene ---------------------------------------------------------------------------------------
    if size > 0:
        bounds = None
        if world > 1 and os.environ.get("VPINT_EQUAL_ROWS", "0") != "1":
            # rows dealt by WORK: the unknown cells of a row (read off the cloud mask) + a little for its set-up cost
            from vpint_b200.sharded import balanced_bounds
            counts = scene_cloud_rows(size, size, seed, bands, discs)
            bounds = balanced_bounds(counts + 0.02 * size, world, min_rows=max(kb, 16))
        shard = RowShard(size, world, rank, halo=kb if world > 1 else 0, bounds=bounds)
        pred, sweeps, launches, ms, cloud_cells, tiles = _solve_rows(shard, comm, size, size, bands, seed, dev, discs, opts, kb,
                                                                     steps=steps, warmup=1)
        phases = {k: round(v, 3) for k, v in tiles[4].items()}
        ok = bool(torch.isfinite(pred).all().item())
        del pred
        t = torch.tensor([ms, cloud_cells, 1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t[:1].clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t[1:].clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, cloud_cells, ok = float(tm.item()), float(ts[0].item()), bool(ts[1].item() == world)
        if rank == 0:
            pxit = float(np.sum(sweeps)) * size * size
            frac = cloud_cells / (float(size) * size)
            rec["c4"] = {"workload": "C4: VPint2 cloud removal of one %dx%dx%d scene from uint16 rasters + cloud mask, "
                                     "row-sharded over %d GPU(s), auto-terminate 1e-4; a step = device mean init (NumPy-exact), "
                                     "fused load, sweeps, result" % (size, size, bands, world),
                         "ms": ms, "steps": steps, "value": pxit / (ms * 1e-3) / 1e6, "unit": "Mpx*it/s", "scaling": "strong",
                         "sweeps_per_band": [int(v) for v in sweeps], "launches": int(launches), "cloud_fraction": round(frac, 4),
                         "rows_rank0": shard.own_rows, "halo": shard.halo, "finite": ok,
                         "row_partition": "by work (unknown cells per row)" if bounds is not None else "equal rows",
                         "row_bounds": bounds, "phases_ms_rank0": phases,
                         "active_tiles_rank0": tiles[0], "total_tiles_rank0": tiles[1], "halo_rows": tiles[2],
                         "halo_fallback_reason": tiles[3],
                         "protocol": "stop decision lags one sweep (all-reduce overlaps the next sweep)" if os.environ.get("VPINT_SHARD_LAGGED", "1") != "0" else "all-reduce and decision every sweep",
                         "unknown_cell_gbps_48B": 48.0 * pxit * frac / (ms * 1e-3) / 1e9}
            # roofline of the sweep phase alone: the ratio stencil moves 8 B state in + 4 B feature (the float32 copy of the feature
            # plane, exact for these uint16 rasters; 8 B otherwise) + 8 B state out per unknown cell per sweep
            try:
                peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
                peak_src = "measured (MEASURED_PEAKS.json)"
            except Exception:
                peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
            if phases.get("sweeps"):
                ach = 20.0 * pxit * frac / world / (phases["sweeps"] * 1e-3) / 1e9
                rec["c4"]["roofline_sweeps"] = {"bound": "hbm", "kernel": "jacobi2d_ratio_bulk_kernel (per GPU, sweep phase of rank 0 incl. the "
                                                "reduction / commit / halo kernels of every sweep)", "bytes_per_unknown_cell_sweep": 20,
                                                "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak}
    torch.cuda.empty_cache()
    return rec if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=10980)
    ap.add_argument("--parity-size", type=int, default=2048)
    ap.add_argument("--bands", type=int, default=13)
    ap.add_argument("--k-block", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--discs", type=int, default=60)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = scene_record(world, rank, dev, size=args.size, parity_size=args.parity_size, steps=args.steps, bands=args.bands,
                       k_block=args.k_block, discs=args.discs)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
