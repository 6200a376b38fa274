#!/usr/bin/env python
"""BASELINE config 4: VPint2 cloud removal of ONE full Sentinel-2-shaped scene (default 10980 x 10980 x 13),
row-sharded over the ranks with halo exchange + all-reduce of the per-sweep sums (vpint_b200/sharded.py).

  python bench_scene.py --size 10980 --k-block 1                      # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_scene.py --k-block 4

Every rank generates only its own rows (+ halos) of the synthetic scene on its device; the per-band init
value (the reference's float32 nanmean of the whole band, VPint/MRP.py:73-77) is formed from all-reduced
fp64 sums.  A step = load + weights + Jacobi sweeps to every band's own stopping sweep + result gather.
Prints one JSON line (rank 0): whole-job Mpx*it/s, sweeps per band, HBM roofline of the stencil kernel.
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def scene_rows(lo, hi, Hs, Ws, B, seed, dev, cloud_frac_discs):
    """Rows [lo, hi) of the synthetic scene: (target, features) float32 (rows, Ws, B), NaN = cloud."""
    import torch
    rng = np.random.default_rng(seed)
    modes = rng.uniform(0, 1, (B, 6, 5))
    discs = rng.uniform(0, 1, (cloud_frac_discs, 3))
    yy = (torch.arange(lo, hi, device=dev, dtype=torch.float64) / Hs).view(-1, 1)
    xx = (torch.arange(Ws, device=dev, dtype=torch.float64) / Ws).view(1, -1)
    rows = hi - lo
    target = torch.empty((rows, Ws, B), dtype=torch.float32, device=dev)
    feats = torch.empty((rows, Ws, B), dtype=torch.float32, device=dev)
    cloud = torch.zeros((rows, Ws), dtype=torch.bool, device=dev)
    for cy, cx, r in discs:
        rad = (100.0 + 500.0 * r) / 10980.0
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
            target[sl, :, b] = torch.round(base[sl] + 20.0 * nz[0, nb]).clamp(1, 65535).to(torch.float32)
            feats[sl, :, b] = torch.round(1.1 * base[sl] + 10.0 * nz[1, nb]).clamp(1, 65535).to(torch.float32)
        del acc, base
    target[cloud] = float("nan")
    return target, feats, cloud


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=10980)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--bands", type=int, default=13)
    ap.add_argument("--k-block", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--discs", type=int, default=150)
    ap.add_argument("--iterations", type=int, default=-1)
    ap.add_argument("--poll", type=int, default=8)
    ap.add_argument("--graph", type=int, default=0, help="replay one captured launch group (kernels + NCCL) per launch")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from vpint_b200.sharded import Comm, RowShard, open_shard, run_sharded, shard_result, stage_shard
    from vpint_b200.wp import wp_run_options

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Hs, Ws, B = args.size, (args.width or args.size), args.bands
    kb = max(1, args.k_block)
    shard = RowShard(Hs, world, rank, halo=kb if world > 1 else 0)
    target, feats, cloud = scene_rows(shard.lo, shard.hi, Hs, Ws, B, 1004, dev, args.discs)
    # per-band float32 mean of the known cells of the WHOLE band (owned rows only, summed over ranks in fp64)
    own = slice(shard.own_row0, shard.own_row0 + shard.own_rows)
    known = ~cloud[own]
    sums = torch.stack([target[own, :, b][known].double().sum() for b in range(B)] + [known.sum().double()])
    if world > 1:
        dist.all_reduce(sums)
    fill = (sums[:B] / sums[B]).to(torch.float32).double().cpu().numpy()
    cloud_cells = cloud[own].sum().double()
    if world > 1:
        dist.all_reduce(cloud_cells)
    opts = wp_run_options(iterations=args.iterations)
    comm = Comm()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solver, eng = open_shard(target, feats, shard, fill, opts, k_block=kb)      # workspace + first staging

    def step():
        stage_shard(solver, target, feats, fill, opts)
        launches = run_sharded(eng, shard, comm, opts["iterations"], poll_every=args.poll,
                               use_graph=bool(args.graph))
        pred, sweeps = shard_result(solver, shard)
        return pred, sweeps, launches, (solver.active_tiles, solver.total_tiles)

    for _ in range(args.warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        pred, sweeps, launches, tiles = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ok = bool(torch.isfinite(pred).all().item())
    solver.close()
    pxit = float(sweeps.sum()) * Hs * Ws
    if rank == 0:
        peak = 6542.4
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        frac_cloud = float(cloud_cells.item()) / (Hs * Ws)
        line = {
            "metric": "vpint2_mpixel_iter_per_s", "value": pxit / (ms * 1e-3) / 1e6, "unit": "Mpx*it/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong", "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C4: VPint2 cloud removal of one %dx%dx%d scene, row-sharded over %d GPU(s), k_block %d, "
                                   "auto-terminate 1e-4" % (Hs, Ws, B, world, kb),
                       "cloud_fraction": round(frac_cloud, 4), "sweeps_per_band": [int(s) for s in sweeps],
                       "launches": int(launches), "active_tiles_rank0": tiles[0], "total_tiles_rank0": tiles[1],
                       "rows_rank0": shard.own_rows, "halo": shard.halo, "finite": ok,
                       "launch_groups": "CUDA graph replay (kernels + NCCL)" if args.graph and world > 0 else "eager"},
            "algorithmic_gbps_whole_grid": 48.0 * pxit / (ms * 1e-3) / 1e9,
            "algorithmic_gbps_unknown_cells": 48.0 * pxit * frac_cloud / (ms * 1e-3) / 1e9,
            "hbm_peak_gbps": peak,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
