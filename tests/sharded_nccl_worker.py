"""torchrun worker: row-sharded VPint2 scene over NCCL, checked against the oracle on rank 0.
Launched by tests/test_sharded_nccl.py (needs >= 2 GPUs)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def main():
    import torch
    import torch.distributed as dist
    from oracle import vpint_oracle as O
    from vpint_b200.sharded import RowShard, solve_scene_sharded
    from vpint_b200.vpint2 import _band_fill_values
    from vpint_b200.wp import wp_run_options
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    rng = np.random.default_rng(77)
    H, W, B = 220, 130, 3
    yy, xx = np.mgrid[0:H, 0:W]
    target = np.empty((H, W, B), dtype=np.float32)
    feats = np.empty((H, W, B), dtype=np.float32)
    for b in range(B):
        base = 1500 + 500 * np.sin(yy / (17.0 + b)) * np.cos(xx / (13.0 + 2 * b))
        target[:, :, b] = np.round(base + rng.normal(0, 15, (H, W)))
        feats[:, :, b] = np.round(1.1 * base + rng.normal(0, 8, (H, W)))
    cloud = ((yy - 100) ** 2 + (xx - 60) ** 2 < 40 ** 2) | ((yy - 30) ** 2 + (xx - 100) ** 2 < 15 ** 2)
    target[cloud, :] = np.nan
    shard = RowShard(H, world, rank, halo=k)
    pred, sweeps = solve_scene_sharded(shard.local(target), shard.local(feats), shard, _band_fill_values(target),
                                       wp_run_options(), k_block=k)
    objs = [None] * world
    dist.all_gather_object(objs, pred.cpu().numpy())          # small scene: gather the owned bands through the host
    parts = [torch.from_numpy(o) for o in objs]
    ok = True
    if rank == 0:
        ref, ref_sweeps = O.vpint2_run(target, feats)
        got = np.moveaxis(np.concatenate([p.cpu().numpy() for p in parts], axis=1), 0, -1)
        rel = float(np.max(np.abs(got - ref) / np.abs(ref)))
        ok = rel <= 1e-9 and list(sweeps) == list(ref_sweeps)
        print("sharded nccl world=%d k=%d rel=%.2e sweeps=%s ref=%s %s" % (world, k, rel, list(sweeps), list(ref_sweeps),
                                                                          "OK" if ok else "MISMATCH"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
