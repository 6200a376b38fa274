"""Host-side logic of the row-sharded multi-rank path on CPU: world_size-2 and -3 gloo
process groups drive vpint_b200.sharded.run_sharded over a NumPy stand-in for the CUDA solver and must
reproduce the single-process oracle (values and per-band stopping sweeps)."""

import os
import socket
import tempfile
import warnings

import numpy as np
import pytest

from oracle import vpint_oracle as O
from vpint_b200.errors import VPintError
from vpint_b200.sharded import RowShard, run_lockstep, run_sharded

warnings.filterwarnings("ignore")


class NumpyShardEngine:
    """CPU test double of the engine protocol: k Jacobi sweeps per launch on the LOCAL rows of every band
    (rows beyond the local slice read as 0, so halo rows decay one row per sweep exactly as on the
    device), per-sweep sums over owned unknown cells, commit-then-break with in-block rollback."""

    def __init__(self, target, feats, shard, k, thr=1e-4, auto=True):
        import torch
        self.torch = torch
        self.sh, self.k, self.thr, self.auto = shard, k, thr, auto
        H, W, B = target.shape
        self.B, self.W, self.n = B, W, float(H * W)
        sl = slice(shard.lo, shard.hi)
        self.orig = [target[sl, :, b].astype(np.float64) for b in range(B)]
        self.state = [O.init_pred(target[:, :, b])[sl].astype(np.float64) for b in range(B)]
        self.w = [O.weights_exact_2d(feats[:, :, b].astype(np.float64).reshape(H, W, 1).copy())[sl] for b in range(B)]
        self.nc = O.neighbour_count((H, W))[sl]
        self.unknown = [np.isnan(o) for o in self.orig]
        own = np.zeros(self.orig[0].shape, dtype=bool)
        own[shard.own_row0:shard.own_row0 + shard.own_rows] = True
        self.own = own
        self._known = torch.zeros(2 * B * 4, dtype=torch.float64)
        kv = self._known.view(2, B, 4)
        for b in range(B):
            sel = own & ~self.unknown[b]
            kv[0, b, 1] = float(self.state[b][sel].sum())
            kv[1, b, 1] = float(self.orig[b][sel].sum())
        self._sums = torch.zeros(B, k, 4, dtype=torch.float64)
        self.msg = {n: torch.zeros(B, k, W, dtype=torch.float64) for n in ("su", "sd", "ru", "rd")}

    def known(self):
        return self._known

    def begin(self, max_iter):
        self.max_iter = max_iter
        self.iters = [0] * self.B
        self.done = [max_iter <= 0] * self.B
        self.first = True

    def _sweep(self, b, p):
        v = np.zeros(p.shape + (4,))
        v[1:, :, 0] = p[:-1]
        v[:, :-1, 1] = p[:, 1:]
        v[:-1, :, 2] = p[1:]
        v[:, 1:, 3] = p[:, :-1]
        new = (v * self.w[b]).sum(axis=2) / self.nc
        return np.where(self.unknown[b], new, self.orig[b])

    def step_begin(self):
        self.snap = []
        self._sums.zero_()
        for b in range(self.B):
            snaps = [self.state[b]]
            if not self.done[b]:
                kr = min(self.k, self.max_iter - self.iters[b])
                for j in range(kr):
                    new = self._sweep(b, snaps[-1])
                    sel = self.own & self.unknown[b]
                    self._sums[b, j, 0] = float(np.abs(new - snaps[-1])[sel].sum())
                    self._sums[b, j, 1] = float(snaps[-1][sel].sum())
                    snaps.append(new)
            self.snap.append(snaps)

    def sums(self):
        return self._sums

    def step_end(self):
        kv = self._known.view(2, self.B, 4)
        for b in range(self.B):
            if self.done[b]:
                continue
            kr = len(self.snap[b]) - 1
            stop = None
            for j in range(kr):
                K = kv[0 if (self.first and j == 0) else 1, b]
                delta = ((float(self._sums[b, j, 0]) + float(K[0])) / self.n) / ((float(self._sums[b, j, 1]) + float(K[1])) / self.n)
                if self.auto and delta <= self.thr:
                    stop = j + 1
                    break
            take = stop if stop is not None else kr
            self.state[b] = self.snap[b][take]
            self.iters[b] += take
            if stop is not None or self.iters[b] >= self.max_iter:
                self.done[b] = True
        self.first = False

    def halo_send(self):
        s, k = self.sh, self.k
        for b in range(self.B):
            self.msg["su"][b] = self.torch.from_numpy(self.state[b][s.own_row0:s.own_row0 + k].copy())
            self.msg["sd"][b] = self.torch.from_numpy(self.state[b][s.own_row0 + s.own_rows - k:s.own_row0 + s.own_rows].copy())
        return self.msg["su"], self.msg["sd"]

    def halo_recv(self):
        return self.msg["ru"], self.msg["rd"]

    def halo_apply(self):
        s, k = self.sh, self.k
        for b in range(self.B):
            st = self.state[b].copy()
            if s.up is not None:
                st[s.own_row0 - k:s.own_row0] = self.msg["ru"][b].numpy()
            if s.down is not None:
                st[s.own_row0 + s.own_rows:s.own_row0 + s.own_rows + k] = self.msg["rd"][b].numpy()
            self.state[b] = st

    def active(self):
        return sum(0 if d else 1 for d in self.done)

    def owned_result(self):
        s = self.sh
        return np.stack([st[s.own_row0:s.own_row0 + s.own_rows] for st in self.state])


def make_scene(H=46, W=30, B=3, seed=5):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    target = np.empty((H, W, B), dtype=np.float32)
    feats = np.empty((H, W, B), dtype=np.float32)
    for b in range(B):
        base = 1500 + 500 * np.sin(yy / (7.0 + b)) * np.cos(xx / (5.0 + 2 * b))
        target[:, :, b] = np.round(base + rng.normal(0, 15, (H, W)))
        feats[:, :, b] = np.round(1.1 * base + rng.normal(0, 8, (H, W)))
    cloud = ((yy - 20) ** 2 + (xx - 12) ** 2 < 81) | ((yy - 38) ** 2 + (xx - 22) ** 2 < 30)
    target[cloud, :] = np.nan
    return target, feats


def _worker(rank, world, port, k, out_dir):
    import torch.distributed as dist
    from vpint_b200.sharded import Comm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        target, feats = make_scene()
        shard = RowShard(target.shape[0], world, rank, halo=k)
        eng = NumpyShardEngine(target, feats, shard, k)
        launches = run_sharded(eng, shard, Comm(), 5000, poll_every=3)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pred=eng.owned_result(), iters=np.array(eng.iters),
                 r0=shard.r0, r1=shard.r1, launches=launches)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,k", [(2, 1), (2, 4), (3, 2)])
def test_run_sharded_over_gloo_matches_oracle(world, k):
    import torch.multiprocessing as mp
    target, feats = make_scene()
    ref, ref_sweeps = O.vpint2_run(target, feats)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), k, d), nprocs=world, join=True)
        got = np.empty_like(ref)
        for r in range(world):
            z = np.load(os.path.join(d, "rank%d.npz" % r))
            got[int(z["r0"]):int(z["r1"])] = np.moveaxis(z["pred"], 0, -1)
            assert list(z["iters"]) == list(ref_sweeps), (r, z["iters"], ref_sweeps)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-9


def test_lockstep_driver_matches_oracle():
    target, feats = make_scene(seed=9)
    ref, ref_sweeps = O.vpint2_run(target, feats, iterations=11)
    world, k = 3, 4
    shards = [RowShard(target.shape[0], world, r, halo=k) for r in range(world)]
    engines = [NumpyShardEngine(target, feats, s, k, auto=False) for s in shards]
    run_lockstep(engines, shards, 11, poll_every=2)
    got = np.concatenate([e.owned_result() for e in engines], axis=1)
    assert np.max(np.abs(np.moveaxis(got, 0, -1) - ref) / np.abs(ref)) <= 1e-9
    assert all(e.iters == [11] * 3 for e in engines)


def test_row_shard_partition():
    H = 10980
    parts = [RowShard(H, 8, r, halo=8) for r in range(8)]
    assert parts[0].r0 == 0 and parts[-1].r1 == H
    assert all(a.r1 == b.r0 for a, b in zip(parts[:-1], parts[1:]))
    assert sorted(set(p.own_rows for p in parts)) == [1372, 1373]
    assert parts[0].own_row0 == 0 and parts[0].local_rows == parts[0].own_rows + 8
    assert parts[3].own_row0 == 8 and parts[3].local_rows == parts[3].own_rows + 16
    assert parts[3].desc() == (H, parts[3].r0 - 8, 8, parts[3].own_rows)
    assert parts[0].up is None and parts[0].down == 1 and parts[7].down is None
    with pytest.raises(VPintError):
        RowShard(3, 8, 0, halo=1)
