"""Stress of the cluster-resident kernel's protocol (mbarrier / DSMEM halo exchange, lagged stop decision, compact-row
staging, per-class cluster sizes) in lieu of compute-sanitizer, which is closed on this pool: randomized shapes, cloud
layouts and batch mixes, each compared with the tiled one-sweep-per-launch path (an independent implementation of the
same sweep: equal stopping sweeps, values within 1e-12) and run repeatedly for bit-identical results.

Reference semantics: VPint/WP_MRP.py:291-362, VPint/SD_MRP.py:112-147."""

import warnings

import numpy as np
import pytest

from oracle import vpint_oracle as O
from vpint_b200.engine import Solver

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


def smooth(rng, shape, lo=300.0, hi=2700.0, modes=4):
    axes = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    acc = np.zeros(shape)
    for _ in range(modes):
        term = np.ones(shape)
        for ax in axes:
            term = term * np.sin(2 * np.pi * (rng.uniform(0.3, 2.0) * ax + rng.uniform()))
        acc += rng.uniform(0.5, 1.0) * term
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
    return lo + (hi - lo) * acc


def cloud_layout(rng, H, W, kind):
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((H, W), dtype=bool)
    if kind == "discs":
        for _ in range(rng.integers(1, 5)):
            cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(min(H, W) / 20, min(H, W) / 4)
            m |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
    elif kind == "band":                       # a few rows only: tiny compact-row sets, smallest cluster class
        y0 = int(rng.integers(0, H))
        m[y0:y0 + int(rng.integers(1, 6)), int(rng.integers(0, W // 2)):] = True
    elif kind == "two_bands":                  # two separate groups of compact rows (seam rows in between)
        for y0 in (int(rng.integers(0, H // 3)), int(rng.integers(2 * H // 3, H))):
            m[y0:y0 + 3, :int(rng.integers(W // 2, W))] = True
    elif kind == "speckle":
        m = rng.uniform(size=(H, W)) < rng.uniform(0.01, 0.4)
    elif kind == "edges":                      # unknown cells on the grid border and in the corners
        m[0, :] = True; m[-1, ::2] = True; m[:, 0] = True; m[::3, -1] = True
    elif kind == "full_rows":
        m[::7, :] = True
    elif kind == "none":
        pass
    elif kind == "most":
        m = rng.uniform(size=(H, W)) < 0.9
    m[H // 2, W // 2] = m[H // 2, W // 2] and (m.sum() < H * W)   # keep at least one known cell
    if m.all():
        m[0, 0] = False
    return m


def solve_batch(t, f, k_block, max_iter=5000, auto=True, clip=np.inf, scalar=None):
    """t, f: (P, H, W) float64; returns (result, sweeps)."""
    import torch
    P, H, W = t.shape
    s = Solver(2, P, H, W, planes=scalar is None, k_block=k_block)
    try:
        td = torch.from_numpy(t).cuda()
        fill = np.array([np.nanmean(x) if np.isfinite(x).any() else 0.0 for x in t])
        s.load(td, fill=fill)
        if scalar is None:
            s.weights(torch.from_numpy(f).cuda().unsqueeze(-1))
        else:
            s.discounts(scalar)
        s.run(max_iter, auto_terminate=auto, clip_val=clip)
        o, n = s.result()
        return o.cpu().numpy(), n.cpu().numpy()
    finally:
        s.close()


KINDS = ["discs", "band", "two_bands", "speckle", "edges", "full_rows", "none", "most"]


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_and_layouts_against_the_tiled_path(seed):
    rng = np.random.default_rng(9000 + seed)
    H = int(rng.choice([16, 33, 64, 100, 128, 200, 256]))
    W = int(rng.choice([16, 36, 64, 100, 128, 256]))
    P = 70 + int(rng.integers(0, 40))                       # > 66: ordered queue + per-class launches
    t = np.empty((P, H, W))
    f = np.empty((P, H, W))
    for p in range(P):
        base = smooth(rng, (H, W))
        t[p] = np.round(base + rng.normal(0, 15, (H, W)))
        f[p] = np.round(1.1 * base + rng.normal(0, 8, (H, W)))
        t[p][cloud_layout(rng, H, W, KINDS[p % len(KINDS)])] = np.nan
    max_iter = int(rng.choice([1, 2, 7, 40, 5000]))
    auto = bool(max_iter == 5000)
    res, nres = solve_batch(t, f, 0, max_iter, auto)
    til, ntil = solve_batch(t, f, 1, max_iter, auto)
    assert np.array_equal(nres, ntil), (H, W, max_iter)
    err = np.abs(res - til) / np.abs(til)
    assert np.nanmax(err) <= 1e-12, (H, W, float(np.nanmax(err)))
    assert np.array_equal(np.isnan(res), np.isnan(til))
    for _ in range(3):                                      # same launch again: bit-identical
        r2, n2 = solve_batch(t, f, 0, max_iter, auto)
        assert np.array_equal(r2, res, equal_nan=True) and np.array_equal(n2, nres)
    # one problem against the oracle
    p = int(rng.integers(0, P))
    o = O.wp_run_2d(t[p], np.where(np.isnan(t[p]), np.nanmean(t[p]), t[p]), f[p].reshape(H, W, 1).copy(), iterations=-1 if auto else max_iter)
    assert o[2] == nres[p]
    assert np.max(np.abs(res[p] - o[0]) / np.abs(o[0])) <= 1e-9


@pytest.mark.parametrize("n_unknown_segments", [479, 480, 481, 959, 960, 961, 1441, 2400, 2401])
def test_work_lists_around_multiples_of_the_worker_count(n_unknown_segments):
    """480 worker threads per CTA: lists just below / at / above k * 480 segments switch the per-warp sweep variants."""
    rng = np.random.default_rng(n_unknown_segments)
    H, W = 64, 256                                           # one CTA per problem (C = 1 class) holds 64 rows
    P = 4
    t = np.empty((P, H, W))
    f = np.empty((P, H, W))
    for p in range(P):
        base = smooth(rng, (H, W))
        t[p] = np.round(base + rng.normal(0, 15, (H, W)))
        f[p] = np.round(1.1 * base + rng.normal(0, 8, (H, W)))
        m = np.zeros((H, W // 4), dtype=bool)
        m.reshape(-1)[rng.permutation(H * W // 4)[:n_unknown_segments + p]] = True     # segments with >= 1 unknown cell
        cells = np.repeat(m, 4, axis=1) & (rng.uniform(size=(H, W)) < 0.7)
        cells[:, ::4] |= m                                   # every chosen segment keeps an unknown cell
        cells[0, 0] = False
        t[p][cells] = np.nan
    res, nres = solve_batch(t, f, 0)
    til, ntil = solve_batch(t, f, 1)
    assert np.array_equal(nres, ntil)
    assert np.nanmax(np.abs(res - til) / np.abs(til)) <= 1e-12


def test_fifty_repeated_runs_are_bit_identical():
    rng = np.random.default_rng(4242)
    H = W = 256
    P = 80
    t = np.empty((P, H, W))
    f = np.empty((P, H, W))
    for p in range(P):
        base = smooth(rng, (H, W))
        t[p] = np.round(base + rng.normal(0, 15, (H, W)))
        f[p] = np.round(1.1 * base + rng.normal(0, 8, (H, W)))
        t[p][cloud_layout(rng, H, W, "discs" if p % 3 else "band")] = np.nan
    first = solve_batch(t, f, 0)
    for _ in range(49):
        again = solve_batch(t, f, 0)
        assert np.array_equal(again[1], first[1])
        assert np.array_equal(again[0], first[0], equal_nan=True)
    til = solve_batch(t, f, 1)
    assert np.array_equal(first[1], til[1])
    assert np.nanmax(np.abs(first[0] - til[0]) / np.abs(til[0])) <= 1e-12


def test_scalar_gamma_and_clip_variants_in_a_mixed_batch():
    rng = np.random.default_rng(77)
    H, W, P = 100, 100, 90
    t = np.empty((P, H, W))
    for p in range(P):
        t[p] = smooth(rng, (H, W), 10, 30) + rng.normal(0, 0.5, (H, W))
        t[p][cloud_layout(rng, H, W, KINDS[p % len(KINDS)])] = np.nan
    gam = rng.uniform(0.5, 0.99, P)
    res, nres = solve_batch(t, None, 0, scalar=gam)
    til, ntil = solve_batch(t, None, 1, scalar=gam)
    assert np.array_equal(nres, ntil)
    assert np.nanmax(np.abs(res - til) / np.abs(til)) <= 1e-12
    f = np.stack([np.round(1.1 * smooth(rng, (H, W))) for _ in range(P)])
    tt = np.stack([np.round(smooth(rng, (H, W))) for _ in range(P)])
    for p in range(P):
        tt[p][cloud_layout(rng, H, W, "discs")] = np.nan
    res, nres = solve_batch(tt, f, 0, clip=1500.0)
    til, ntil = solve_batch(tt, f, 1, clip=1500.0)
    assert np.array_equal(nres, ntil)
    assert np.nanmax(np.abs(res - til) / np.abs(til)) <= 1e-12
