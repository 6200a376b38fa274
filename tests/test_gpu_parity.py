"""Parity of the CUDA path (through the C ABI) with the reference.

Three layers:
  1. reference-generated golden fixtures (tests/golden/*.npz),
  2. the NumPy oracle on the same seeded inputs at sizes it finishes in seconds,
  3. size-independent properties at larger sizes.

Bar (BASELINE.json north_star): fp64 mode within 1e-9 relative per pixel with the
SAME iteration count; fp32-storage mode within 1e-4 relative.
"""

import warnings

import numpy as np
import pytest

from oracle import vpint_oracle as O
from vpint_b200 import SD_SMRP, SD_STMRP, VPint2_interpolator, VPintError, WP_SMRP, WP_STMRP, run_patches
from vpint_b200 import solve as solve_mod
from vpint_b200.batch import sd_trials
from vpint_b200.engine import Solver

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")

RTOL64 = 1e-9     # north_star: fp64 mode
RTOL32 = 1e-4     # north_star: fp32-storage mode
K_BLOCKS = [0, 1, 4, 8]    # 0 = library default: cluster-resident solve where a problem fits on chip
# Per-sweep deltas are ratios of grid means; once a solve has converged to the last ulp the numerator is
# pure rounding noise (~1e-16 of the state), so deltas get an absolute floor next to the 1e-9 relative bar.
DELTA_ATOL = 1e-14


def close(a, b, rtol=RTOL64, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    inf = np.isinf(b)
    assert np.array_equal(np.isinf(a), inf) and np.array_equal(a[inf], b[inf]), "inf pattern differs"
    ok = ~(np.isnan(b) | inf)
    err = np.abs(a[ok] - b[ok])
    bound = rtol * np.abs(b[ok]) + atol + 1e-300
    assert (err <= bound).all(), "max rel err %.3e" % float(np.max(err / (np.abs(b[ok]) + 1e-300)))


def closed(a, b, rtol=RTOL64):
    close(a, b, rtol=rtol, atol=DELTA_ATOL)


@pytest.fixture(params=K_BLOCKS, ids=lambda k: "k%d" % k)
def kblock(request, monkeypatch):
    monkeypatch.setattr(solve_mod, "DEFAULT_K_BLOCK", request.param)
    return request.param


# ---- 1. golden fixtures produced by the reference itself --------------------------------------
def test_golden_sd2d(golden, kblock):
    g = golden("sd2d")
    grid = g["grid"]
    m = SD_SMRP(grid.copy(), gamma=0.9)
    p, d = m.run(track_delta=True)
    assert len(d) == len(g["d_auto"])
    close(p, g["pred_auto"]); closed(d, g["d_auto"])
    assert p is m.pred_grid and p.dtype == np.float64
    m = SD_SMRP(grid.copy(), gamma=0.73)
    p, d = m.run(iterations=7, track_delta=True)
    close(p, g["pred_fix"]); closed(d, g["d_fix"])
    m = SD_SMRP(grid.copy(), gamma=0.9, init_strategy="zero")
    p, d = m.run(iterations=3, track_delta=True)
    close(p, g["pred_zero"]); closed(d, g["d_zero"])
    m = SD_SMRP(grid.copy(), gamma=0.73)       # resume: run(3); run(4) == run(7)
    m.run(iterations=3)
    close(m.run(iterations=4), g["pred_resume"])
    m = SD_SMRP(grid.copy(), gamma=0.9)        # reset() leaves NaN in the start state
    m.reset()
    p, d = m.run(iterations=4, track_delta=True)
    close(p, g["pred_reset"]); closed(d, g["d_reset"])


def test_golden_wp2d(golden, kblock):
    g = golden("wp2d")
    grid, feat = g["grid"], g["feat"]
    m = WP_SMRP(grid.copy(), feat.copy())
    p, d = m.run(track_delta=True)
    assert len(d) == len(g["d_auto"]) == 14
    close(p, g["pred_auto"]); closed(d, g["d_auto"])
    assert m.run_state and m.run_method == "exact"
    p, d = WP_SMRP(grid.copy(), feat.copy()).run(iterations=9, track_delta=True)
    close(p, g["pred_fix"]); closed(d, g["d_fix"])
    m = WP_SMRP(grid.copy(), feat.copy())
    p, d = m.run(iterations=6, track_delta=True, clip_val=float(g["clip_val"]))
    close(p, g["pred_clip"]); closed(d, g["d_clip"])
    assert m.clip_val == float(g["clip_val"])
    for method in ("cosine_similarity", "exact_inverse"):
        m = WP_SMRP(grid.copy(), feat.copy(), min_gamma=0.2, max_gamma=1.05)
        p, d = m.run(iterations=5, method=method, track_delta=True)
        close(p, g["pred_" + method]); closed(d, g["d_" + method])


def test_golden_wp2d_edge_cases(golden, kblock):
    g = golden("wp2d_edge")
    m = WP_SMRP(g["z_grid"].copy(), g["z_feat"].copy())          # zeros in the features
    p, d = m.run(iterations=5, track_delta=True)
    close(p, g["z_pred"]); closed(d, g["z_d"])
    assert np.array_equal(m.feature_grid, g["z_feat_after"])
    p, d = WP_SMRP(g["f9_grid"].copy(), g["f9_feat"].copy()).run(iterations=4, track_delta=True)
    close(p, g["f9_pred"]); closed(d, g["f9_d"])
    p, d = WP_SMRP(g["f9_grid"].copy(), g["f19_feat"].copy()).run(iterations=3, track_delta=True)
    close(p, g["f19_pred"]); closed(d, g["f19_d"])
    for tag in ("col", "row"):                                   # H or W == 1: neighbour count 2 / 0
        p, d = WP_SMRP(g[tag + "_grid"].copy(), g[tag + "_feat"].copy()).run(iterations=3, track_delta=True)
        close(p, g[tag + "_pred"]); closed(d, g[tag + "_d"])
    p, d = WP_SMRP(g["inf_grid"].copy(), g["inf_feat"].copy()).run(iterations=8, track_delta=True)   # blow-up to inf / NaN
    close(p, g["inf_pred"]); closed(d, g["inf_d"])
    p, d = WP_SMRP(g["inf_grid"].copy(), g["inf_feat"].copy()).run(track_delta=True)
    assert len(d) == 3
    close(p, g["inf_pred_auto"]); closed(d, g["inf_d_auto"])


def test_golden_st3d(golden):
    g = golden("st3d")
    cube = g["cube"]
    p, d = SD_STMRP(cube.copy(), gamma=0.9, tau=0.8).run(track_delta=True)
    assert len(d) == len(g["sd_d_auto"])
    close(p, g["sd_auto"]); closed(d, g["sd_d_auto"])
    p, d = SD_STMRP(cube.copy(), gamma=0.6, tau=0.95).run(iterations=6, track_delta=True)
    close(p, g["sd_fix"]); closed(d, g["sd_d_fix"])
    feat = g["feat"]
    p, d = WP_STMRP(cube.copy(), feat.copy()).run(iterations=5, method="exact", track_delta=True)
    close(p, g["wp_exact"]); closed(d, g["wp_d_exact"])
    p, d = WP_STMRP(cube.copy(), feat.copy()).run(iterations=4, method="cosine_similarity", track_delta=True)
    close(p, g["wp_cos"]); closed(d, g["wp_d_cos"])
    p, d = WP_STMRP(cube.copy(), feat.copy()).run(method="exact", track_delta=True)
    assert len(d) == len(g["wp_d_auto"])
    close(p, g["wp_auto"]); closed(d, g["wp_d_auto"])
    p, d = SD_STMRP(g["flat"].copy(), gamma=0.9, tau=0.9).run(iterations=3, track_delta=True)
    close(p, g["sd_flat"]); closed(d, g["sd_d_flat"])


def test_golden_vpint2(golden, kblock):
    g = golden("vpint2")
    vp = VPint2_interpolator(g["target"], g["features"], mask=g["mask"], threshold=10)
    p = vp.run()
    assert p.shape == g["pred_par"].shape and p.dtype == np.float64
    close(p, g["pred_par"])
    ps = vp.run_serial()
    assert ps.dtype == np.float32 and ps.flags["C_CONTIGUOUS"]
    close(ps, g["pred_serial"], rtol=2e-7)     # float32 rounding of an fp64 result within 1e-9
    close(vp.run_serial(iterations=5), g["pred_fix"], rtol=2e-7)
    close(vp.run(iterations=4, clip_val=1500.0), g["pred_clip"])
    from vpint_b200.vpint2 import solve_bands
    from vpint_b200.wp import wp_run_options
    _, sweeps = solve_bands(vp.target, vp.features, wp_run_options(), return_sweeps=True)
    assert list(sweeps) == list(g["sweeps"])


# ---- 2. the oracle on seeded inputs ------------------------------------------------------------
def smooth(rng, shape, lo=300.0, hi=2700.0, modes=5):
    axes = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    acc = np.zeros(shape)
    for _ in range(modes):
        term = np.ones(shape)
        for ax in axes:
            term = term * np.sin(2 * np.pi * (rng.uniform(0.3, 2.0) * ax + rng.uniform()))
        acc += rng.uniform(0.5, 1.0) * term
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
    return lo + (hi - lo) * acc


def discs(rng, shape, n, rmin, rmax):
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    m = np.zeros(shape, dtype=bool)
    for _ in range(n):
        cy, cx, r = rng.uniform(0, shape[0]), rng.uniform(0, shape[1]), rng.uniform(rmin, rmax)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
    return m


def test_oracle_sd2d_config1_shape(kblock):
    """BASELINE config 1: SD_SMRP on a 100x100 grid, ~20 % uniformly hidden."""
    rng = np.random.default_rng(1001)
    grid = smooth(rng, (100, 100), 15, 25) + rng.normal(0, 1.0, (100, 100))
    grid[rng.uniform(size=grid.shape) < 0.2] = np.nan
    ref, dref, n = O.sd_run_2d(grid, O.init_pred(grid), gamma=0.9)
    p, d = SD_SMRP(grid.copy(), gamma=0.9).run(track_delta=True)
    assert len(d) == n
    close(p, ref); closed(d, dref)
    ref, dref, _ = O.sd_run_2d(grid, O.init_pred(grid), gamma=0.9, iterations=100)
    p, d = SD_SMRP(grid.copy(), gamma=0.9).run(iterations=100, track_delta=True)
    close(p, ref); closed(d, dref)


@pytest.mark.parametrize("shape", [(256, 256), (130, 203), (67, 300)])
def test_oracle_wp2d(shape, kblock):
    """BASELINE config 2 recipe (WP_SMRP, 3 feature layers, disc clouds) at oracle-friendly sizes,
    including ragged shapes that do not divide the tile or the row pitch."""
    rng = np.random.default_rng(1002 + shape[1])
    base = smooth(rng, shape)
    grid = base + rng.normal(0, 20, shape)
    feat = np.stack([rng.uniform(0.7, 1.3) * base + rng.normal(0, 10, shape) for _ in range(3)], axis=-1)
    feat = np.maximum(feat, 50.0)
    grid[discs(rng, shape, 6, 8, 40)] = np.nan
    ref, dref, n = O.wp_run_2d(grid, O.init_pred(grid), feat.copy())
    p, d = WP_SMRP(grid.copy(), feat.copy()).run(track_delta=True)
    assert len(d) == n, (len(d), n)
    close(p, ref); closed(d, dref)
    ref, dref, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=25, clip_val=2000.0)
    p, d = WP_SMRP(grid.copy(), feat.copy()).run(iterations=25, track_delta=True, clip_val=2000.0)
    close(p, ref); closed(d, dref)


def test_oracle_st3d():
    rng = np.random.default_rng(1005)
    shape = (24, 20, 12)
    cube = smooth(rng, shape, 10, 30) + rng.normal(0, 0.3, shape)
    for t in range(shape[2]):
        cube[discs(rng, shape[:2], 2, 2, 6), t] = np.nan
    ref, dref, n = O.sd_run_3d(cube, O.init_pred(cube), gamma=0.9, tau=0.9)
    p, d = SD_STMRP(cube.copy(), gamma=0.9, tau=0.9).run(track_delta=True)
    assert len(d) == n
    close(p, ref); closed(d, dref)
    feat = smooth(rng, shape[:2] + (1,), 1, 3)
    ref, dref, n = O.wp_run_3d(cube, O.init_pred(cube), feat.copy(), method="exact")
    p, d = WP_STMRP(cube.copy(), feat.copy()).run(method="exact", track_delta=True)
    assert len(d) == n
    close(p, ref); closed(d, dref)


def make_patch(rng, H=64, W=64, B=4):
    target = np.stack([np.round(smooth(rng, (H, W))) for _ in range(B)], axis=-1).astype(np.uint16)
    feats = np.stack([np.round(1.1 * target[:, :, b] + rng.normal(0, 10, (H, W))) for b in range(B)], axis=-1)
    feats = np.clip(feats, 1, 65535).astype(np.uint16)
    mask = np.where(discs(rng, (H, W), 3, 4, 14), 100.0, 0.0)
    return VPint2_interpolator(target, feats, mask=mask, threshold=10)


def test_oracle_vpint2_patch_batch(kblock):
    """BASELINE config 3 recipe: a batch of Sentinel-2-shaped patches, every (patch, band) its own
    problem with its own stopping sweep."""
    rng = np.random.default_rng(1003)
    vps = [make_patch(rng) for _ in range(5)]
    outs = run_patches(vps)
    for vp, out in zip(vps, outs):
        ref, sweeps = O.vpint2_run(vp.target, vp.features)
        close(out, ref)
        assert out.shape == ref.shape
    one = vps[2].run()
    close(one, outs[2], rtol=0)      # batching must not change a problem's result at all
    ser = run_patches(vps[:2], serial=True, iterations=12)
    ref, _ = O.vpint2_run(vps[1].target, vps[1].features, iterations=12)
    close(ser[1], ref.astype(np.float32), rtol=2e-7)


def test_oracle_sweep_counts_per_problem(kblock):
    rng = np.random.default_rng(77)
    vp = make_patch(rng, 96, 80, 5)
    from vpint_b200.vpint2 import solve_bands
    from vpint_b200.wp import wp_run_options
    out, sweeps = solve_bands(vp.target, vp.features, wp_run_options(), return_sweeps=True)
    ref, ref_sweeps = O.vpint2_run(vp.target, vp.features)
    assert list(sweeps) == list(ref_sweeps)
    assert len(set(ref_sweeps)) > 1, "fixture should stop bands at different sweeps"
    close(np.moveaxis(out, 0, -1), ref)


def test_find_gamma_batched_trials_match_reference_rng_order():
    rng = np.random.default_rng(5)
    grid = smooth(rng, (30, 30), 10, 30)
    grid[rng.uniform(size=grid.shape) < 0.3] = np.nan
    np.random.seed(11)
    m = SD_SMRP(grid.copy())
    best = m.find_gamma(6, 0.5, sub_iterations=20)
    # replay the reference's RNG order on the host and score with the oracle
    np.random.seed(11)
    sub = grid.copy()
    for i in range(30):
        for j in range(30):
            if not np.isnan(sub[i, j]) and np.random.rand() < 0.5:
                sub[i, j] = np.nan
    best_ref, best_loss = 0.9, np.inf
    for _ in range(6):
        gam = np.random.uniform(low=0, high=1)
        pred, _, _ = O.sd_run_2d(sub, O.init_pred(sub), gamma=gam, iterations=20)
        sel = ~np.isnan(grid) & np.isnan(sub)
        mae = np.abs(pred[sel] - grid[sel]).sum() / sel.sum()
        if mae < best_loss:
            best_ref, best_loss = gam, mae
    assert isinstance(best, float) and best == best_ref and m.gamma == best_ref


# ---- 3. properties at larger sizes ---------------------------------------------------------------
def test_properties_large_wp(kblock):
    rng = np.random.default_rng(2024)
    shape = (1024, 1024)                                    # BASELINE config 2 size
    base = smooth(rng, shape)
    feat = np.stack([rng.uniform(0.7, 1.3) * base for _ in range(3)], axis=-1)
    cloud = discs(rng, shape, 12, 30, 120)
    grid = base.copy()
    grid[cloud] = np.nan
    m = WP_SMRP(grid.copy(), feat.copy())
    p, d = m.run(iterations=40, track_delta=True)
    assert np.array_equal(p[~cloud], grid[~cloud]), "known cells must come back bit-exact"
    assert not np.isnan(p).any() and d[-1] < d[0]
    # linearity: scaling the target scales the result, the stopping statistic is scale free
    m2 = WP_SMRP(4.0 * grid, feat.copy())
    p2, d2 = m2.run(iterations=40, track_delta=True)
    close(p2, 4.0 * p, rtol=1e-12); closed(d2, d, rtol=1e-12)
    # resume == one longer run
    m3 = WP_SMRP(grid.copy(), feat.copy())
    m3.run(iterations=15)
    close(m3.run(iterations=25), p, rtol=0)
    # target proportional to a single feature is a fixed point of the exact weights
    m4 = WP_SMRP(np.where(cloud, np.nan, 0.5 * base), base.copy())
    m4.pred_grid = 0.5 * base
    close(m4.run(iterations=3), 0.5 * base, rtol=1e-12)


def test_known_answers():
    g = np.full((3, 3), 2.0)
    g[1, 1] = np.nan
    g[0, 1], g[1, 2], g[2, 1], g[1, 0] = 1.0, 2.0, 3.0, 6.0
    assert SD_SMRP(g, gamma=1.0).run(1)[1, 1] == 3.0          # mean of the 4 neighbours
    c = np.full((5, 5), 7.0)
    for (i, j) in ((2, 2), (0, 2), (0, 0)):                   # interior / edge / corner: exactly gamma * c
        h = c.copy()
        h[i, j] = np.nan
        assert SD_SMRP(h, gamma=0.5).run(1)[i, j] == 3.5


def test_start_state_differs_from_original_on_known_cells(kblock):
    """pred_grid edited by the caller: the first sweep reads it, known cells are then restored."""
    rng = np.random.default_rng(9)
    grid = smooth(rng, (40, 40), 10, 20)
    grid[rng.uniform(size=grid.shape) < 0.3] = np.nan
    start = O.init_pred(grid) * rng.uniform(0.9, 1.1, grid.shape)
    ref, dref, _ = O.sd_run_2d(grid, start.copy(), gamma=0.8, iterations=6)
    m = SD_SMRP(grid.copy(), gamma=0.8)
    m.pred_grid = start.copy()
    p, d = m.run(iterations=6, track_delta=True)
    close(p, ref); closed(d, dref)


def test_fp32_storage_mode():
    rng = np.random.default_rng(31)
    shape = (200, 160)
    base = smooth(rng, shape)
    grid = base + rng.normal(0, 20, shape)
    feat = (1.1 * base + rng.normal(0, 10, shape)).reshape(shape + (1,))
    grid[discs(rng, shape, 5, 8, 30)] = np.nan
    ref, _, n = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=60)
    res, _, sweeps = solve_mod.propagate(grid, O.init_pred(grid), planes=True, features=feat, max_iter=60,
                                         auto_terminate=False, threshold=1e-4, storage="float32")
    assert sweeps == 60
    close(res, ref, rtol=RTOL32)


def test_fp32_storage_mode_auto_terminate_on_benchmark_patches():
    """north_star's fp32 mode on the C3 recipe (256 x 256 x 13 uint16-valued patches, disc clouds), run to the
    reference's own stopping rule: float32 state between the sweeps, fp64 arithmetic and fp64 stopping sums -- the
    same number of sweeps per band as the fp64 reference and values within 1e-4."""
    from vpint_b200.vpint2 import solve_bands
    from vpint_b200.wp import wp_run_options
    import bench
    t, f, m = bench.make_patches_numpy(2, 515)
    total = equal = 0
    for i in range(2):
        vp = VPint2_interpolator(t[i], f[i], mask=m[i], threshold=10)
        out, sweeps = solve_bands(vp.target, vp.features, wp_run_options(), storage="float32", k_block=1, return_sweeps=True)
        ref, ref_sweeps = O.vpint2_run(vp.target, vp.features)
        close(np.moveaxis(out, 0, -1), ref, rtol=RTOL32)
        total += len(ref_sweeps)
        equal += int(np.sum(np.asarray(sweeps) == np.asarray(ref_sweeps)))
        assert np.max(np.abs(np.asarray(sweeps) - np.asarray(ref_sweeps))) <= 1, (list(sweeps), list(ref_sweeps))
    # rounding the state to float32 moves delta by ~1e-7 relative: a band stops one sweep apart only if its delta
    # passes the threshold that closely
    assert equal >= total - 1, (equal, total)


def test_batched_trials_and_stride0_inputs():
    rng = np.random.default_rng(8)
    grid = smooth(rng, (50, 50), 10, 30)
    grid[rng.uniform(size=grid.shape) < 0.25] = np.nan
    gammas = np.array([0.3, 0.9, 0.55])
    out = sd_trials(grid, gammas, 15)
    for i, gam in enumerate(gammas):
        ref, _, _ = O.sd_run_2d(grid, O.init_pred(grid), gamma=gam, iterations=15)
        close(out[i], ref)


def test_solver_introspection_and_tile_skipping():
    import torch
    grid = np.ones((256, 512))
    grid[10:20, 10:20] = np.nan                     # one small cloud: most tiles inactive
    s = Solver(2, 1, 256, 512, planes=False)
    s.load(torch.from_numpy(grid).cuda().unsqueeze(0), fill=[1.0])
    assert 0 < s.active_tiles < s.total_tiles
    s.discounts(0.9)
    s.run(5, auto_terminate=False)
    out, niter = s.result()
    assert int(niter[0]) == 5 and s.launches > 0
    ref, _, _ = O.sd_run_2d(grid[:, :256].copy(), np.where(np.isnan(grid[:, :256]), 1.0, grid[:, :256]), gamma=0.9, iterations=5)
    close(out[0].cpu().numpy()[:20, :20], ref[:20, :20])
    s.close()
    with pytest.raises(VPintError):
        WP_SMRP(np.ones((4, 4)), np.ones((4, 4))).run(iterations=1, save_gif=True)


# ---- 4. row-sharded scene on one device: shards in lockstep through the CUDA halo kernels ---------
@pytest.mark.parametrize("world,k", [(2, 1), (3, 4), (4, 8)])
def test_row_sharded_scene_lockstep(world, k):
    from vpint_b200.sharded import RowShard, open_shard, run_lockstep, shard_result
    from vpint_b200.vpint2 import _band_fill_values
    from vpint_b200.wp import wp_run_options
    rng = np.random.default_rng(404)
    vp = make_patch(rng, 150, 90, 3)
    ref, ref_sweeps = O.vpint2_run(vp.target, vp.features)
    fill = _band_fill_values(vp.target)
    shards = [RowShard(150, world, r, halo=k) for r in range(world)]
    opened = [open_shard(s.local(vp.target), s.local(vp.features), s, fill, wp_run_options(), k_block=k) for s in shards]
    try:
        run_lockstep([e for _, e in opened], shards, wp_run_options()["iterations"], poll_every=4)
        parts = [shard_result(sv, s) for (sv, _), s in zip(opened, shards)]
    finally:
        for sv, _ in opened:
            sv.close()
    got = np.concatenate([p.cpu().numpy() for p, _ in parts], axis=1)
    for _, sweeps in parts:
        assert list(sweeps) == list(ref_sweeps)
    close(np.moveaxis(got, 0, -1), ref)


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("iterations", [-1, 7, 1])
def test_row_sharded_scene_lockstep_lagged(world, iterations):
    """The low-latency protocol of the row-sharded loop on one device: stop decision lagging one sweep, halo rows
    pushed into the neighbour's receive slots with a flag (the kernels that run across GPUs over NVLink)."""
    from vpint_b200.sharded import RowShard, open_shard, run_lockstep_lagged, shard_result
    from vpint_b200.vpint2 import _band_fill_values
    from vpint_b200.wp import wp_run_options
    rng = np.random.default_rng(405 + world)
    vp = make_patch(rng, 150, 90, 3)
    opts = wp_run_options(iterations=iterations)
    ref, ref_sweeps = O.vpint2_run(vp.target, vp.features, iterations=iterations)
    fill = _band_fill_values(vp.target)
    shards = [RowShard(150, world, r, halo=1) for r in range(world)]
    opened = [open_shard(s.local(vp.target), s.local(vp.features), s, fill, opts, k_block=1) for s in shards]
    try:
        for _ in range(2):                                  # a second run() on the same engines (sequence numbers go on)
            for (sv, e), s in zip(opened, shards):
                from vpint_b200.sharded import stage_shard
                import torch
                stage_shard(sv, torch.from_numpy(np.ascontiguousarray(s.local(vp.target))).cuda(),
                            torch.from_numpy(np.ascontiguousarray(s.local(vp.features))).cuda(), fill, opts)
            run_lockstep_lagged([e for _, e in opened], shards, opts["iterations"])
            parts = [shard_result(sv, s) for (sv, _), s in zip(opened, shards)]
            got = np.concatenate([p.cpu().numpy() for p, _ in parts], axis=1)
            for _, sweeps in parts:
                assert list(sweeps) == list(ref_sweeps)
            close(np.moveaxis(got, 0, -1), ref)
    finally:
        for sv, _ in opened:
            sv.close()


def test_in_library_sharded_loop_single_rank():
    """vpint_solver_run_sharded with a world of one (the peer block is an ordinary device buffer): the peer commit
    kernel, the launch record in mapped pinned memory and the host's run-ahead polling, against the ordinary run()."""
    import torch
    from vpint_b200.wp import wp_run_options
    rng = np.random.default_rng(407)
    vp = make_patch(rng, 120, 200, 4)
    ref, ref_sweeps = O.vpint2_run(vp.target, vp.features)
    for iterations in (-1, 9, 1):
        rr, rs = O.vpint2_run(vp.target, vp.features, iterations=iterations) if iterations != -1 else (ref, ref_sweeps)
        opts = wp_run_options(iterations=iterations)
        s = Solver(2, 4, 120, 200, nprob_inner=4, planes=True, k_block=1)
        try:
            t = torch.from_numpy(vp.target[None]).cuda()
            f = torch.from_numpy(vp.features[None]).cuda()
            block = torch.zeros(s.peer_bytes(1), dtype=torch.uint8, device="cuda")
            s.peer_bind(1, 0, [block.data_ptr()])
            for _ in range(2):                                   # twice: the launch counter runs on across runs
                s.ingest_bands(t, f, None, value_dtype="float32")
                n = s.run_sharded(opts["iterations"], auto_terminate=opts["auto_terminate"], threshold=opts["threshold"])
                out, sweeps = s.result()
                assert list(sweeps.cpu().numpy()) == list(rs)
                close(np.moveaxis(out.cpu().numpy(), 0, -1), rr)
                assert n >= max(rs)
        finally:
            s.close()


# ---- 5. A1 on the device and the pipelined host path -----------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(256, 256), (100, 100), (37, 91), (2, 3), (1, 130), (300, 301)])
def test_device_mean_init_is_numpy_nanmean(shape, dtype):
    """load(fill=None) must start unknown cells at np.nanmean(band) taken in the band's dtype, bit for bit
    (VPint/MRP.py:73-77): NumPy's pairwise summation order is reproduced on the device."""
    import torch
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    B = 3
    stack = rng.uniform(0, 4000, shape + (B,)).astype(dtype)
    stack[rng.uniform(size=shape) < 0.3, :] = np.nan
    stack[0, 0, :] = np.nan                      # at least one unknown and one known cell
    stack[-1, -1, :] = 1234.5
    stack[..., 2] = np.round(stack[..., 2])
    s = Solver(2, B, shape[0], shape[1], nprob_inner=B, planes=False)
    try:
        t = torch.from_numpy(stack).cuda()
        s.load(t.permute(2, 0, 1).unsqueeze(0), fill=None)
        out, _ = s.result()
        got = out.cpu().numpy()
    finally:
        s.close()
    for b in range(B):
        band = stack[:, :, b]
        want = np.nanmean(band.copy())
        assert want.dtype == dtype
        filled = got[b][np.isnan(band)]
        assert filled.size > 0 and (filled == np.float64(want)).all(), (b, filled[0], want)
        assert np.array_equal(got[b][~np.isnan(band)], band[~np.isnan(band)].astype(np.float64))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_mean_init_band_interleaved_stack(dtype):
    """The band-interleaved fast paths (13 bands, several patches, H * W a power of two): same bit-for-bit
    np.nanmean per band, known cells copied unchanged."""
    import torch
    rng = np.random.default_rng(77)
    N, H, W, B = 3, 64, 128, 13
    stack = np.round(rng.uniform(0, 9000, (N, H, W, B))).astype(dtype)
    for i in range(N):
        stack[i, rng.uniform(size=(H, W)) < 0.1 + 0.3 * i, :] = np.nan
    stack[:, 0, 0, :] = np.nan
    stack[:, -1, -1, :] = 4321.0
    s = Solver(2, N * B, H, W, nprob_inner=B, planes=False)
    try:
        t = torch.from_numpy(stack).cuda()
        s.load(t.permute(0, 3, 1, 2), fill=None)
        out, _ = s.result()
        got = out.cpu().numpy().reshape(N, B, H, W)
    finally:
        s.close()
    for i in range(N):
        for b in range(B):
            band = stack[i, :, :, b]
            want = np.nanmean(band.copy())
            filled = got[i, b][np.isnan(band)]
            assert filled.size > 0 and (filled == np.float64(want)).all(), (i, b, filled[0], want)
            assert np.array_equal(got[i, b][~np.isnan(band)], band[~np.isnan(band)].astype(np.float64))


@pytest.mark.parametrize("dtype,storage", [(np.float32, "float64"), (np.float64, "float64"), (np.float32, "float32")])
def test_large_grid_float_stack_takes_its_mean_on_the_device(dtype, storage, monkeypatch):
    """Float stacks above the single-kernel limit (131 072 cells): the leaf / subtree kernels give the bit-identical
    np.nanmean per band, no host pass (VPint/MRP.py:73-77), then the ordinary load; the class entry point agrees."""
    import torch
    from vpint_b200 import vpint2 as V
    rng = np.random.default_rng(4242)
    N, H, W, B = 2, 397, 411, 3                         # 163 167 cells, not a multiple of 8
    stack = np.round(rng.uniform(0, 9000, (N, H, W, B))).astype(dtype)
    for i in range(N):
        stack[i, rng.uniform(size=(H, W)) < 0.15 + 0.2 * i, :] = np.nan
    monkeypatch.setattr(V, "_band_fill_values", lambda *a, **k: pytest.fail("host mean taken"))
    s = Solver(2, N * B, H, W, nprob_inner=B, planes=False, storage=storage)
    try:
        t = torch.from_numpy(stack).cuda()
        V._device_band_means(s, t)
        s.load(t.permute(0, 3, 1, 2), fill=None)
        out, _ = s.result(dtype=torch.float64)
        got = out.cpu().numpy().reshape(N, B, H, W)
    finally:
        s.close()
    for i in range(N):
        for b in range(B):
            band = stack[i, :, :, b]
            want = np.nanmean(band.copy())
            filled = got[i, b][np.isnan(band)]
            want = np.float64(np.float32(want)) if storage == "float32" else np.float64(want)
            assert filled.size > 0 and (filled == want).all(), (i, b, filled[0], want)
    if storage == "float64":
        feats = np.round(rng.uniform(100, 9000, (N, H, W, B))).astype(dtype)
        from vpint_b200.wp import wp_run_options
        res = V.solve_bands(stack, feats, wp_run_options(iterations=12, auto_terminate=False), out_layout="hwb")
        ref, _ = O.vpint2_run(stack[1], feats[1], iterations=12, auto_terminate=False)
        close(np.asarray(res)[1], ref)


def test_pipelined_host_stack_equals_single_solve():
    import torch
    from vpint_b200.vpint2 import _solve_bands_torch
    from vpint_b200.wp import wp_run_options
    rng = np.random.default_rng(606)
    vps = [make_patch(rng, 64, 64, 3) for _ in range(7)]
    tgt = torch.from_numpy(np.stack([v.target for v in vps])).pin_memory()
    fts = torch.from_numpy(np.stack([v.features for v in vps])).pin_memory()
    dev = torch.device("cuda", 0)
    opts = wp_run_options()
    out = torch.empty((7, 3, 64, 64), dtype=torch.float64).pin_memory()
    res, sweeps = _solve_bands_torch(tgt, fts, opts, "bhw", torch.float64, "float64", None, dev, out, None, chunk_patches=2)
    one, sweeps1 = _solve_bands_torch(tgt.cuda(), fts.cuda(), opts, "bhw", torch.float64, "float64", None, dev, None, None)
    assert torch.equal(sweeps, sweeps1)
    close(res.numpy(), one.numpy(), rtol=0)
    ref, ref_sweeps = O.vpint2_run(vps[5].target, vps[5].features)
    close(np.moveaxis(res.numpy()[5], 0, -1), ref)
    assert list(sweeps[5].numpy()) == list(ref_sweeps)


# ---- 6. in-loop options of WP_SMRP.run: identity priority, known-value bias, resistance -----------
def test_golden_wp2d_priority_and_resistance(golden, kblock):
    g = golden("wp2d")
    grid, feat = g["grid"], g["feat"]
    for beta, tag in ((0, "0"), (1, "1"), (2.5, "2p5")):
        m = WP_SMRP(grid.copy(), feat.copy())
        p, d = m.run(iterations=6, track_delta=True, prioritise_identity=True, priority_intensity=beta)
        close(p, g["pred_prio_" + tag]); closed(d, g["d_prio_" + tag])
        assert m.priority_grid.shape == grid.shape + (4,)
    w = O.weights_exact_2d(feat.astype(float).reshape(feat.shape[0], feat.shape[1], -1).copy())
    close(m.priority_grid, O.priority_planes(w, 2.5), rtol=1e-14)
    mu = float(g["res_mu"])
    m = WP_SMRP(grid.copy(), feat.copy())
    p, d = m.run(iterations=6, track_delta=True, resistance=True, epsilon=0.02, mu=mu)
    close(p, g["pred_res"]); closed(d, g["d_res"])
    assert not np.array_equal(p[~np.isnan(grid)], grid[~np.isnan(grid)]), "resistance moves known cells too"
    m = WP_SMRP(grid.copy(), feat.copy())
    p, d = m.run(iterations=5, track_delta=True, resistance=True, epsilon=0.01, mu=mu, prioritise_identity=True,
                 priority_intensity=1.5, clip_val=2.0 * mu)
    close(p, g["pred_res_prio"]); closed(d, g["d_res_prio"])
    p2 = m.run(iterations=2, resistance=True, epsilon=0.01, mu=mu)      # resume from a state whose known cells moved
    ref, _, _ = O.wp_run_2d(grid, g["pred_res_prio"].copy(), feat.astype(float).reshape(grid.shape + (-1,)).copy(),
                            iterations=2, resistance=True, epsilon=0.01, mu=mu)
    close(p2, ref)


def test_golden_known_value_bias(golden, kblock):
    g = golden("wp2d_edge")
    m = WP_SMRP(g["kvb_grid"].copy(), g["kvb_feat"].copy())
    p, d = m.run(iterations=5, track_delta=True, prioritise_identity=True, priority_intensity=1, known_value_bias=0.3)
    close(p, g["kvb_pred"]); closed(d, g["kvb_d"])


def test_vpint2_batch_with_priority_and_resistance(kblock):
    """The notebook's recommended call passes these through VPint2_interpolator.run(**kwargs)."""
    rng = np.random.default_rng(808)
    vp = make_patch(rng, 48, 48, 3)
    kw = dict(iterations=12, prioritise_identity=True, priority_intensity=1, known_value_bias=0.2, resistance=True,
              epsilon=0.01, mu=1500.0, clip_val=10000)
    got = vp.run(**kw)
    ref, _ = O.vpint2_run(vp.target, vp.features, **kw)
    close(got, ref)


def test_find_discounts_batched_trials_match_reference_rng_order():
    """SD_STMRP.find_discounts (VPint/SD_MRP.py:409-476) incl. its reset() quirk: epochs after the first start
    from the NaN-holding cube."""
    rng = np.random.default_rng(15)
    cube = smooth(rng, (10, 9, 5), 10, 30)
    cube[rng.uniform(size=cube.shape) < 0.15] = np.nan
    np.random.seed(21)
    m = SD_STMRP(cube.copy())
    best = m.find_discounts(4, 0.3, sub_iterations=30)
    np.random.seed(21)
    sub = cube.copy()
    for i in range(10):
        for j in range(9):
            for t in range(5):
                if not np.isnan(sub[i, j, t]) and np.random.rand() < 0.3:
                    sub[i, j, t] = np.nan
    best_ref, best_loss = (0.9, 0.9), np.inf
    state = O.init_pred(sub)
    for _ in range(4):
        gam, tau = np.random.rand(), np.random.rand()
        pred, _, _ = O.sd_run_3d(sub, state, gamma=gam, tau=tau, iterations=30)
        sel = ~np.isnan(cube) & np.isnan(sub)
        mae = np.add.accumulate(np.abs(pred[sel] - cube[sel]))[-1] / sel.sum()
        if mae < best_loss:
            best_ref, best_loss = (gam, tau), mae
        state = sub                          # reset(): the next epoch starts from the cube with its NaNs
    assert best == best_ref and (m.gamma, m.tau) == best_ref


# ---- 7. device-side loop: CUDA graph with a conditional WHILE node ----------------------------------------------
@pytest.mark.parametrize("k", [1, 4])
def test_graph_loop_mode_matches_chunked(k, monkeypatch):
    from vpint_b200 import engine, _native
    monkeypatch.setattr(solve_mod, "DEFAULT_K_BLOCK", k)
    rng = np.random.default_rng(4242)
    shape = (130, 203)
    base = smooth(rng, shape)
    grid = base + rng.normal(0, 20, shape)
    feat = np.stack([rng.uniform(0.7, 1.3) * base + rng.normal(0, 10, shape) for _ in range(3)], axis=-1)
    grid[discs(rng, shape, 6, 8, 40)] = np.nan
    ref, dref, n = O.wp_run_2d(grid, O.init_pred(grid), feat.copy())
    p0, d0 = WP_SMRP(grid.copy(), feat.copy()).run(track_delta=True)
    monkeypatch.setattr(engine, "DEFAULT_LOOP_MODE", _native.LOOP_GRAPH)
    p1, d1 = WP_SMRP(grid.copy(), feat.copy()).run(track_delta=True)
    assert len(d1) == n
    close(p1, ref); closed(d1, dref)
    close(p1, p0, rtol=0)                       # same kernels, same order: bit-identical to the host-driven loop
    p2, d2 = WP_SMRP(grid.copy(), feat.copy()).run(iterations=17, track_delta=True)
    ref2, dref2, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=17)
    close(p2, ref2); closed(d2, dref2)
    vps = [make_patch(rng, 64, 64, 3) for _ in range(3)]       # batched, per-problem stopping sweeps
    outs = run_patches(vps)
    for vp, out in zip(vps, outs):
        r, _ = O.vpint2_run(vp.target, vp.features)
        close(out, r)
    cube = smooth(rng, (12, 10, 6), 10, 30)
    cube[rng.uniform(size=cube.shape) < 0.2] = np.nan
    r3, d3, n3 = O.sd_run_3d(cube, O.init_pred(cube), gamma=0.9, tau=0.8)
    p3, dd3 = SD_STMRP(cube.copy(), gamma=0.9, tau=0.8).run(track_delta=True)
    assert len(dd3) == n3
    close(p3, r3)


# ---- 8. auto_adapt: the reference's host-side parameter search with device trials ---------------------------------
def test_golden_auto_adapt(golden):
    """Same picks as the reference under the same seed (incl. its quirks: 'sequential' always fails and falls back
    to the given parameters), same final result, same position of the global RNG afterwards."""
    g = golden("adapt")
    grid, feat = g["grid"], g["feat"]
    runs = {
        "rand": dict(auto_adaptation_epochs=5, auto_adaptation_strategy='random'),
        "rand_sub": dict(auto_adaptation_epochs=4, auto_adaptation_strategy='random',
                         auto_adaptation_subsample_strategy='random', auto_adaptation_proportion=0.3),
        "diff": dict(auto_adaptation_epochs=3, auto_adaptation_strategy='random',
                     auto_adaptation_subsample_strategy='max_diff'),
        "hill": dict(auto_adaptation_epochs=7, auto_adaptation_strategy='hill_climbing', auto_adaptation_max_iter=30),
        "seq": dict(auto_adaptation_epochs=2, auto_adaptation_strategy='sequential', auto_adaptation_max_iter=15),
    }
    for tag, kw in runs.items():
        np.random.seed(77)
        m = WP_SMRP(grid.copy(), feat.copy())
        p, d = m.run(iterations=25, track_delta=True, auto_adapt=True, prioritise_identity=True, resistance=True,
                     clip_val=5000.0, **kw)
        assert np.random.rand() == float(g["rng_" + tag]), tag
        close(p, g["pred_" + tag]); closed(d, g["d_" + tag])
    np.random.seed(78)
    p = WP_SMRP(grid.copy(), feat.copy()).run(iterations=10, auto_adapt=True, prioritise_identity=True,
                                              auto_adaptation_epochs=4)
    assert np.random.rand() == float(g["rng_beta_only"])
    close(p, g["pred_beta_only"])


def test_vpint2_auto_adapt_rng_semantics():
    rng = np.random.default_rng(909)
    vp = make_patch(rng, 32, 32, 2)
    kw = dict(iterations=8, auto_adapt=True, auto_adaptation_epochs=3, prioritise_identity=True, resistance=True,
              mu=1500.0, clip_val=10000)
    np.random.seed(5)
    got = vp.run(**kw)
    assert np.random.rand() == np.random.RandomState(5).rand()      # parent stream untouched, like after fork + join
    for b in range(2):                                              # every band saw the same stream
        np.random.seed(5)
        ref = WP_SMRP(vp.target[:, :, b], vp.features[:, :, b]).run(**kw)
        close(got[:, :, b], ref, rtol=0)
    np.random.seed(5)
    ser = vp.run_serial(**kw)
    np.random.seed(5)
    r0 = WP_SMRP(vp.target[:, :, 0], vp.features[:, :, 0]).run(**kw)
    r1 = WP_SMRP(vp.target[:, :, 1], vp.features[:, :, 1]).run(**kw)    # continues the stream
    close(ser[:, :, 0], r0.astype(np.float32), rtol=0)
    close(ser[:, :, 1], r1.astype(np.float32), rtol=0)


# ---- 9. edge cases of the single-feature (ratio) forms and degenerate grids ----------------------------------------
@pytest.mark.parametrize("shape", [(37, 91), (130, 203), (64, 5), (3, 300), (300, 301), (2, 2), (9, 1), (1, 9)])
def test_single_feature_ragged_shapes(shape, kblock):
    """One feature layer takes the ratio forms (resident kernel for k_block 0, ratio tiles for k_block 1): widths
    that are not multiples of the 4-cell segments / 64-element pitch, single rows and columns."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    base = smooth(rng, shape, 500, 2500)
    grid = np.round(base + rng.normal(0, 20, shape))
    feat = np.round(1.1 * base + rng.normal(0, 10, shape))
    hide = rng.uniform(size=shape) < 0.35
    hide.flat[0] = True
    hide.flat[-1] = False
    grid[hide] = np.nan
    ref, dref, n = O.wp_run_2d(grid, O.init_pred(grid), feat.reshape(shape + (1,)).astype(float).copy())
    p, d = WP_SMRP(grid.copy(), feat.copy()).run(track_delta=True)
    assert len(d) == n, (len(d), n)
    close(p, ref); closed(d, dref)
    assert np.array_equal(p[~hide], grid[~hide])
    ref, dref, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.reshape(shape + (1,)).astype(float).copy(), iterations=7,
                               clip_val=1800.0)
    p, d = WP_SMRP(grid.copy(), feat.copy()).run(iterations=7, track_delta=True, clip_val=1800.0)
    close(p, ref); closed(d, dref)


@pytest.mark.parametrize("shape", [(3, 4), (4, 8), (5, 260), (16, 16), (48, 64), (64, 256), (129, 128), (256, 256),
                                   (200, 512), (384, 384)])
def test_resident_fast_sweep_borders(shape):
    """Grids the resident kernel's branch-free sweep takes (W % 4 == 0, H >= 3, no clip_val) with unknown cells on
    every border and in every corner: neighbour counts 2 and 3 go through the multiply + fma-residual division,
    which must equal the reference's true division; row blocks are dealt round robin to the CTAs of a cluster."""
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    base = smooth(rng, shape, 500, 2500)
    grid = np.round(base + rng.normal(0, 20, shape))
    feat = np.round(1.1 * base + rng.normal(0, 10, shape))
    hide = rng.uniform(size=shape) < 0.3
    hide[0, :] = True; hide[-1, :] = True; hide[:, 0] = True; hide[:, -1] = True
    hide[shape[0] // 2, 1:-1] = False                             # something known in every column
    if shape[0] * shape[1] > 4000:
        hide |= discs(rng, shape, 3, 5, min(shape) // 3)
        hide[shape[0] // 2, 1:-1] = False
    grid[hide] = np.nan
    ref, dref, n = O.wp_run_2d(grid, O.init_pred(grid), feat.reshape(shape + (1,)).astype(float).copy())
    p, d = WP_SMRP(grid.copy(), feat.copy()).run(track_delta=True)
    assert len(d) == n, (len(d), n)
    close(p, ref); closed(d, dref)
    assert np.array_equal(p[~hide], grid[~hide])
    if shape[0] == shape[1]:                                      # SD_SMRP.run needs square grids (NumPy broadcast)
        ref, dref, n = O.sd_run_2d(grid, O.init_pred(grid), gamma=0.93, iterations=40, auto_terminate=False)
        p, d = SD_SMRP(grid.copy(), gamma=0.93).run(iterations=40, track_delta=True)
        close(p, ref); closed(d, dref)


def test_resident_feature_width_switch_on_one_solver():
    """The resident kernel keeps features on chip as floats when every feature of the batch is a float32 value and
    as doubles otherwise; the flag behind that choice is re-evaluated by every weights() call (also when a zero
    feature is replaced by 0.01, which is not a float32 value)."""
    import torch
    rng = np.random.default_rng(909)
    shape = (96, 128)
    base = smooth(rng, shape, 500, 2500)
    grid = np.round(base + rng.normal(0, 20, shape))
    grid[discs(rng, shape, 4, 6, 25)] = np.nan
    feats = {"float32 values": np.round(1.1 * base + rng.normal(0, 10, shape)),
             "fp64 values": 1.1 * base + rng.normal(0, 10, shape),
             "float32 values with zeros": np.round(1.1 * base + rng.normal(0, 10, shape))}
    feats["float32 values with zeros"][rng.uniform(size=shape) < 0.02] = 0.0
    s = Solver(2, 1, shape[0], shape[1], planes=True, k_block=0)
    try:
        for name in ("float32 values", "fp64 values", "float32 values with zeros", "float32 values"):
            f = feats[name]
            ref, _, n = O.wp_run_2d(grid, O.init_pred(grid), f.reshape(shape + (1,)).astype(float).copy())
            s.load(torch.from_numpy(grid).cuda().unsqueeze(0), fill=[float(np.nanmean(grid))])
            s.weights(torch.from_numpy(f.reshape((1,) + shape + (1,))).cuda(), method="exact")
            s.run(5000, auto_terminate=True)
            out, niter = s.result()
            assert int(niter[0]) == n, (name, int(niter[0]), n)
            close(out[0].cpu().numpy(), ref)
    finally:
        s.close()


def test_eo_wrapper_multiband_vpint():
    """multiband_VPint (VPint/utils/EO_wrapper.py:15-21) = per-band WP_SMRP runs, here as one batched solve."""
    from vpint_b200.utils.EO_wrapper import multiband_VPint
    rng = np.random.default_rng(515)
    shape = (60, 72)
    B = 4
    base = np.stack([smooth(rng, shape, 500, 2500) for _ in range(B)], axis=-1)
    target = np.round(base + rng.normal(0, 20, base.shape))
    feats = np.round(1.1 * base + rng.normal(0, 10, base.shape))
    target[discs(rng, shape, 3, 5, 18), :] = np.nan
    keep = target.copy()
    got = multiband_VPint(target, feats)
    assert got.dtype == target.dtype and got.shape == target.shape
    assert np.array_equal(target, keep, equal_nan=True)            # the caller's array is not touched
    for b in range(B):
        ref, _, _ = O.wp_run_2d(target[:, :, b], O.init_pred(target[:, :, b]), feats[:, :, b:b + 1].copy())
        close(got[:, :, b], ref)
    fixed = multiband_VPint(target, feats, iterations=9)
    for b in range(B):
        ref, _, _ = O.wp_run_2d(target[:, :, b], O.init_pred(target[:, :, b]), feats[:, :, b:b + 1].copy(), iterations=9)
        close(fixed[:, :, b], ref)
    pair = multiband_VPint(target, feats, iterations=5, method='cosine_similarity', min_gamma=0.2, max_gamma=1.05)
    for b in range(B):
        ref, _, _ = O.wp_run_2d(target[:, :, b], O.init_pred(target[:, :, b]), feats[:, :, b:b + 1].copy(), iterations=5,
                                method='cosine_similarity', min_gamma=0.2, max_gamma=1.05)
        close(pair[:, :, b], ref)


def test_degenerate_grids(kblock):
    rng = np.random.default_rng(12)
    full = smooth(rng, (20, 20), 10, 20)                       # nothing to interpolate: one sweep, delta 0
    p, d = WP_SMRP(full.copy(), full.copy()).run(track_delta=True)
    ref, dref, n = O.wp_run_2d(full, O.init_pred(full), full.reshape(20, 20, 1).copy())
    assert len(d) == n == 1 and d[0] == 0.0
    close(p, ref, rtol=0)
    p, d = SD_SMRP(full.copy()).run(track_delta=True)
    assert len(d) == 1 and np.array_equal(p, full)
    empty = np.full((12, 12), np.nan)                           # nothing known: NaN mean, NaN deltas, never "converges"
    ref, dref, n = O.sd_run_2d(empty, O.init_pred(empty), gamma=0.9, iterations=3)
    p, d = SD_SMRP(empty.copy()).run(iterations=3, track_delta=True)
    close(p, ref); closed(d, dref)
    one = full.copy()                                           # a single unknown cell in a corner
    one[0, 0] = np.nan
    ref, dref, n = O.wp_run_2d(one, O.init_pred(one), full.reshape(20, 20, 1).copy())
    p, d = WP_SMRP(one.copy(), full.copy()).run(track_delta=True)
    assert len(d) == n
    close(p, ref); closed(d, dref)


def test_resume_single_feature_paths(kblock):
    """run(a); run(b) continues from pred_grid on the single-feature paths as well; the ratio forms
    round-trip through value space between the two calls, so equality with one long run is to ~1 ulp."""
    rng = np.random.default_rng(33)
    shape = (70, 90)
    base = smooth(rng, shape, 500, 2500)
    grid = base + rng.normal(0, 20, shape)
    feat = 1.1 * base + rng.normal(0, 10, shape)
    grid[discs(rng, shape, 4, 6, 20)] = np.nan
    ref, _, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.reshape(shape + (1,)).copy(), iterations=30)
    m = WP_SMRP(grid.copy(), feat.copy())
    m.run(iterations=12)
    close(m.run(iterations=18), ref)
