"""Host-side mirror of the reference interface (constructors, argument handling,
error behaviour) against reference-generated fixtures.  No GPU needed: run() is
only checked to fail loudly without a device."""

import numpy as np
import pytest

import vpint_b200
from vpint_b200 import SD_SMRP, SD_STMRP, VPint2_interpolator, VPintError, WP_SMRP, WP_STMRP
from vpint_b200.wp import wp_run_options

try:
    import torch
    HAS_CUDA = torch.cuda.is_available()
except Exception:  # pragma: no cover
    HAS_CUDA = False


def same(a, b):
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(a, b, equal_nan=True)


def test_reference_module_names_importable():
    from vpint_b200.SD_MRP import SD_SMRP as A
    from vpint_b200.WP_MRP import WP_SMRP as B, VPintError as E
    from vpint_b200.VPint2 import VPint2_interpolator as V
    from vpint_b200.MRP import MRP, SMRP, STMRP
    assert A is SD_SMRP and B is WP_SMRP and V is VPint2_interpolator and E is VPintError
    assert issubclass(SMRP, MRP) and issubclass(STMRP, MRP)


def test_init_strategies(golden):
    g = golden("sd2d")["grid"]
    m = SD_SMRP(g.copy())
    same(m.original_grid, g)
    assert not np.isnan(m.pred_grid).any()
    assert m.pred_grid[np.isnan(g)][0] == np.nanmean(g)
    m = SD_SMRP(g.copy(), init_strategy="zero")
    assert (m.pred_grid[np.isnan(g)] == 0).all()
    np.random.seed(3)
    m = SD_SMRP(g.copy(), init_strategy="random")
    np.random.seed(3)
    ref = np.random.normal(np.nanmean(g), np.nanstd(g), size=int(np.isnan(g).sum()))
    assert np.array_equal(m.pred_grid[np.isnan(g)], ref)
    with pytest.raises(VPintError):
        SD_SMRP(g.copy(), init_strategy="bogus")
    m = SD_SMRP(g.copy())
    m.reset()
    assert m.pred_grid is m.original_grid


def test_float32_mean_init_keeps_dtype():
    g = np.random.default_rng(0).uniform(1, 2, (9, 9)).astype(np.float32)
    g[2, 3] = np.nan
    m = WP_SMRP(g, np.ones((9, 9), dtype=np.float32))
    assert m.pred_grid.dtype == np.float32 and m.feature_grid.dtype == np.float64
    assert m.pred_grid[2, 3] == np.nanmean(g)
    assert m.feature_grid.shape == (9, 9, 1)


def test_mrp_apply_mask_reference_quirk():
    # VPint/MRP.py:100-113 masks cells whose VALUE is 1 and edits the caller's array
    g = np.array([[1.0, 2.0], [3.0, 1.0]])
    mask = np.zeros((2, 2))
    m = SD_SMRP(g, mask=mask)
    assert np.isnan(m.original_grid[0, 0]) and np.isnan(m.original_grid[1, 1])
    assert np.isnan(g[0, 0])
    with pytest.raises(VPintError):
        SD_SMRP(np.ones((2, 2)), mask=np.zeros((2, 1, 2)))


def test_wp_constructor_errors():
    with pytest.raises(VPintError):
        WP_SMRP(np.ones((4, 4)), np.ones((4, 5)))
    with pytest.raises(VPintError):
        WP_SMRP(np.ones((4, 4, 2)), np.ones((4, 4)))
    with pytest.raises(VPintError):
        WP_SMRP(np.ones((4, 4)), np.ones((4, 4, 1, 1)))
    m = WP_SMRP(np.ones((4, 4, 1)), np.ones((4, 4)))
    assert m.original_grid.shape == (4, 4)


def test_run_option_handling():
    o = wp_run_options()
    assert o["iterations"] == 5000 and o["auto_terminate"] and o["method"] == "exact"
    o = wp_run_options(iterations=7)
    assert o["iterations"] == 7 and not o["auto_terminate"]
    for kw in (dict(method="predict"), dict(method="nope"), dict(save_gif=True), dict(store_gif_source=True)):
        with pytest.raises(VPintError):
            wp_run_options(**kw)
    o = wp_run_options(auto_adapt=True, prioritise_identity=True, resistance=True, epsilon=0.2)
    assert o["auto_adapt"] and o["prioritise_identity"] and o["resistance"] and o["epsilon"] == 0.2
    assert o["auto_adaptation_epochs"] == 100 and o["auto_adaptation_subsample_strategy"] == "max_contrast"
    with pytest.raises(TypeError):
        wp_run_options(no_such_option=1)


def test_sd_nonsquare_raises_like_numpy():
    m = SD_SMRP(np.ones((4, 5)))
    with pytest.raises(ValueError):
        m.run(1)


def test_zero_iterations_returns_start_state_without_gpu():
    g = np.ones((5, 5))
    g[1, 1] = np.nan
    m = SD_SMRP(g)
    out = m.run(iterations=0)
    assert out is m.pred_grid and m.run_state
    p, d = m.run(iterations=0, track_delta=True)
    assert d.shape == (0,)
    c = np.ones((3, 3, 2))
    assert SD_STMRP(c).run(iterations=0).shape == (3, 3, 2)


@pytest.mark.skipif(HAS_CUDA, reason="checks the no-GPU failure mode")
def test_run_fails_loudly_without_cuda():
    g = np.ones((5, 5))
    g[1, 1] = np.nan
    with pytest.raises(VPintError, match="no CPU fallback"):
        SD_SMRP(g).run(3)
    with pytest.raises(VPintError, match="no CPU fallback"):
        WP_SMRP(g, np.ones((5, 5))).run(3)
    with pytest.raises(VPintError, match="no CPU fallback"):
        VPint2_interpolator(np.ones((4, 4, 2)), np.ones((4, 4, 2))).run()


def test_stmrp_dim_check_and_timesteps():
    m = SD_STMRP(np.ones((4, 3)))
    assert m.original_grid.shape == (4, 3, 1)
    data = {"2020-01-01 00:00:00": np.ones((2, 2)), "2020-01-03 00:00:00": 2 * np.ones((2, 2)),
            "2020-01-04 00:00:00": 3 * np.ones((2, 2))}
    m = SD_STMRP(data, auto_timesteps=True)
    assert m.original_grid.shape == (2, 2, 4) and m.true_time_indices == [0, 2, 3]
    assert np.isnan(m.original_grid[:, :, 1]).all() and not np.isnan(m.pred_grid).any()
    w = WP_STMRP(np.ones((3, 3, 2)), np.ones((3, 3, 1)))
    with pytest.raises(VPintError):
        w.run(iterations=1)          # default method='predict'
    with pytest.raises(VPintError):
        w.run(iterations=1, method="nope")


def test_vpint2_constructor_matches_reference(golden):
    g = golden("vpint2_ctor")
    for tag in ("wide", "tall"):
        vp = VPint2_interpolator(g[tag + "_target"], g[tag + "_features"], mask=g[tag + "_mask"], buffer_mask=True,
                                 mask_buffer_size=3, mask_buffer_passes=2)
        same(vp.target, g[tag + "_stored"])
        assert vp.features.dtype == np.float32
        vp = VPint2_interpolator(g[tag + "_target"], g[tag + "_features"], mask=g[tag + "_mask"], clip_target=150,
                                 clip_features=(110.0, 190.0))
        same(vp.target, g[tag + "_stored_clip"])
        same(vp.features, g[tag + "_features_clip"])
    vp = VPint2_interpolator(np.moveaxis(g["wide_target"], -1, 0), np.moveaxis(g["wide_features"], -1, 0),
                             mask=g["wide_mask"], bands_first=True, dtype=np.float64, threshold=0.05)
    same(vp.target, g["bf_stored"])
    v = golden("vpint2")
    vp = VPint2_interpolator(v["target"], v["features"], mask=v["mask"], threshold=10)
    same(vp.target, v["stored_target"])
    with pytest.raises(AssertionError):
        VPint2_interpolator(np.ones((4, 4, 2)), np.ones((4, 4, 3)))


def test_package_exports():
    for name in vpint_b200.__all__:
        assert hasattr(vpint_b200, name)


def test_contrast_map_and_subsampling_match_reference_semantics():
    """Host-side pieces of auto_adapt (VPint/WP_MRP.py:441-492, 893-940): contrast map incl. the reference's
    one-short neighbour slices, and which cells the three sub-sampling strategies hide."""
    from vpint_b200.adapt import _subsample, contrast_map
    rng = np.random.default_rng(3)
    g = rng.uniform(1, 9, (7, 6))
    g[2, 3] = np.nan
    c = contrast_map(g)
    assert c.shape == g.shape and np.nanmin(c) == 0.0 and np.nanmax(c) == 1.0
    # hand evaluation of one interior cell: its right / down planes are shifted by the short slices
    i, j = 3, 2
    up, right, down, left = g[i - 1, j], g[i, j + 1], g[i + 1, j], g[i, j - 1]
    raw = (abs(up - g[i, j]) + abs(right - g[i, j]) + abs(down - g[i, j]) + abs(left - g[i, j])) / 4
    nb = np.zeros((7, 6, 4)); nb[1:-1, :, 0] = g[0:-2]; nb[:, 0:-2, 1] = g[:, 1:-1]; nb[0:-2, :, 2] = g[1:-1]; nb[:, 1:-1, 3] = g[:, 0:-2]
    cnt = np.full(g.shape, 4.0); cnt[:, 0] -= 1; cnt[:, -1] -= 1; cnt[0] -= 1; cnt[-1] -= 1
    full = np.nansum(np.abs(nb - g[:, :, None]), axis=-1) / cnt
    assert np.isclose(full[i, j], raw)
    m = WP_SMRP(g.copy(), rng.uniform(1, 2, (7, 6)))
    np.random.seed(1)
    s_rand = _subsample(m, 'random', 0.5)
    np.random.seed(1)
    assert np.array_equal(np.isnan(s_rand), np.isnan(g) | (np.random.rand(42).reshape(7, 6) < 0.5))
    for strat in ('max_contrast', 'max_diff'):
        s = _subsample(m, strat, 0.25)
        assert np.isnan(s).sum() == 1 + int(0.25 * 41)
    with pytest.raises(VPintError):
        _subsample(m, 'bogus', 0.5)


def test_pipeline_wait_policy(monkeypatch):
    """Pipeline threads spin in their device waits only while every one of them can have a core
    (vpint_b200/vpint2.py:_pipeline_blocking_waits); VPINT_PIPELINE_BLOCKING overrides."""
    import os
    from vpint_b200 import vpint2 as V
    monkeypatch.delenv("VPINT_PIPELINE_BLOCKING", raising=False)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(16)), raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert V._pipeline_blocking_waits(6) is False
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    assert V._pipeline_blocking_waits(6) is True          # 14 threads want more than half of 16 cores
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(32)), raising=False)
    assert V._pipeline_blocking_waits(6) is False
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "4")
    assert V._pipeline_blocking_waits(6) is True
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert V._pipeline_blocking_waits(6) is True
    monkeypatch.setenv("VPINT_PIPELINE_BLOCKING", "0")
    assert V._pipeline_blocking_waits(6) is False
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    monkeypatch.setenv("VPINT_PIPELINE_BLOCKING", "1")
    assert V._pipeline_blocking_waits(6) is True


def test_bench_numa_binding_degrades_gracefully():
    """bench.bind_to_gpu_node must never raise: without NVML (this container) it reports 'unbound'."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    msg = bench.bind_to_gpu_node(0)
    assert isinstance(msg, str) and msg
    if msg.startswith("unbound"):
        assert os.sched_getaffinity(0) == before
    else:                                         # a box with NVML: restore for the tests that follow
        os.sched_setaffinity(0, before)


def test_eo_wrapper_apply_cloud_mask_matches_reference_loop():
    """VPint/utils/EO_wrapper.py:23-37 restated as the reference's pixel loop."""
    from vpint_b200.utils.EO_wrapper import apply_cloud_mask
    rng = np.random.default_rng(4)
    target = rng.uniform(1, 2, (9, 11, 4)).astype(np.float32)
    mask = rng.uniform(0, 1, (9, 11, 2))
    mask[3, 5, 0] = np.nan

    def loop(target_img, mask, threshold=None):
        cloud_img = target_img.copy()
        for i in range(target_img.shape[0]):
            for j in range(target_img.shape[1]):
                hit = np.isnan(mask[i, j, 0]) if threshold is None else (mask[i, j, 0] > threshold)
                if hit:
                    cloud_img[i, j, :] = np.nan
        return cloud_img

    for thr in (None, 0.35, 2.0):
        got = apply_cloud_mask(target, mask, thr)
        same(got, loop(target, mask, thr))
        assert got is not target and not np.isnan(target).any()


@pytest.mark.skipif(HAS_CUDA, reason="checks the no-GPU failure mode")
def test_eo_wrapper_multiband_fails_loudly_without_cuda():
    from vpint_b200.utils.EO_wrapper import multiband_VPint
    t = np.ones((5, 5, 2))
    t[1, 1, :] = np.nan
    with pytest.raises(VPintError, match="no CPU fallback"):
        multiband_VPint(t, np.ones((5, 5, 2)))


def test_metrics_match_reference(golden):
    """MRP.r_squared / MAE / RMSE / PSNR / SMRP.mean_absolute_error against values produced by the reference
    (tests/golden/make_golden.py case_metrics; VPint/MRP.py:116-233,284-325)."""
    g = golden("metrics")
    for tag in ("2d", "3d"):
        m = SD_SMRP(g["grid_" + tag].copy(), gamma=0.9) if tag == "2d" else SD_STMRP(g["grid_" + tag].copy(), auto_timesteps=False)
        m.pred_grid = g["pred_" + tag].copy()
        t = g["true_" + tag]
        assert np.isclose(m.r_squared(t.copy()), g["r2_" + tag], rtol=1e-12, atol=0)
        assert np.isclose(m.MAE(t.copy()), g["mae_" + tag], rtol=1e-12, atol=0)
        assert np.isclose(m.RMSE(t.copy()), g["rmse_" + tag], rtol=1e-12, atol=0)
        assert np.isclose(m.PSNR(t.copy()), g["psnr_" + tag], rtol=1e-12, atol=0)
    m = SD_SMRP(g["grid_2d"].copy(), gamma=0.9)
    m.pred_grid = g["pred_2d"].copy()
    mae, eg = m.mean_absolute_error(g["truefull_2d"].copy(), scale_factor=1, gridded=True)
    assert np.isclose(mae, g["smae_2d"], rtol=1e-12, atol=0) and np.allclose(eg, g["smae_grid_2d"], rtol=1e-12, atol=0)
    assert np.isclose(m.mean_absolute_error(g["truefull_2d"].copy(), scale_factor=2), g["smae2_2d"], rtol=1e-12, atol=0)


def test_balanced_row_bounds():
    """Row shards dealt by work: boundaries on the cumulative per-row weight, every rank at least min_rows rows."""
    from vpint_b200.sharded import RowShard, balanced_bounds
    rng = np.random.default_rng(5)
    w = np.zeros(1000)
    w[100:300] = 50.0                                  # one cloud bank
    w[800:820] = 200.0
    w += 1.0
    for world in (1, 2, 3, 8):
        b = balanced_bounds(w, world, min_rows=4)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == world + 1 and all(x1 - x0 >= 4 for x0, x1 in zip(b, b[1:]))
        work = [w[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(work) <= 1.25 * (w.sum() / world) + w.max()
        shards = [RowShard(1000, world, r, halo=1, bounds=b) for r in range(world)]
        assert [s.r0 for s in shards] == b[:-1] and shards[-1].r1 == 1000
        assert all(s.own_rows == b[r + 1] - b[r] for r, s in enumerate(shards))
    assert balanced_bounds(np.zeros(10), 3) == [0, 4, 7, 10] or len(balanced_bounds(np.zeros(10), 3)) == 4
    with pytest.raises(VPintError):
        RowShard(10, 2, 0, 1, bounds=[0, 5, 9])
    with pytest.raises(VPintError):
        balanced_bounds(np.ones(3), 4)


def test_interpolator_keeps_raw_rasters_and_materialises_on_demand(golden):
    """The lazy constructor: uint16 / float rasters are kept as they came; .target / .features are built on the host
    with the reference's steps only when somebody reads them (fixture: tests/golden/vpint2.npz)."""
    g = golden("vpint2")
    vp = VPint2_interpolator(g["target"], g["features"], mask=g["mask"], threshold=10)
    assert vp._raw is not None and vp._target is None
    assert vp._raw[0].dtype == g["target"].dtype and vp._raw[0] is not g["target"]      # a copy, not a reference
    t = vp.target
    same(t, g["stored_target"])
    assert vp._target is t and vp.features.dtype == np.float32
    vp.features = vp.features * 2
    assert vp._raw is None, "assigning switches to the host-built arrays"
    # inputs the device constructor does not take are built on the host at once
    vp2 = VPint2_interpolator(g["target"].astype(np.int32), g["features"].astype(np.int32), mask=g["mask"], threshold=10)
    assert vp2._raw is None
    same(vp2.target, g["stored_target"])


def test_solve_raw_fails_loudly_without_a_device():
    if HAS_CUDA:
        pytest.skip("needs a box without CUDA")
    from vpint_b200.vpint2 import solve_raw
    t = np.ones((1, 16, 16, 2), dtype=np.uint16)
    with pytest.raises(VPintError):
        solve_raw(t, t, None)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pairwise_tree_restatement_is_numpy(dtype):
    """oracle/numpy_pairwise.py -- the leaf / tree decomposition the device mean init implements -- against NumPy
    itself, bit for bit: sums of ragged sizes around every split rule, and np.nanmean with holes."""
    from oracle import numpy_pairwise as PW
    rng = np.random.default_rng(11)
    sizes = [1, 5, 7, 8, 9, 15, 16, 64, 127, 128, 129, 136, 137, 255, 256, 257, 263, 1000, 1023, 1025, 4097, 10980, 65536, 70001]
    for n in sizes:
        a = (rng.uniform(0, 9000, n)).astype(dtype)
        assert PW.pairwise_sum(a) == np.add.reduce(a), n
        # every element belongs to exactly one leaf, leaves start at multiples of 8 and hold 64..128 elements (n > 128)
        lv = PW.leaves(n)
        assert lv[0][0] == 0 and sum(m for _, m in lv) == n
        assert all(lo1 + m1 == lo2 for (lo1, m1), (lo2, _) in zip(lv, lv[1:]))
        if n > 128:
            assert all(lo % 8 == 0 and 64 <= m <= 128 for lo, m in lv)
        holes = a.copy()
        holes[rng.uniform(size=n) < 0.3] = np.nan
        if np.isnan(holes).all():
            continue
        got, want = PW.nanmean(holes), np.nanmean(holes)
        assert got.dtype == want.dtype and got == want, (n, got, want)
    grid = rng.uniform(0, 9000, (397, 411)).astype(dtype)
    grid[rng.uniform(size=grid.shape) < 0.2] = np.nan
    assert PW.nanmean(grid) == np.nanmean(grid)
