"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

The reference (LaurensArp/VPint 0.2.4) targets numpy~=1.26 and calls
``np.product``, removed in NumPy 2; the one-line shim below (outside the
reference tree) is the only accommodation.  Every case stores its seeded inputs
and the reference's outputs (result grid, per-sweep delta vector), so the
fixtures are self-contained.
"""

import os
import sys
import warnings

import numpy as np

np.product = np.prod  # NumPy>=2 shim, see module docstring

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from VPint.SD_MRP import SD_SMRP, SD_STMRP      # noqa: E402
from VPint.WP_MRP import WP_SMRP, WP_STMRP      # noqa: E402
from VPint.VPint2 import VPint2_interpolator    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def smooth(rng, shape, lo=300.0, hi=2700.0, modes=4):
    """Strictly positive low-frequency field (keeps WP weights near 1)."""
    axes = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    acc = np.zeros(shape)
    for _ in range(modes):
        term = np.ones(shape)
        for ax in axes:
            term = term * np.sin(2 * np.pi * (rng.uniform(0.3, 1.5) * ax + rng.uniform()))
        acc += rng.uniform(0.5, 1.0) * term
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
    return lo + (hi - lo) * acc


def holes(rng, shape, frac):
    return rng.uniform(size=shape) < frac


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def case_sd2d():
    rng = np.random.default_rng(1001)
    base = smooth(rng, (20, 20), 10, 30) + rng.normal(0, 0.5, (20, 20))
    grid = base.copy()
    grid[holes(rng, grid.shape, 0.25)] = np.nan
    m = SD_SMRP(grid.copy(), gamma=0.9)
    pred_auto, d_auto = m.run(track_delta=True)
    m = SD_SMRP(grid.copy(), gamma=0.73)
    pred_fix, d_fix = m.run(iterations=7, track_delta=True)
    m = SD_SMRP(grid.copy(), gamma=0.9, init_strategy="zero")
    pred_zero, d_zero = m.run(iterations=3, track_delta=True)
    # resume semantics: run(3); run(4) == run(7)
    m = SD_SMRP(grid.copy(), gamma=0.73)
    m.run(iterations=3)
    pred_resume = m.run(iterations=4)
    # reset() leaves NaNs in pred_grid; the next run starts from them
    m = SD_SMRP(grid.copy(), gamma=0.9)
    m.reset()
    pred_reset, d_reset = m.run(iterations=4, track_delta=True)
    save("sd2d", grid=grid, pred_auto=pred_auto, d_auto=d_auto, pred_fix=pred_fix, d_fix=d_fix,
         pred_zero=pred_zero, d_zero=d_zero, pred_resume=pred_resume,
         pred_reset=pred_reset, d_reset=d_reset)


def case_wp2d():
    rng = np.random.default_rng(1002)
    H, W, F = 24, 18, 3
    base = smooth(rng, (H, W))
    grid = base + rng.normal(0, 20, (H, W))
    feat = np.stack([rng.uniform(0.7, 1.3) * base + rng.normal(0, 10, (H, W)) for _ in range(F)], axis=-1)
    feat = np.maximum(feat, 50.0)
    grid[holes(rng, grid.shape, 0.3)] = np.nan
    out = dict(grid=grid, feat=feat)

    m = WP_SMRP(grid.copy(), feat.copy())
    out["pred_auto"], out["d_auto"] = m.run(track_delta=True)
    m = WP_SMRP(grid.copy(), feat.copy())
    out["pred_fix"], out["d_fix"] = m.run(iterations=9, track_delta=True)
    clip = float(np.nanpercentile(grid, 60))
    m = WP_SMRP(grid.copy(), feat.copy())
    out["pred_clip"], out["d_clip"] = m.run(iterations=6, track_delta=True, clip_val=clip)
    out["clip_val"] = np.array(clip)
    for beta in (0, 1, 2.5):
        m = WP_SMRP(grid.copy(), feat.copy())
        p, d = m.run(iterations=6, track_delta=True, prioritise_identity=True, priority_intensity=beta)
        out["pred_prio_%s" % str(beta).replace(".", "p")] = p
        out["d_prio_%s" % str(beta).replace(".", "p")] = d
    mu = float(np.nanmean(grid))
    m = WP_SMRP(grid.copy(), feat.copy())
    out["pred_res"], out["d_res"] = m.run(iterations=6, track_delta=True, resistance=True, epsilon=0.02, mu=mu)
    out["res_mu"] = np.array(mu)
    m = WP_SMRP(grid.copy(), feat.copy())
    out["pred_res_prio"], out["d_res_prio"] = m.run(iterations=5, track_delta=True, resistance=True,
                                                     epsilon=0.01, mu=mu, prioritise_identity=True,
                                                     priority_intensity=1.5, clip_val=2.0 * mu)
    for method in ("cosine_similarity", "exact_inverse"):
        m = WP_SMRP(grid.copy(), feat.copy(), min_gamma=0.2, max_gamma=1.05)
        out["pred_" + method], out["d_" + method] = m.run(iterations=5, method=method, track_delta=True)
    # non-vectorised 'exact' is only reachable through predict_weight; the run()
    # branch for method='exact' is the vectorised one, already covered above.
    save("wp2d", **out)


def case_wp2d_edge():
    rng = np.random.default_rng(1003)
    out = {}
    # F=1 with zeros, a negative and a tiny feature; 2-D feature array accepted
    H, W = 9, 11
    grid = smooth(rng, (H, W), 1, 5)
    feat = smooth(rng, (H, W), 1, 5)
    feat[2, 3] = 0.0
    feat[0, 0] = 0.0
    feat[5, 5] = -0.0
    grid[holes(rng, grid.shape, 0.4)] = np.nan
    grid[0, 0] = np.nan
    grid[H - 1, W - 1] = np.nan
    m = WP_SMRP(grid.copy(), feat.copy())
    out["z_grid"], out["z_feat"] = grid, feat
    out["z_pred"], out["z_d"] = m.run(iterations=5, track_delta=True)
    out["z_feat_after"] = m.feature_grid.copy()
    # F=9 exercises NumPy's 8-accumulator pairwise mean over the feature axis
    H, W, F = 7, 8, 9
    grid = smooth(rng, (H, W), 100, 200)
    feat = np.stack([smooth(rng, (H, W), 100, 200) for _ in range(F)], axis=-1)
    grid[holes(rng, grid.shape, 0.35)] = np.nan
    m = WP_SMRP(grid.copy(), feat.copy())
    out["f9_grid"], out["f9_feat"] = grid, feat
    out["f9_pred"], out["f9_d"] = m.run(iterations=4, track_delta=True)
    # F=19: 2 full blocks of 8 plus a remainder of 3
    F = 19
    feat = np.stack([smooth(rng, (H, W), 100, 200) for _ in range(F)], axis=-1)
    m = WP_SMRP(grid.copy(), feat.copy())
    out["f19_feat"] = feat
    out["f19_pred"], out["f19_d"] = m.run(iterations=3, track_delta=True)
    # single column / single row grids: neighbour count hits 2 and 0
    for tag, shp in (("col", (6, 1)), ("row", (1, 6))):
        grid = smooth(rng, shp, 10, 20)
        feat = smooth(rng, shp, 10, 20)
        grid[2 if shp[0] > 1 else 0, 0 if shp[1] == 1 else 3] = np.nan
        m = WP_SMRP(grid.copy(), feat.copy())
        out[tag + "_grid"], out[tag + "_feat"] = grid, feat
        out[tag + "_pred"], out[tag + "_d"] = m.run(iterations=3, track_delta=True)
    # blow-up: huge feature ratios drive values to inf, then inf-inf=NaN is
    # skipped by nanmean
    H, W = 8, 8
    grid = smooth(rng, (H, W), 1e150, 2e150)
    feat = np.ones((H, W))
    feat[::2, :] = 1e160
    grid[2:6, 2:6] = np.nan
    m = WP_SMRP(grid.copy(), feat.copy())
    out["inf_grid"], out["inf_feat"] = grid, feat
    out["inf_pred"], out["inf_d"] = m.run(iterations=8, track_delta=True)
    m = WP_SMRP(grid.copy(), feat.copy())
    out["inf_pred_auto"], out["inf_d_auto"] = m.run(track_delta=True)
    # known_value_bias needs a square grid (it calls SD_SMRP)
    H = W = 10
    base = smooth(rng, (H, W))
    grid = base + rng.normal(0, 10, (H, W))
    feat = 1.1 * base + rng.normal(0, 5, (H, W))
    grid[holes(rng, grid.shape, 0.3)] = np.nan
    m = WP_SMRP(grid.copy(), feat.copy())
    out["kvb_grid"], out["kvb_feat"] = grid, feat
    out["kvb_pred"], out["kvb_d"] = m.run(iterations=5, track_delta=True, prioritise_identity=True,
                                          priority_intensity=1, known_value_bias=0.3)
    save("wp2d_edge", **out)


def case_st3d():
    rng = np.random.default_rng(1005)
    H, W, T = 8, 7, 5
    cube = smooth(rng, (H, W, T), 10, 30) + rng.normal(0, 0.3, (H, W, T))
    cube[holes(rng, cube.shape, 0.3)] = np.nan
    out = dict(cube=cube)
    m = SD_STMRP(cube.copy(), gamma=0.9, tau=0.8)
    out["sd_auto"], out["sd_d_auto"] = m.run(track_delta=True)
    m = SD_STMRP(cube.copy(), gamma=0.6, tau=0.95)
    out["sd_fix"], out["sd_d_fix"] = m.run(iterations=6, track_delta=True)
    feat = np.stack([smooth(rng, (H, W), 1, 3) for _ in range(2)], axis=-1)
    feat[3, 3, :] = 0.0
    feat[1, 2, 0] = 0.0
    out["feat"] = feat
    m = WP_STMRP(cube.copy(), feat.copy())
    out["wp_exact"], out["wp_d_exact"] = m.run(iterations=5, method="exact", track_delta=True)
    m = WP_STMRP(cube.copy(), feat.copy())
    out["wp_cos"], out["wp_d_cos"] = m.run(iterations=4, method="cosine_similarity", track_delta=True)
    m = WP_STMRP(cube.copy(), feat.copy())
    out["wp_auto"], out["wp_d_auto"] = m.run(method="exact", track_delta=True)
    # T == 1 cube (dim_check turns a 2-D grid into this)
    flat = cube[:, :, 0].copy()
    m = SD_STMRP(flat.copy(), gamma=0.9, tau=0.9)
    out["flat"] = flat
    out["sd_flat"], out["sd_d_flat"] = m.run(iterations=3, track_delta=True)
    save("st3d", **out)


def case_vpint2():
    rng = np.random.default_rng(1004)
    H, W, B = 16, 14, 3
    target = np.stack([np.round(smooth(rng, (H, W))) for _ in range(B)], axis=-1).astype(np.uint16)
    features = np.stack([np.round(1.1 * target[:, :, b] + rng.normal(0, 10, (H, W))) for b in range(B)],
                        axis=-1).astype(np.uint16)
    mask = np.zeros((H, W))
    yy, xx = np.mgrid[0:H, 0:W]
    mask[(yy - 6) ** 2 + (xx - 5) ** 2 < 16] = 100
    mask[(yy - 12) ** 2 + (xx - 11) ** 2 < 6] = 100
    vp = VPint2_interpolator(target, features, mask=mask, threshold=10)
    pred_serial = vp.run_serial()
    pred_par = vp.run()
    sweeps = []
    for b in range(B):
        m = WP_SMRP(vp.target[:, :, b], vp.features[:, :, b])
        _, d = m.run(track_delta=True)
        sweeps.append(len(d))
    pred_fix = vp.run_serial(iterations=5)
    pred_clip = vp.run(iterations=4, clip_val=1500.0)
    save("vpint2", target=target, features=features, mask=mask, stored_target=vp.target,
         pred_serial=pred_serial, pred_par=np.ascontiguousarray(pred_par), sweeps=np.array(sweeps),
         pred_fix=pred_fix, pred_clip=np.ascontiguousarray(pred_clip))


def case_vpint2_ctor():
    rng = np.random.default_rng(1006)
    out = {}
    for tag, (H, W) in (("wide", (12, 20)), ("tall", (20, 12))):
        B = 2
        target = rng.uniform(100, 200, (H, W, B))
        features = rng.uniform(100, 200, (H, W, B))
        mask = np.zeros((H, W))
        mask[3, 4] = 1.0
        mask[H - 2, W - 3] = 0.5
        mask[H // 2, W - 1] = 0.1      # below the 0.2 default threshold
        vp = VPint2_interpolator(target, features, mask=mask, buffer_mask=True, mask_buffer_size=3,
                                 mask_buffer_passes=2)
        out[tag + "_target"], out[tag + "_features"], out[tag + "_mask"] = target, features, mask
        out[tag + "_stored"] = vp.target
        vp1 = VPint2_interpolator(target, features, mask=mask, clip_target=150, clip_features=(110.0, 190.0))
        out[tag + "_stored_clip"] = vp1.target
        out[tag + "_features_clip"] = vp1.features
    tb = np.moveaxis(out["wide_target"], -1, 0)
    fb = np.moveaxis(out["wide_features"], -1, 0)
    vp = VPint2_interpolator(tb, fb, mask=out["wide_mask"], bands_first=True, dtype=np.float64, threshold=0.05)
    out["bf_stored"] = vp.target
    save("vpint2_ctor", **out)


def case_metrics():
    """MRP.r_squared / MAE / RMSE / PSNR and SMRP.mean_absolute_error (VPint/MRP.py:116-233,284-325) on a fixed
    prediction: pred_grid is set by hand, so the fixture pins the formulas, not a solve."""
    rng = np.random.default_rng(1010)
    out = {}
    for tag, shape in (("2d", (24, 24)), ("3d", (10, 12, 5))):
        true = smooth(rng, shape, 10, 30) + rng.normal(0, 0.5, shape)
        grid = true.copy()
        grid[holes(rng, shape, 0.3)] = np.nan
        true_nan = true.copy()
        true_nan[holes(rng, shape, 0.05)] = np.nan           # the metrics replace NaNs of the ground truth by its nanmean
        m = SD_SMRP(grid.copy(), gamma=0.9) if len(shape) == 2 else SD_STMRP(grid.copy(), auto_timesteps=False)
        pred = np.where(np.isnan(grid), true + rng.normal(0, 1.0, shape), grid)
        m.pred_grid = pred.copy()
        out.update({"grid_" + tag: grid, "true_" + tag: true_nan, "pred_" + tag: pred,
                    "r2_" + tag: np.float64(m.r_squared(true_nan.copy())), "mae_" + tag: np.float64(m.MAE(true_nan.copy())),
                    "rmse_" + tag: np.float64(m.RMSE(true_nan.copy())), "psnr_" + tag: np.float64(m.PSNR(true_nan.copy()))})
        if len(shape) == 2:
            mae, eg = m.mean_absolute_error(true.copy(), scale_factor=1, gridded=True)
            out.update({"truefull_2d": true, "smae_2d": np.float64(mae), "smae_grid_2d": eg,
                        "smae2_2d": np.float64(m.mean_absolute_error(true.copy(), scale_factor=2))})
    save("metrics", **out)


def case_adapt():
    """WP_SMRP.run(auto_adapt=True): the host-side search of VPint/WP_MRP.py:212-239, 429-759 with seeded
    global RNG.  Written to its own file so that the earlier fixtures stay untouched."""
    rng = np.random.default_rng(1010)
    shape = (26, 26)
    base = smooth(rng, shape)
    grid = base + rng.normal(0, 15, shape)
    feat = 1.1 * base + rng.normal(0, 8, shape)
    grid[holes(rng, shape, 0.25)] = np.nan
    out = {"grid": grid, "feat": feat}
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
        out["pred_" + tag], out["d_" + tag] = p, d
        out["rng_" + tag] = np.array(np.random.rand())        # where the global RNG stands afterwards
    np.random.seed(78)
    m = WP_SMRP(grid.copy(), feat.copy())
    out["pred_beta_only"] = m.run(iterations=10, auto_adapt=True, prioritise_identity=True, auto_adaptation_epochs=4)
    out["rng_beta_only"] = np.array(np.random.rand())
    save("adapt", **out)


if __name__ == "__main__":
    case_vpint2_ctor()
    case_sd2d()
    case_wp2d()
    case_wp2d_edge()
    case_st3d()
    case_vpint2()
    case_adapt()
    case_metrics()
