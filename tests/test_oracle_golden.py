"""The oracle (oracle/vpint_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  Bit-for-bit:
the oracle issues the same NumPy primitives as the reference."""

import warnings

import numpy as np
import pytest

from oracle import vpint_oracle as O

warnings.filterwarnings("ignore")


def same(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b, equal_nan=True), float(np.nanmax(np.abs(a - b)))


def test_sd2d(golden):
    g = golden("sd2d")
    grid = g["grid"]
    p, d, n = O.sd_run_2d(grid, O.init_pred(grid), gamma=0.9)
    same(p, g["pred_auto"]); same(d, g["d_auto"]); assert n == len(g["d_auto"])
    p, d, n = O.sd_run_2d(grid, O.init_pred(grid), gamma=0.73, iterations=7)
    same(p, g["pred_fix"]); same(d, g["d_fix"])
    p, d, n = O.sd_run_2d(grid, O.init_pred(grid, "zero"), gamma=0.9, iterations=3)
    same(p, g["pred_zero"]); same(d, g["d_zero"])
    p, _, _ = O.sd_run_2d(grid, O.init_pred(grid), gamma=0.73, iterations=3)
    p, _, _ = O.sd_run_2d(grid, p, gamma=0.73, iterations=4)
    same(p, g["pred_resume"]); same(p, g["pred_fix"])
    p, d, _ = O.sd_run_2d(grid, grid.copy(), gamma=0.9, iterations=4)   # after reset()
    same(p, g["pred_reset"]); same(d, g["d_reset"])


def test_sd2d_nonsquare_raises():
    grid = np.ones((4, 5))
    with pytest.raises(ValueError):
        O.sd_run_2d(grid, grid.copy())


def test_wp2d(golden):
    g = golden("wp2d")
    grid, feat = g["grid"], g["feat"]
    p, d, n = O.wp_run_2d(grid, O.init_pred(grid), feat.copy())
    same(p, g["pred_auto"]); same(d, g["d_auto"]); assert n == 14
    p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=9)
    same(p, g["pred_fix"]); same(d, g["d_fix"])
    p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=6, clip_val=float(g["clip_val"]))
    same(p, g["pred_clip"]); same(d, g["d_clip"])
    for beta, tag in ((0, "0"), (1, "1"), (2.5, "2p5")):
        p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=6,
                              prioritise_identity=True, priority_intensity=beta)
        same(p, g["pred_prio_" + tag]); same(d, g["d_prio_" + tag])
    mu = float(g["res_mu"])
    p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=6, resistance=True, epsilon=0.02, mu=mu)
    same(p, g["pred_res"]); same(d, g["d_res"])
    p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=5, resistance=True, epsilon=0.01,
                          mu=mu, prioritise_identity=True, priority_intensity=1.5, clip_val=2.0 * mu)
    same(p, g["pred_res_prio"]); same(d, g["d_res_prio"])
    for method in ("cosine_similarity", "exact_inverse"):
        p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), feat.copy(), iterations=5, method=method,
                              min_gamma=0.2, max_gamma=1.05)
        same(p, g["pred_" + method]); same(d, g["d_" + method])


def test_wp2d_edge(golden):
    g = golden("wp2d_edge")
    f = g["z_feat"].astype(float).reshape(g["z_feat"].shape + (1,)).copy()
    p, d, _ = O.wp_run_2d(g["z_grid"], O.init_pred(g["z_grid"]), f, iterations=5)
    same(p, g["z_pred"]); same(d, g["z_d"]); same(f, g["z_feat_after"])
    p, d, _ = O.wp_run_2d(g["f9_grid"], O.init_pred(g["f9_grid"]), g["f9_feat"].copy(), iterations=4)
    same(p, g["f9_pred"]); same(d, g["f9_d"])
    p, d, _ = O.wp_run_2d(g["f9_grid"], O.init_pred(g["f9_grid"]), g["f19_feat"].copy(), iterations=3)
    same(p, g["f19_pred"]); same(d, g["f19_d"])
    for tag in ("col", "row"):
        grid = g[tag + "_grid"]
        f = g[tag + "_feat"].reshape(grid.shape + (1,)).copy()
        p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), f, iterations=3)
        same(p, g[tag + "_pred"]); same(d, g[tag + "_d"])
    grid = g["inf_grid"]
    f = g["inf_feat"].reshape(grid.shape + (1,))
    p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), f.copy(), iterations=8)
    same(p, g["inf_pred"]); same(d, g["inf_d"])
    p, d, n = O.wp_run_2d(grid, O.init_pred(grid), f.copy())
    same(p, g["inf_pred_auto"]); same(d, g["inf_d_auto"]); assert n == 3
    grid = g["kvb_grid"]
    f = g["kvb_feat"].reshape(grid.shape + (1,)).copy()
    p, d, _ = O.wp_run_2d(grid, O.init_pred(grid), f, iterations=5, prioritise_identity=True,
                          priority_intensity=1, known_value_bias=0.3)
    same(p, g["kvb_pred"]); same(d, g["kvb_d"])


def test_st3d(golden):
    g = golden("st3d")
    cube = g["cube"]
    p, d, n = O.sd_run_3d(cube, O.init_pred(cube), gamma=0.9, tau=0.8)
    same(p, g["sd_auto"]); same(d, g["sd_d_auto"]); assert n == len(g["sd_d_auto"])
    p, d, _ = O.sd_run_3d(cube, O.init_pred(cube), gamma=0.6, tau=0.95, iterations=6)
    same(p, g["sd_fix"]); same(d, g["sd_d_fix"])
    feat = g["feat"]
    p, d, _ = O.wp_run_3d(cube, O.init_pred(cube), feat.copy(), iterations=5, method="exact")
    same(p, g["wp_exact"]); same(d, g["wp_d_exact"])
    p, d, _ = O.wp_run_3d(cube, O.init_pred(cube), feat.copy(), iterations=4, method="cosine_similarity")
    same(p, g["wp_cos"]); same(d, g["wp_d_cos"])
    p, d, _ = O.wp_run_3d(cube, O.init_pred(cube), feat.copy(), method="exact")
    same(p, g["wp_auto"]); same(d, g["wp_d_auto"])
    flat = g["flat"].reshape(g["flat"].shape + (1,))
    p, d, _ = O.sd_run_3d(flat, O.init_pred(flat), iterations=3)
    same(p, g["sd_flat"]); same(d, g["sd_d_flat"])


def test_vpint2(golden):
    g = golden("vpint2")
    stored = g["stored_target"]
    assert stored.dtype == np.float32
    feats = g["features"].astype(np.float32)
    p, sweeps = O.vpint2_run(stored, feats)
    same(p, g["pred_par"]); assert list(sweeps) == list(g["sweeps"])
    same(p.astype(np.float32), g["pred_serial"])
    p, _ = O.vpint2_run(stored, feats, iterations=5)
    same(p.astype(np.float32), g["pred_fix"])
    p, _ = O.vpint2_run(stored, feats, iterations=4, clip_val=1500.0)
    same(p, g["pred_clip"])


def test_neighbour_count():
    nc = O.neighbour_count((3, 4))
    assert nc[0, 0] == 2 and nc[0, 1] == 3 and nc[1, 1] == 4
    assert O.neighbour_count((5, 1))[0, 0] == 1 and O.neighbour_count((5, 1))[2, 0] == 2
    assert O.neighbour_count((2, 3, 4))[0, 0, 0] == 3 and O.neighbour_count((3, 3, 3))[1, 1, 1] == 6
