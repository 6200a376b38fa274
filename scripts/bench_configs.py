#!/usr/bin/env python
"""Device timings of the other BASELINE configs (1, 2, 5) -- parity cases, not the headline bench line.
Prints one JSON object per config: run() device time with inputs resident, sweeps, Mpx*it/s."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def smooth(rng, shape, lo, hi, modes=5):
    axes = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij", sparse=True)
    acc = 0.0
    for _ in range(modes):
        term = 1.0
        for ax in axes:
            term = term * np.sin(2 * np.pi * (rng.uniform(0.3, 2.0) * ax + rng.uniform()))
        acc = acc + rng.uniform(0.5, 1.0) * term
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
    return lo + (hi - lo) * acc


def timed_solver(make, reps=3):
    import torch
    best = None
    for _ in range(reps):
        s, run = make()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out, nit = s.result()
        sweeps = nit.cpu().numpy()
        s.close()
        best = ms if best is None else min(best, ms)
    return best, sweeps


def c5(torch, Solver, dev, res):
    # C5: SD_STMRP 2048x2048x64, 30 % missing (disc clouds per time slice), generated on the device
    Hc, Wc, Tc = 2048, 2048, 64
    g = torch.Generator(device=dev); g.manual_seed(1005)
    yy = (torch.arange(Hc, device=dev, dtype=torch.float64) / Hc).view(-1, 1, 1)
    xx = (torch.arange(Wc, device=dev, dtype=torch.float64) / Wc).view(1, -1, 1)
    tt = (torch.arange(Tc, device=dev, dtype=torch.float64) / Tc).view(1, 1, -1)
    cube = 20.0 + 5.0 * torch.sin(2 * np.pi * (0.7 * yy + 0.1)) * torch.cos(2 * np.pi * (1.3 * xx + 0.4)) * (1.0 + 0.2 * torch.sin(2 * np.pi * tt))
    cube = cube + 0.3 * torch.randn((Hc, Wc, Tc), generator=g, device=dev, dtype=torch.float64)
    miss = torch.zeros((Hc, Wc, Tc), dtype=torch.bool, device=dev)
    discs = torch.rand((Tc, 8, 3), generator=g, device=dev, dtype=torch.float64)
    for t in range(Tc):
        m2 = torch.zeros((Hc, Wc), dtype=torch.bool, device=dev)
        for k in range(8):
            cy, cx, r = discs[t, k]
            rad = 0.05 + 0.12 * r
            m2 |= (yy[:, :, 0] - cy) ** 2 + (xx[:, :, 0] - cx) ** 2 < rad * rad
        miss[:, :, t] = m2
    mean = float(cube[~miss].mean().item())
    cube[miss] = float("nan")
    c_dev = cube.unsqueeze(0)
    frac = float(miss.double().mean().item())
    for it in (-1, 50):
        def make(it=it):
            s = Solver(3, 1, Hc, Wc, Tc, planes=False, device=dev)
            s.load(c_dev, fill=[mean])
            s.discounts(0.9, 0.9)
            return s, (lambda s: s.run(10000 if it < 0 else it, auto_terminate=it < 0))
        ms, sw = timed_solver(make, reps=2)
        res.append({"config": "C5 SD_STMRP 2048x2048x64 %.0f%% missing, %s" % (100 * frac, "auto-terminate" if it < 0 else "run(50)"),
                    "ms": ms, "sweeps": int(sw[0]), "mcell_it_per_s": Hc * Wc * Tc * int(sw[0]) / ms / 1e3,
                    "algorithmic_gbps_unknown_cells": 16.0 * Hc * Wc * Tc * frac * int(sw[0]) / ms / 1e6})


def main():
    import torch
    from vpint_b200.engine import Solver
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1001)
    res = []
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""     # "c5": the 3-D cube alone (ncu captures)
    if only == "c5":
        c5(torch, Solver, dev, res)
        for r in res:
            print(json.dumps(r))
        return
    # C1: SD_SMRP 100x100, 20 % hidden, gamma 0.9
    grid = smooth(rng, (100, 100), 15, 25) + rng.normal(0, 1, (100, 100))
    grid[rng.uniform(size=grid.shape) < 0.2] = np.nan
    g_dev = torch.from_numpy(grid).to(dev).unsqueeze(0)

    def mk1(it):
        def make():
            s = Solver(2, 1, 100, 100, planes=False, device=dev)
            s.load(g_dev, fill=[np.nanmean(grid)])
            s.discounts(0.9)
            return s, (lambda s: s.run(10000 if it < 0 else it, auto_terminate=it < 0))
        return make
    for it in (-1, 100):
        ms, sw = timed_solver(mk1(it))
        res.append({"config": "C1 SD_SMRP 100x100 20%% hidden, %s" % ("auto-terminate" if it < 0 else "run(100)"), "ms": ms,
                    "sweeps": int(sw[0]), "mpx_it_per_s": 1e4 * int(sw[0]) / ms / 1e3})
    # C2: WP_SMRP 1024x1024, 3 feature layers, 12 disc clouds
    shape = (1024, 1024)
    base = smooth(rng, shape, 300, 2700, 6)
    tgt = base + rng.normal(0, 20, shape)
    feat = np.stack([rng.uniform(0.7, 1.3) * base + rng.normal(0, 10, shape) for _ in range(3)], axis=-1)
    feat = np.maximum(feat, 50.0)
    yy, xx = np.mgrid[0:1024, 0:1024]
    cloud = np.zeros(shape, dtype=bool)
    for _ in range(12):
        cy, cx, r = rng.uniform(0, 1024), rng.uniform(0, 1024), rng.uniform(30, 120)
        cloud |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
    tgt[cloud] = np.nan
    t_dev = torch.from_numpy(tgt).to(dev).unsqueeze(0)
    f_dev = torch.from_numpy(feat).to(dev).unsqueeze(0)
    for kb in (1, 4):
        for it in (-1, 200):
            def make(kb=kb, it=it):
                s = Solver(2, 1, 1024, 1024, planes=True, k_block=kb, device=dev)
                s.load(t_dev, fill=[np.nanmean(tgt)])
                s.weights(f_dev, method="exact")
                return s, (lambda s: s.run(5000 if it < 0 else it, auto_terminate=it < 0))
            ms, sw = timed_solver(make)
            res.append({"config": "C2 WP_SMRP 1024x1024 F=3 %.0f%% cloud, k_block %d, %s" % (100 * cloud.mean(), kb, "auto-terminate" if it < 0 else "run(200)"),
                        "ms": ms, "sweeps": int(sw[0]), "mpx_it_per_s": 1024 * 1024 * int(sw[0]) / ms / 1e3})
    # C2 through the drop-in class (host arrays in, host array out, default path: 4 sweeps per launch for F > 1)
    from vpint_b200 import WP_SMRP
    WP_SMRP(tgt.copy(), feat.copy()).run()
    walls = []
    for _ in range(3):
        m = WP_SMRP(tgt.copy(), feat.copy())
        t0 = time.perf_counter()
        m.run()
        walls.append(1e3 * (time.perf_counter() - t0))
    res.append({"config": "C2 WP_SMRP(...).run() through the class API, wall clock incl. H2D / D2H", "ms": min(walls)})
    c5(torch, Solver, dev, res)
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
