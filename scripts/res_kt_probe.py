"""Per-sweep cost of the resident kernel against the segments per thread (run on the GPU box).

Grids of 256x256 whose first k segments of every row are unknown: every CTA of a 4-CTA cluster then holds
64*k segments, i.e. k/8 per thread.  Fixed sweep count, no stop rule -> time / sweeps is the cost of one sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vpint_b200.engine import Solver

H = W = 256
P = int(sys.argv[1]) if len(sys.argv) > 1 else 132
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
KS = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 4, 8, 16, 24, 32, 48, 63]
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 3
for k in KS:
    t = rng.uniform(1000, 2000, (P, H, W))
    f = np.round(rng.uniform(1000, 2000, (P, H, W, 1)))   # float32-exact, like VPint2 features
    if k:
        t[:, :, 4:4 + 4 * k] = np.nan
    else:
        t[:, 100, 100] = np.nan
    tt, ff = torch.from_numpy(t).to(dev), torch.from_numpy(f).to(dev)
    s = Solver(2, P, H, W, planes=True, k_block=0, device=dev)
    best = 1e9
    for rep in range(REPS):
        s.load(tt, fill=np.full(P, 1500.0))
        s.weights(ff, method="exact")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run(sweeps, auto_terminate=False); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rounds = -(-P // 33)
    print("k=%2d segs/thread=%.2f  run %.3f ms  -> %.3f us per sweep (%d rounds of 33 clusters)" % (
        k, 64 * k / 512.0, best, best * 1e3 / (rounds * sweeps), rounds), flush=True)
    del s
