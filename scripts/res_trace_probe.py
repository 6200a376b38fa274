"""Stage timing inside the resident kernel (needs a library built with -DVPINT_RES_TRACE; GPU box)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vpint_b200 import _native
from vpint_b200.engine import Solver

H = W = 256
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = 33
dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
t = rng.uniform(1000, 2000, (P, H, W)); f = rng.uniform(1000, 2000, (P, H, W, 1))
if k: t[:, :, 4:4 + 4 * k] = np.nan
else: t[:, 100, 100] = np.nan
s = Solver(2, P, H, W, planes=True, k_block=0, device=dev)
for rep in range(2):
    s.load(torch.from_numpy(t).to(dev), fill=np.full(P, 1500.0)); s.weights(torch.from_numpy(f).to(dev), method="exact")
    s.run(100, auto_terminate=False); torch.cuda.synchronize()
lib = _native.load_library()
buf = np.zeros((8, 2, 8, 10), dtype=np.int64)
rc = lib.vpint_debug_trace(buf.ctypes.data_as(C.c_void_p), buf.nbytes)
assert rc == 0, rc
names = ["top", "halo_wait", "compute", "warp_sum", "decide", "bar+verdict", "push", "stores", "syncwarp+arrive"]
print("k=%d  stage deltas in cycles, mean over sweeps 10..17; period = top(t+1) - top(t)" % k)
for cta in range(4):
    for who in range(2):
        tr = buf[cta, who].astype(np.float64)       # [sweep][stage]
        d = np.diff(tr[:, :9], axis=1).mean(axis=0)
        period = np.diff(tr[:, 0]).mean()
        print("cta %d %s  period %6.0f | " % (cta, "w0 " if who == 0 else "syn", period) +
              "  ".join("%s %5.0f" % (n, v) for n, v in zip(names[1:], d)))
