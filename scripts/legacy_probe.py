"""Per-call time of the round-1 host pipeline (solve_bands on pinned float32 arrays), with and without a RawPipeline alive."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vpint_b200.vpint2 import RawPipeline, solve_bands
from vpint_b200.wp import wp_run_options
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
target, feats, t_raw, f_raw, mask = bench.make_patches_torch(n, 1003, dev)
opts = wp_run_options()
t32 = torch.empty(target.shape, dtype=torch.float32).pin_memory(); t32.copy_(target)
f32 = torch.empty(feats.shape, dtype=torch.float32).pin_memory(); f32.copy_(feats)
out = torch.empty((n, 13, 256, 256), dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
def legacy():
    t0 = time.perf_counter()
    solve_bands(t32, f32, opts, out_layout="bhw", return_sweeps=True, device=dev, out=out)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0)
print("legacy alone:", ["%.1f" % legacy() for _ in range(8)])
pipe = RawPipeline(dev)
od = torch.empty((n, 13, 256, 256), dtype=torch.float64, device=dev)
for _ in range(3):
    pipe.solve(target, feats, None, opts, out=od)
torch.cuda.synchronize()
print("legacy after raw pipeline:", ["%.1f" % legacy() for _ in range(8)])
print(torch.cuda.memory_reserved() / 1e9, "GB reserved")
