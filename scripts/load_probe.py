"""Times the passes around run(): load with / without the device-side nanmean, weights, result (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_patches_torch, H, W, B
from vpint_b200.engine import Solver
from vpint_b200.vpint2 import _band_fill_values

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
target, feats = make_patches_torch(n, 1003, dev)[:2]
fill = _band_fill_values(target.cpu().numpy()).reshape(-1)
s = Solver(2, n * B, H, W, nprob_inner=B, planes=True, k_block=0, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tp, fp = target.permute(0, 3, 1, 2), feats.permute(0, 3, 1, 2).unsqueeze(-1)
print("patches %d: load(fill given) %.3f ms, load(device nanmean) %.3f ms, weights %.3f ms" % (
    n, timed(lambda: s.load(tp, fill=fill)), timed(lambda: s.load(tp, fill=None)), timed(lambda: s.weights(fp, method="exact"))))
s.load(tp, fill=None); s.weights(fp, method="exact"); s.run(5000, auto_terminate=True)
out = torch.empty((n, B, H, W), dtype=torch.float64, device=dev)
print("result %.3f ms" % timed(lambda: s.result(out=out)))
gb = n * B * H * W * 8 / 1e9
print("state plane = %.2f GB" % gb)
# one full step with events between the phases
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for rep in range(3):
    torch.cuda.synchronize()
    ev[0].record(); s.load(tp, fill=None)
    ev[1].record(); s.weights(fp, method="exact")
    ev[2].record(); s.run(5000, auto_terminate=True)
    ev[3].record(); s.result(out=out)
    ev[4].record(); torch.cuda.synchronize()
    print("step: load %.3f weights %.3f run %.3f result %.3f total %.3f ms" % tuple(
        [ev[i].elapsed_time(ev[i + 1]) for i in range(4)] + [ev[0].elapsed_time(ev[4])]))
