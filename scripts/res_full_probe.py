"""The resident kernel at the benchmark's size: ONE solver holding the whole 512-patch batch (6656 band problems), fused
load, run(), repeated; prints per-launch run() time.  Target of the round-2 `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_patches_torch, H, W, B
from vpint_b200.engine import Solver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
target, feats, _, _, _ = make_patches_torch(n, 1003, dev)
s = Solver(2, n * B, H, W, nprob_inner=B, planes=True, k_block=0, device=dev)
sc = s.fill_scratch("float32")
for rep in range(reps):
    s.ingest_bands(target, feats, None, value_dtype="float32", scratch=sc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.run(5000, auto_terminate=True); e1.record()
    torch.cuda.synchronize()
    out, nit = s.result()
    print("rep %d run %.3f ms, problem-sweeps %d, max %d" % (rep, e0.elapsed_time(e1), int(nit.sum().item()), int(nit.max().item())))
