"""Stage timing of the fused load on a shard-sized problem (one image of ROWS x 10980 x 13 uint16), one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_scene
from vpint_b200.engine import Solver

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1373
W, B = 10980, 13
dev = torch.device("cuda", 0)
t, f, c = bench_scene.scene_rows(0, rows, rows, W, B, 1004, dev, 60)
s = Solver(2, B, rows, W, nprob_inner=B, planes=True, k_block=1, device=dev)
sc = s.fill_scratch("float32")
t4, f4, c3 = t.unsqueeze(0), f.unsqueeze(0), c.unsqueeze(0)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e = [ev() for _ in range(8)]
    e[0].record(); s.fill_leaves(t4, c3, sc)
    e[1].record(); s.fill_combine(sc)
    e[2].record(); s.ingest_bands(t4, f4, c3, fill_ready=True)
    e[3].record(); s.begin_run(5000)
    e[4].record(); s.step_begin()
    e[5].record(); s.step_end()
    for _ in range(3):
        s.step_begin(); s.step_end()
    e[6].record(); out, nit = s.result()
    e[7].record(); torch.cuda.synchronize()
    names = ["fill_leaves", "fill_combine", "ingest+reduce", "begin_run(tiles)", "first step (to_ratio + sweep)", "decide + 3 sweeps", "result (from_ratio + gather)"]
    print("rep", rep, ", ".join("%s %.3f" % (n, e[i].elapsed_time(e[i + 1])) for i, n in enumerate(names)))
