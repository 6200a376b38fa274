#!/bin/bash
# Runs on the GPU box via gpurun: parity tests, smoke, small benches per k_block.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -40
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
for kb in ${KBS:-1 4 6 8}; do
  echo "== bench small k_block=$kb"
  timeout 600 python bench.py --patches 64 --steps 2 --warmup 1 --k-block $kb --no-cpu-baseline > gpurun_out/bench_small_k$kb.json 2> gpurun_out/bench_small_k$kb.err; echo "rc=$?"
  tail -2 gpurun_out/bench_small_k$kb.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_small_k$kb.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["config"]["mean_sweeps_per_problem"])
except Exception as e:
    print("no json", e)
PY
done
