#!/bin/bash
# Runs on the GPU box via gpurun: parity tests, smoke, bench, then an ncu launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
echo "== bench small"
timeout 600 python bench.py --patches 64 --steps 2 --warmup 1 --cpu-patches 2 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small rc=$?"
tail -2 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
