#!/bin/bash
# Copy-engine-fed ratio stencil: parity first (bounded), then timing against the register-fed kernel.
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_raw.py -x -q -k "library or lagged or lockstep or copy_engine or large_scene" 2>&1 | tail -2
for v in "VPINT_RATIO_KERNEL=4" "VPINT_NO_F32_PLANE=1" "VPINT_X=1"; do
  echo "$v"
  env $v timeout -k 10 200 python scripts/load_probe_raw.py 1373 2>&1 | tail -1
done
if [ "${1:-}" = "full" ]; then
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -k 10 600 python bench.py --steps 2 --warmup 1 --e2e-steps 2 --roofline-launches 4 --no-cpu-baseline --no-variants --patches 64 > gpurun_out/bulk_bench.json 2> gpurun_out/bulk_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bulk_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); s=d['scene']; print('parity', s['parity_ok'], s['parity']['max_rel'], 'c4 ms', s['c4']['ms'], s['c4']['phases_ms_rank0'], s['c4']['sweeps_per_band'])
PY
tail -2 gpurun_out/bulk_bench.err | cut -c1-300
fi
