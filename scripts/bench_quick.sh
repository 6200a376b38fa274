#!/bin/bash
# quick C3 comparison: prints device / kernel / e2e ms for a set of env / flag variants
one() {
  label="$1"; shift
  out=gpurun_out/q_$label.json
  "$@" > $out 2> gpurun_out/q_$label.err
  python - "$label" $out <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("%-28s dev %.2f ms  kernel %.2f ms  e2e %.2f ms (min %.2f) ok=%s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_min"], d["e2e"]["result_check"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
B="python bench.py --no-scene --no-variants --no-cpu-baseline"
one prio16 $B
one prio32 $B --chunk-patches 32
one prio32s6 $B --chunk-patches 32 --streams 6
one prio64s3 $B --chunk-patches 64 --streams 3
