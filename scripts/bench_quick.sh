#!/bin/bash
# quick C3 comparison: prints device / kernel / e2e ms for a set of env / flag variants
one() {
  label="$1"; shift
  out=gpurun_out/q_$label.json
  "$@" > $out 2> gpurun_out/q_$label.err
  python - "$label" $out <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print("%-28s dev %.2f ms  kernel %.2f ms  e2e %.2f ms (min %.2f) ok=%s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_min"], d["e2e"]["result_check"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
B="python bench.py --no-scene --no-variants --no-cpu-baseline --e2e-steps 5"
one classes_234 $B
VPINT_CLASS_MASK=4 one classes_24 $B
VPINT_CLASS_MASK=8 one classes_34 $B
