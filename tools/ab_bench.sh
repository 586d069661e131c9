#!/bin/bash
# A/B of step-kernel variants on the GPU box: tools/ab_bench.sh "ENV=VAL ..." envs [steps]
# prints value / ms_per_step / frac of a short bench run (no sweep, no CPU baseline).
set -u
vars="$1"; envs="$2"; steps="${3:-300}"
out=$(env $vars python bench.py --envs "$envs" --steps "$steps" --warmup 10 --no-sweep --no-cpu-baseline 2>&1 | tail -1)
python - "$vars" "$envs" <<PY
import json,sys
try:
    d=json.loads('''$out''')
    print("%-45s envs=%-6s ms/step=%.5f  env-steps/s=%.4g  frac=%.3f  e2e=%.4g" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e); print('''$out'''[-600:])
PY
