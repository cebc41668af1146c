#!/bin/bash
for v in "$@"; do
  EF_B200_HOST_PLAN=$v timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/e2e_plan.json 2> gpurun_out/e2e_plan.err
  python - $v <<'PY'
import json, sys
d=json.load(open('gpurun_out/e2e_plan.json'))
print("PLAN", sys.argv[1], "device ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "e2e Mpix/s", round(d["e2e"]["value"]))
PY
done
