#!/bin/bash
# A/B of kernel variants selected by an environment variable: tools/gpu_ab_env.sh VAR v1 v2 ...   (short bench per value)
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  python - "$VAR" "$v" <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/ab_{sys.argv[1]}_{sys.argv[2]}.json'))
print(sys.argv[1], sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done
