#!/bin/bash
for v in "$@"; do
  EF_B200_HOST_CHUNK=$v timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/e2e_$v.json 2> gpurun_out/e2e_$v.err
  python - $v <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/e2e_{sys.argv[1]}.json'))
print("HOST_CHUNK", sys.argv[1], "device ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "e2e Mpix/s", round(d["e2e"]["value"]))
PY
done
