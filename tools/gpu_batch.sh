#!/bin/bash
for b in 32 64; do
  timeout 600 python bench.py --batch $b --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/batch_$b.json 2> gpurun_out/batch_$b.err
  python - $b <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/batch_{sys.argv[1]}.json'))
print("BATCH", sys.argv[1], "value", round(d["value"]), "ms/frame", round(d["ms_per_step"]/int(sys.argv[1]),4), "e2e", round(d["e2e"]["value"]), "e2e ms/frame", round(d["e2e"]["ms_per_step"]/int(sys.argv[1]),4))
PY
done
