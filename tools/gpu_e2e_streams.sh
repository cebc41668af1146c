#!/bin/bash
# host-buffer pipeline on one or two compute streams: host-API parity tests, then the e2e line of bench.py for both
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_api" 2>&1 | tail -1
for v in 1 2 1 2; do
  EF_B200_HOST_STREAMS=$v timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/e2e_streams.json 2> gpurun_out/e2e_streams.err
  python - $v <<'PY'
import json, sys
d=json.load(open('gpurun_out/e2e_streams.json'))
print("STREAMS", sys.argv[1], "device ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "e2e Mpix/s", round(d["e2e"]["value"]))
PY
done
