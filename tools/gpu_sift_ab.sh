#!/bin/bash
# parity of the descriptor paths + a short device-only bench at batch 16 (stage times comparable with profiles/r02)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_natural_images.py tests/test_gpu_adapter.py -m gpu -x -q 2>&1 | tail -3
python bench.py --batch 16 --no-extras --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})"
