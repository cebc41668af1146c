#!/bin/bash
# Quick GPU-box session: parity tests + short bench.  Logs -> gpurun_out/
mkdir -p gpurun_out
echo "### pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "### bench"; timeout 900 python bench.py --steps 5 --warmup 3 --batch 8 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "fps", round(d["frames_per_s"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1))
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench.err
