#!/bin/bash
# Quick GPU-box session: parity tests + short bench.  Logs -> gpurun_out/
#   QUICK=1: only the stage-by-stage test and one detectAndCompute case per descriptor
mkdir -p gpurun_out
SEL=""
if [ -n "$QUICK" ]; then SEL="-k test_stages_match_oracle or (test_detect_and_compute_matches_oracle and 800) or test_detect_only_and_params or test_host_api_and_batch"; fi
echo "### pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short ${NOX:--x} -p no:cacheprovider --durations=4 ${SEL:+"$SEL"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
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
