#!/bin/bash
# A/B session: full GPU parity tests, then short device-only benches under the env switches given in AB_ENVS (";"-separated, "-" = none)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { echo "### pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short ${NOX:--x} -p no:cacheprovider --durations=4 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log; }
IFS=';' read -ra ENVS <<< "${AB_ENVS:--}"
i=0
for e in "${ENVS[@]}"; do
  [ "$e" = "-" ] && e=""
  echo "### bench [$e]"
  env $e timeout 600 python bench.py --steps 5 --warmup 3 --batch ${AB_BATCH:-8} ${AB_E2E:---no-e2e} --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_ab$i.log 2> gpurun_out/bench_ab$i.err; echo "exit $?"
  python - "$i" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_ab{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
  i=$((i+1))
done
