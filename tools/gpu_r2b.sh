#!/bin/bash
# quick check of a detector-kernel change: stage / detector parity tests, then a short bench without the extras
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kernels.py tests/test_gpu_band_sharding.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
tail -3 gpurun_out/r2b_bench.err
