#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_natural_images.py tests/test_gpu_adapter.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2e_pytest.log
for d in HASH_SIFT_256 HASH_SIFT_512 BAD_256 BAD_512; do
timeout 600 python bench.py --desc $d --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2e_bench_$d.json 2> gpurun_out/r2e_bench_$d.err
python - $d <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/r2e_bench_{sys.argv[1]}.json'))
print(sys.argv[1], "value", round(d["value"]), "fps", round(d["frames_per_s"]), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms_per_step"].items() if v>0})
PY
done
