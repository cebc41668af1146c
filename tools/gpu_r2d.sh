#!/bin/bash
# natural-image table (reference CLI over the adapter + reference kernels) and the ncu capture of one step
mkdir -p gpurun_out profiles/r02
python - <<'PY'
import sys; sys.path.insert(0,'tests')
import test_gpu_adapter as t; t.sidecars()
PY
timeout 1500 python tools/natural_bench.py --iterations 30 > gpurun_out/r2_natural_images.json 2> gpurun_out/r2_natural_images.err; echo "natural exit $?"; tail -3 gpurun_out/r2_natural_images.err
B="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-extras"
K='regex:ef_(resize|score|nms|compact|select|angle|blur|hashsift|bad)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 15 -c 15 --csv --log-file gpurun_out/r2_launches_hs.csv $B > gpurun_out/r2_launches_hs.out 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 15 -c 15 -o gpurun_out/r2_prof_all -f $B > gpurun_out/r2_prof_all.out 2>&1; echo "full capture exit $?"
ls -la gpurun_out | grep r2_
