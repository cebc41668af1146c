#!/bin/bash
# ncu evidence of one bench step for the current kernels: launch list (gpu__time_duration) + one --set full capture with sources
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-extras"
K='regex:ef_(resize|score|nms|compact|select|angle|blur|hashsift|bad)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 15 -c 15 --csv --log-file gpurun_out/r2_launches_hs.csv $B > gpurun_out/r2_launches_hs.out 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 15 -c 15 -o gpurun_out/r2_prof_all -f $B > gpurun_out/r2_prof_all.out 2>&1; echo "full capture exit $?"
ls -la gpurun_out | grep r2_prof_all
