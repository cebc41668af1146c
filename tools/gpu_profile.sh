#!/bin/bash
# ncu launch list + full capture of the hot kernels (1 GPU).  Outputs -> gpurun_out/
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline"
echo "### launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 15 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.out 2>&1
echo "exit $?"; tail -3 gpurun_out/launches.out
echo "### full capture"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"ef_hashsift_pipe|ef_nms|ef_score|ef_hashsift_project|ef_resize|ef_blur|ef_bad_pipe" -s 12 -c 12 -o gpurun_out/prof_full -f $B > gpurun_out/prof_full.out 2>&1
echo "exit $?"; tail -3 gpurun_out/prof_full.out; ls -la gpurun_out
