#!/bin/bash
# Focused capture: ncu --set full of the kernels matching $KREGEX (skip $SKIP launches, capture $COUNT) in one bench step -> gpurun_out/prof_one.ncu-rep
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline ${BENCH_ARGS}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-ef_}" -s ${SKIP:-0} -c ${COUNT:-8} -o gpurun_out/prof_one -f $B > gpurun_out/prof_one.out 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/prof_one.out; ls -la gpurun_out/prof_one.ncu-rep
