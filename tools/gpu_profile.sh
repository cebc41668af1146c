#!/bin/bash
# Evidence run for profiles/: ncu launch list (gpu__time_duration) + full capture of every kernel of one step (1 GPU), matcher capture.
# Outputs -> gpurun_out/; summarise here with tools/ncu_summary.py (see profiles/README.md).
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline"
echo "### launch list (HashSIFT-512)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ef_" -s 15 -c 15 --csv --log-file gpurun_out/launches_hs.csv $B > gpurun_out/launches_hs.out 2>&1; echo "exit $?"
echo "### launch list (BAD-512)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ef_" -s 14 -c 14 --csv --log-file gpurun_out/launches_bad.csv $B --desc BAD_512 > gpurun_out/launches_bad.out 2>&1; echo "exit $?"
echo "### full capture (HashSIFT-512 step)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ef_" -s 15 -c 15 -o gpurun_out/prof_all -f $B > gpurun_out/prof_all.out 2>&1; echo "exit $?"
echo "### full capture (BAD-512 describe)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ef_bad_pipe" -s 1 -c 1 -o gpurun_out/prof_bad -f $B --desc BAD_512 > gpurun_out/prof_bad.out 2>&1; echo "exit $?"
echo "### full capture (matcher 40k x 40k x 512)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ef_match" -s 8 -c 4 -o gpurun_out/prof_match -f python tools/match_bench.py > gpurun_out/prof_match.out 2>&1; echo "exit $?"
ls -la gpurun_out | tail -12
