#!/bin/bash
# Round-end evidence on one B200: parity tests, sanitizer, bench lines, ncu launch lists and the --set full capture of one step.
# Outputs -> gpurun_out/final_*; summarise with tools/ncu_summary.py and copy into profiles/ (see profiles/README.md).
mkdir -p gpurun_out
echo "### pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/final_pytest.log
echo "### sanitizer"
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/final_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|probe OK" gpurun_out/final_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/final_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|probe OK" gpurun_out/final_racecheck.log
echo "### bench default"; timeout 600 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "exit $?"; cut -c1-400 gpurun_out/final_bench_default.json
echo "### bench BAD_512"; timeout 300 python bench.py --desc BAD_512 --no-cpu-baseline > gpurun_out/final_bench_bad512.json 2> gpurun_out/final_bench_bad512.err; echo "exit $?"; cut -c1-200 gpurun_out/final_bench_bad512.json
B="python bench.py --steps 1 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline"
echo "### launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ef_" -s 15 -c 15 --csv --log-file gpurun_out/final_launches_hs.csv $B > gpurun_out/final_launches_hs.out 2>&1; echo "exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ef_" -s 14 -c 14 --csv --log-file gpurun_out/final_launches_bad.csv $B --desc BAD_512 > gpurun_out/final_launches_bad.out 2>&1; echo "exit $?"
echo "### full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ef_" -s 15 -c 15 -o gpurun_out/final_prof_all -f $B > gpurun_out/final_prof_all.out 2>&1; echo "exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ef_bad_pipe" -s 1 -c 1 -o gpurun_out/final_prof_bad -f $B --desc BAD_512 > gpurun_out/final_prof_bad.out 2>&1; echo "exit $?"
ls -la gpurun_out | grep final_
