#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|probe OK|Invalid|at ef_" gpurun_out/r2_memcheck.log | head
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|probe OK|hazard|at ef_" gpurun_out/r2_racecheck.log | head
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2_synccheck.log 2>&1; echo "synccheck exit $?"; grep -E "ERROR SUMMARY|probe OK|at ef_" gpurun_out/r2_synccheck.log | head
