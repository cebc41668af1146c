#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  echo "== EF_BLUR_TMA=$v"
  EF_BLUR_TMA=$v timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_probe.py --w 400 --h 300 --nfeat 800 > gpurun_out/dbg_$v.log 2>&1
  grep -E "ERROR SUMMARY|Invalid|Illegal|illegal|at 0x|by thread|in ef_|kernel" gpurun_out/dbg_$v.log | head -12
done
