#!/bin/bash
# One GPU-box session: probe, sanitizer, parity tests, short bench.  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "### probe"; timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit $?"; tail -5 gpurun_out/probe.log
echo "### sanitizer"; timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python tools/gpu_probe.py --w 400 --h 300 --nfeat 800 > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitizer.log | head -10
echo "### pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
echo "### bench"; timeout 900 python bench.py --steps 5 --warmup 3 --batch 8 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
