#!/bin/bash
# round 2, first GPU session: full -m gpu suite (incl. the new photograph / 8K / adapter tests), smoke, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "### pytest"; timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/r2a_pytest.log
echo "### smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2a_smoke.log
echo "### bench"; timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; cat gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
