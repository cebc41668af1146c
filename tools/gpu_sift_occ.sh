#!/bin/bash
# occupancy limiter + stall mix of the HashSIFT feature kernel (one launch under ncu), then the parity + short bench of gpu_sift_ab.sh
ncu --metrics launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:hashsift_pipe -c 1 python bench.py --batch 8 --no-extras --no-cpu-baseline --no-e2e --steps 1 --warmup 1 2>&1 | grep -E "launch__|sm__|smsp__|gpu__"
bash tools/gpu_sift_ab.sh
