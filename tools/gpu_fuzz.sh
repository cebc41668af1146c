#!/bin/bash
# wider sweep of the randomised differential test: tools/gpu_fuzz.sh [cases]
EF_FUZZ_CASES=${1:-160} EF_FUZZ_COMPUTE_CASES=${2:-24} EF_FUZZ_BAND_CASES=${3:-12} EF_FUZZ_MATCH_CASES=${4:-16} timeout 1500 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_band_sharding.py -m gpu -q 2>&1 | tail -40
