#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kernels.py tests/test_gpu_band_sharding.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2c_pytest.log
bash tools/gpu_ab_env.sh ${1:-EF_BLUR_TMA} 0 1 0 1
