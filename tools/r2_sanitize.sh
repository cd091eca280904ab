#!/bin/bash
# compute-sanitizer memcheck over every kernel-level and model-level GPU test (tiny and mid shapes)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_dp.py \
  -q -m gpu -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r02_sanitizer_memcheck.log | tail -8
