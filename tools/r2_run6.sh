#!/bin/bash
mkdir -p gpurun_out
rc=0
for t in tests/test_gpu_kernels.py tests/test_gpu_model.py; do
  name=$(basename "$t" .py)
  timeout 900 python -m pytest "$t" -q -m gpu -s --timeout 600 -p no:cacheprovider > "gpurun_out/r2f_${name}.log" 2>&1
  r=$?
  echo "== $t exit $r"; grep -E "passed|failed|error" "gpurun_out/r2f_${name}.log" | tail -3
  [ $r -ne 0 ] && { rc=$r; grep -E "^E |Error|error:|b2:|FAILED" "gpurun_out/r2f_${name}.log" | head -40; }
done
timeout 300 python bench.py --workload decode --steps 3 --warmup 1 2>/dev/null | cut -c1-200
timeout 300 python bench.py --steps 6 --warmup 3 --no-decode --no-precise --no-cpu-baseline 2>/dev/null | cut -c1-200
exit $rc
