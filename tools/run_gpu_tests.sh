#!/bin/bash
# Runs each GPU test module in its own process (a CUDA fault in one cannot poison the others),
# each under a timeout, logs into gpurun_out/.
mkdir -p gpurun_out
rc=0
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 900 python -m pytest "$t" -q -m gpu -s --timeout 300 -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  r=$?
  echo "== $t exit $r"; tail -n 25 "gpurun_out/${name}.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
