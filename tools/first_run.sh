#!/bin/bash
mkdir -p gpurun_out
for st in info gemm attn; do
  echo "=== stage $st"; timeout 300 python tools/gpu_diag.py $st 2>&1 | tail -40
done
