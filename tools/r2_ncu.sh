#!/bin/bash
# round-2 evidence: ncu --set full captures (one launch of every kernel on the prefill and decode paths).  gpurun returns at
# most 64 MiB per call, a --set full report is ~5-7 MB per kernel: run as  tools/r2_ncu.sh prefill | decode  (two calls).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
PRE="python bench.py --steps 1 --warmup 1 --no-decode --no-precise --no-cpu-baseline"
DEC="python bench.py --workload decode --steps 1 --warmup 1 --new-tokens 6"
KR='regex:gemm_sk_kernel|attn_decode_split|embed_norm_row|argmax_kernel|gemm_bf16_kernel'
if [ "$1" = "prefill" ]; then
  # prefill step = 116 launches: embed | 16 x (norm, qkv, attn, o_proj, norm, glu, down) | final norm, lm_head, add_scalar
  # (-k: count library kernels only -- bench.py generates its synthetic checkpoint with torch kernels first)
  KP='regex:gemm2_bf16_kernel|gemm_bf16_kernel|attn_prefill|rmsnorm_kernel|embed_kernel|add_scalar_kernel|gather_last_kernel'
  timeout 600 $NCU --import-source on -k "$KP" -s 116 -c 8 -o gpurun_out/r02_prefill_layer0 -f $PRE > /dev/null 2>&1; echo "ncu prefill layer rc $?"
  timeout 600 $NCU -k "$KP" -s 229 -c 2 -o gpurun_out/r02_prefill_lmhead -f $PRE > /dev/null 2>&1; echo "ncu prefill lm_head rc $?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KP" -s 348 -c 232 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 2 --no-decode --no-precise --no-cpu-baseline > /dev/null 2>&1; echo "launch list rc $?"
else
  # decode (fused path, no graph): embed_norm + one layer (5 kernels) + the next layer's QKV; then lm_head + argmax of the same step
  B2_NO_GRAPH=1 timeout 600 $NCU -k "$KR" -s 85 -c 7 -o gpurun_out/r02_decode_layer0 -f $DEC > /dev/null 2>&1; echo "ncu decode layer rc $?"
  B2_NO_GRAPH=1 timeout 600 $NCU -k "$KR" -s 166 -c 2 -o gpurun_out/r02_decode_lmhead -f $DEC > /dev/null 2>&1; echo "ncu decode lm_head rc $?"
  # Gemma-2: post-norm (add_rmsnorm) and the head_dim-256 attention kernel
  timeout 600 $NCU -k 'regex:add_rmsnorm_kernel|attn_prefill_kernel' -s 8 -c 2 -o gpurun_out/r02_gemma_kernels -f \
    python bench.py --model gemma2-2b --seq 4096 --batch 2 --steps 1 --warmup 1 --no-decode --no-precise --no-cpu-baseline > /dev/null 2>&1; echo "ncu gemma rc $?"
  B2_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_decode_launches.csv $DEC > /dev/null 2>&1; echo "decode launch list rc $?"
fi
# never lose the whole directory to the 64 MiB cap: drop the largest report until it fits
while [ "$(du -sm gpurun_out | cut -f1)" -ge 60 ]; do big=$(ls -S gpurun_out/*.ncu-rep | head -1); echo "dropping $big ($(du -m $big | cut -f1) MB) to stay under the return cap"; rm -f "$big"; done
ls -la gpurun_out/ | head -30; du -sm gpurun_out
