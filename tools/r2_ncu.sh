#!/bin/bash
# round-2 evidence: ncu --set full captures (one launch of every kernel on the prefill and decode paths); everything written
# under gpurun_out/ must stay below 64 MiB in total, so captures are short and only the prefill layer keeps the source view
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
PRE="python bench.py --steps 1 --warmup 1 --no-decode --no-precise --no-cpu-baseline"
# prefill step = 116 launches: embed | 16 x (norm, qkv, attn, o_proj, norm, glu, down) | final norm, lm_head, add_scalar
timeout 600 $NCU --import-source on -s 116 -c 8 -o gpurun_out/r02_prefill_layer0 -f $PRE > gpurun_out/r02_ncu_prefill_a.log 2>&1; echo "ncu prefill layer rc $?"
timeout 600 $NCU -s 229 -c 2 -o gpurun_out/r02_prefill_lmhead -f $PRE > gpurun_out/r02_ncu_prefill_b.log 2>&1; echo "ncu prefill lm_head rc $?"
# decode (fused path, no graph): embed_norm + one layer (5 kernels) + the next layer's QKV; then lm_head + argmax of the same step
DEC="python bench.py --workload decode --steps 1 --warmup 1 --new-tokens 6"
KR='regex:gemm_sk_kernel|attn_decode_split|embed_norm_row|argmax_kernel|gemm_bf16_kernel'
B2_NO_GRAPH=1 timeout 600 $NCU -k "$KR" -s 85 -c 7 -o gpurun_out/r02_decode_layer0 -f $DEC > gpurun_out/r02_ncu_decode_a.log 2>&1; echo "ncu decode layer rc $?"
B2_NO_GRAPH=1 timeout 600 $NCU -k "$KR" -s 166 -c 2 -o gpurun_out/r02_decode_lmhead -f $DEC > gpurun_out/r02_ncu_decode_b.log 2>&1; echo "ncu decode lm_head rc $?"
# Gemma-2: post-norm (add_rmsnorm) and the head_dim-256 attention kernel
timeout 600 $NCU -k 'regex:add_rmsnorm_kernel|attn_prefill_kernel' -s 8 -c 2 -o gpurun_out/r02_gemma_kernels -f \
  python bench.py --model gemma2-2b --seq 4096 --batch 2 --steps 1 --warmup 1 --no-decode --no-precise --no-cpu-baseline > gpurun_out/r02_ncu_gemma.log 2>&1; echo "ncu gemma rc $?"
# launch list (gpu__time_duration only) of the default prefill run: shares of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 348 -c 232 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 2 --no-decode --no-precise --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc $?"
B2_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_decode_launches.csv $DEC > /dev/null 2>&1; echo "decode launch list rc $?"
rm -f gpurun_out/*.log
ls -la gpurun_out/ | head -30; du -sh gpurun_out
