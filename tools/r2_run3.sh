#!/bin/bash
# round-2 GPU run #3: persistent attention kernel (tests + A/B against the one-tile-per-CTA kernel), decode path v2, baseline shapes
mkdir -p gpurun_out
rc=0
for t in tests/test_gpu_kernels.py tests/test_gpu_model.py; do
  name=$(basename "$t" .py)
  timeout 900 python -m pytest "$t" -q -m gpu -s --timeout 600 -p no:cacheprovider -x > "gpurun_out/r2c_${name}.log" 2>&1
  r=$?
  echo "== $t exit $r"; grep -E "passed|failed|error" "gpurun_out/r2c_${name}.log" | tail -3
  [ $r -ne 0 ] && { rc=$r; grep -E "^E |Error|error:|b2:" "gpurun_out/r2c_${name}.log" | head -30; }
done
if [ $rc -ne 0 ]; then echo "kernel/model tests failed: A/B with the one-tile attention kernel for the rest"; fi
pre() { timeout 300 python bench.py --steps 6 --warmup 3 --no-decode --no-precise --no-cpu-baseline --event-every 1 "$@" 2>gpurun_out/r2c_pre.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['roofline']['attention']
    print(round(d['value']), 'tok/s', round(d['ms_per_step'],3), 'ms; attn ms/step', round(a['ms_per_step'],3), 'attn TF', round(a['achieved'],1), 'gemm TF', round(d['roofline']['achieved'],1))
except Exception as e: print('failed', e); print(open('gpurun_out/r2c_pre.err').read()[-1500:])"; }
for rep in 1 2; do
  echo "== prefill 1B persistent attention (rep $rep)"; pre
  echo "== prefill 1B one-tile-per-CTA attention (rep $rep)"; B2_ATTN_CO_ONESHOT=1 pre
done
echo "== prefill 3B S=8192 persistent"; pre --model 3b --seq 8192 --batch 2 --steps 4
echo "== prefill 3B S=8192 one-tile"; B2_ATTN_CO_ONESHOT=1 pre --model 3b --seq 8192 --batch 2 --steps 4
dec() { timeout 300 python bench.py --workload decode --steps 3 --warmup 1 "$@" 2>gpurun_out/r2c_dec.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'tok/s', round(d['ms_per_token_step'],4), 'ms/step frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], 'e2e', round(d['e2e']['value']))
except Exception as e: print('decode bench failed', e); print(open('gpurun_out/r2c_dec.err').read()[-1500:])"; }
for rep in 1 2; do
  echo "== decode fused v2 (rep $rep)"; dec
  echo "== decode 7-kernel path (rep $rep)"; B2_NO_FUSED_DECODE=1 dec
done
echo "== decode launch list (fused v2)"
B2_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_decode_launches.csv python bench.py --workload decode --steps 1 --warmup 1 --new-tokens 6 > gpurun_out/r2c_decode_ncu.log 2>&1; echo "ncu rc $?"
s=$(date +%s)
timeout 1100 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s --timeout 900 -p no:cacheprovider > gpurun_out/r2c_test_gpu_baseline_shapes.log 2>&1
r=$?; echo "== baseline shapes exit $r ($(( $(date +%s) - s )) s)"; grep -E "passed|failed|error" gpurun_out/r2c_test_gpu_baseline_shapes.log | tail -3
[ $r -ne 0 ] && { rc=$r; grep -E "^E " gpurun_out/r2c_test_gpu_baseline_shapes.log | head -20; }
grep -h "\[parity\]\|\[oracle\]" gpurun_out/r2c_test_gpu_baseline_shapes.log | cut -c1-250
exit $rc
