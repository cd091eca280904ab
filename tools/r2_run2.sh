#!/bin/bash
# round-2 GPU run #2: cluster split-K decode path, tests again, A/B of exp2 offload, decode variants
mkdir -p gpurun_out
rc=0
for t in tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_baseline_shapes.py; do
  name=$(basename "$t" .py)
  s=$(date +%s)
  timeout 1100 python -m pytest "$t" -q -m gpu -s --timeout 900 -p no:cacheprovider -x > "gpurun_out/r2b_${name}.log" 2>&1
  r=$?
  echo "== $t exit $r ($(( $(date +%s) - s )) s)"; grep -E "passed|failed|error" "gpurun_out/r2b_${name}.log" | tail -3
  [ $r -ne 0 ] && { rc=$r; grep -E "^E |Error|error:|b2:" "gpurun_out/r2b_${name}.log" | head -30; }
done
grep -h "\[parity\]\|\[oracle\]" gpurun_out/r2b_test_gpu_baseline_shapes.log | cut -c1-250
dec() { timeout 300 python bench.py --workload decode --steps 3 --warmup 1 "$@" 2>gpurun_out/r2b_dec.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'tok/s', round(d['ms_per_token_step'],4), 'ms/step frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], 'e2e', round(d['e2e']['value']))
except Exception as e: print('decode bench failed', e); print(open('gpurun_out/r2b_dec.err').read()[-1500:])"; }
echo "== decode fused (default)"; dec
echo "== decode 7-kernel path"; B2_NO_FUSED_DECODE=1 dec
echo "== decode fused, per-head attention kernel for short caches"; B2_DECODE_SPLIT_MIN_T=256 dec
echo "== decode fused P=2048"; dec --prompt 2048
echo "== decode 7-kernel P=2048"; B2_NO_FUSED_DECODE=1 dec --prompt 2048
echo "== bench default"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc $?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_frac_of_peak")}, d["roofline"]["achieved"], d["roofline"]["attention"], d["e2e"]["value"])
    print("decode", {k: d["decode"][k] for k in ("value", "ms_per_token_step")}, d["decode"]["roofline"]["frac"], d["decode"]["e2e"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2b_bench.err").read()[-3000:])
P
echo "== A/B exp2 offload"
cp llm_np_cp_b200/libb200fwd.so /tmp/keep.so
for rep in 1 2; do
  for f in ab_base.so ab_poly8.so ab_poly4.so; do
    cp "$f" llm_np_cp_b200/libb200fwd.so
    timeout 300 python bench.py --steps 6 --warmup 3 --no-decode --no-precise --no-cpu-baseline --event-every 1 2>gpurun_out/r2b_ab.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['roofline']['attention']
    print('$f rep $rep', round(d['ms_per_step'],3), 'attn ms/step', round(a['ms_per_step'],3), 'attn TF', round(a['achieved'],1))
except Exception as e: print('$f failed', e); print(open('gpurun_out/r2b_ab.err').read()[-800:])"
  done
done
cp /tmp/keep.so llm_np_cp_b200/libb200fwd.so
echo "== decode launch list (fused path)"
B2_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_decode_launches.csv python bench.py --workload decode --steps 1 --warmup 1 --new-tokens 6 > gpurun_out/r2b_decode_ncu.log 2>&1; echo "ncu rc $?"
exit $rc
