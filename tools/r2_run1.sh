#!/bin/bash
# round-2 GPU run #1: all GPU tests, the default bench line, an A/B of exp2-offload variants, decode launch list
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,power.limit --format=csv; nproc; free -g | head -2; } > gpurun_out/r2_info.txt 2>&1
rc=0
for t in tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_dp.py tests/test_loader_generate.py tests/test_gpu_baseline_shapes.py; do
  name=$(basename "$t" .py)
  s=$(date +%s)
  timeout 1100 python -m pytest "$t" -q -m gpu -s --timeout 900 -p no:cacheprovider > "gpurun_out/r2_${name}.log" 2>&1
  r=$?
  echo "== $t exit $r ($(( $(date +%s) - s )) s)"; grep -E "passed|failed|error" "gpurun_out/r2_${name}.log" | tail -3
  [ $r -ne 0 ] && rc=$r
done
grep -h "\[parity\]\|\[oracle\]\|\[setup\]" gpurun_out/r2_test_gpu_baseline_shapes.log | cut -c1-260
echo "== bench default"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc $?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_frac_of_peak")}, d["roofline"]["achieved"], d["roofline"]["attention"], d["e2e"]["value"])
    print("decode", {k: d["decode"][k] for k in ("value", "ms_per_token_step")}, d["decode"]["roofline"]["frac"], d["decode"]["e2e"]["value"])
    print("precise", d.get("precise")); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2_bench1.err").read()[-3000:])
P
echo "== A/B exp2 offload"
cp llm_np_cp_b200/libb200fwd.so /tmp/keep.so
for rep in 1 2; do
  for f in ab_base.so ab_poly8.so ab_poly4.so; do
    cp "$f" llm_np_cp_b200/libb200fwd.so
    timeout 300 python bench.py --steps 6 --warmup 3 --no-decode --no-precise --no-cpu-baseline --event-every 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['roofline']['attention']
print('$f rep $rep', round(d['ms_per_step'],3), 'attn ms/step', round(a['ms_per_step'],3), 'attn TF', round(a['achieved'],1))"
  done
done
cp /tmp/keep.so llm_np_cp_b200/libb200fwd.so
echo "== decode launch list"
B2_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_decode_launches.csv python bench.py --workload decode --steps 1 --warmup 1 --new-tokens 6 > gpurun_out/r2_decode_ncu.log 2>&1; echo "ncu rc $?"
exit $rc
