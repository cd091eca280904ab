#!/bin/bash
# round-2 GPU run #4: full GPU test suite exactly as the driver runs it, default bench line, lm_head rasterisation A/B, ncu evidence
mkdir -p gpurun_out
s=$(date +%s)
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2d_pytest_gpu.log 2>&1
rc=$?; echo "== pytest -m gpu exit $rc ($(( $(date +%s) - s )) s)"; tail -4 gpurun_out/r2d_pytest_gpu.log
[ $rc -ne 0 ] && grep -E "^E |Error|b2:" gpurun_out/r2d_pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench default"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc $?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_frac_of_peak")}, d["roofline"]["achieved"], d["roofline"]["attention"], d["e2e"]["value"])
    print("decode", {k: d["decode"][k] for k in ("value", "ms_per_token_step")}, d["decode"]["roofline"]["frac"], d["decode"]["e2e"]["value"])
    print("precise", d.get("precise")); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2d_bench.err").read()[-3000:])
P
pre() { timeout 300 python bench.py --steps 6 --warmup 3 --no-decode --no-precise --no-cpu-baseline --event-every 1 "$@" 2>gpurun_out/r2d_pre.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['roofline']['attention']
    print(round(d['value']), 'tok/s', round(d['ms_per_step'],3), 'ms; attn ms/step', round(a['ms_per_step'],3), 'gemm TF', round(d['roofline']['achieved'],1), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],3))
except Exception as e: print('failed', e); print(open('gpurun_out/r2d_pre.err').read()[-1500:])"; }
for rep in 1 2; do
  echo "== lm_head one band + cache hints (rep $rep)"; pre
  echo "== lm_head banded rasterisation (rep $rep)"; B2_LMHEAD_BANDS=1 pre
done
echo "== gemma2-2b S=4096 B=4 (reference flags / hf flags)"; pre --model gemma2-2b --seq 4096 --batch 4 --steps 4; pre --model gemma2-2b --seq 4096 --batch 4 --steps 4 --hf-flags
echo "== decode P=2048"; timeout 300 python bench.py --workload decode --prompt 2048 --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/r2d_decode_p2048.json | cut -c1-400
echo "== reference arm (short)"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/r2d_reference_arm.json | cut -c1-900
tools/r2_ncu.sh
exit $rc
