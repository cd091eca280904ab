#!/bin/bash
# round-2 GPU run #5: JSON lines of every figure DESIGN.md quotes + the ncu evidence set (kept under 64 MiB)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc $?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "model_frac_of_peak")}, d["roofline"]["achieved"], d["roofline"]["attention"], d["e2e"]["value"])
print("decode", {k: d["decode"][k] for k in ("value", "ms_per_token_step")}, d["decode"]["roofline"]["frac"], d["decode"]["e2e"]["value"])
P
timeout 400 python bench.py --model gemma2-2b --seq 4096 --batch 4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_gemma2_ref_flags.json 2>/dev/null; echo "gemma rc $?"
timeout 400 python bench.py --model gemma2-2b --seq 4096 --batch 4 --steps 4 --warmup 3 --no-cpu-baseline --hf-flags --no-precise > gpurun_out/r02_bench_gemma2_hf_flags.json 2>/dev/null; echo "gemma hf rc $?"
timeout 400 python bench.py --model 3b --seq 8192 --batch 2 --steps 4 --warmup 3 --no-cpu-baseline --no-decode --no-precise > gpurun_out/r02_bench_3b_s8192_b2_1gpu.json 2>/dev/null; echo "3b rc $?"
timeout 300 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/r02_bench_decode_p16.json 2>/dev/null; echo "decode rc $?"
timeout 300 python bench.py --workload decode --prompt 2048 --steps 3 --warmup 1 > gpurun_out/r02_bench_decode_p2048.json 2>/dev/null; echo "decode p2048 rc $?"
B2_NO_FUSED_DECODE=1 timeout 300 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/r02_bench_decode_p16_7kernel_path.json 2>/dev/null; echo "decode 7k rc $?"
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "reference arm rc $?"
for f in gpurun_out/r02_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d.get('metric'), round(d['value'],1), round(d['ms_per_step'],3), (d.get('precise') or {}).get('value'), (d.get('cpu_baseline_long') or {}))"; done
timeout 120 tools/mb/pipes 2>&1 | tail -30 > gpurun_out/r02_pipe_microbench.txt; tail -12 gpurun_out/r02_pipe_microbench.txt
tools/r2_ncu.sh prefill
