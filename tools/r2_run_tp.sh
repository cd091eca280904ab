#!/bin/bash
# multi-GPU run (gpurun --gpus N): TP / DP tests, bench --tp N (prefill 3B S=8192 one sequence, decode B=64), DP scaling lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}.txt 2>&1
rc=0
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_tp.py tests/test_gpu_dp.py -q -m gpu -s --timeout 800 -p no:cacheprovider > gpurun_out/r2_tp_tests_${N}.log 2>&1
r=$?; echo "== tp/dp tests exit $r"; grep -E "passed|failed|error|parity-tp" gpurun_out/r2_tp_tests_${N}.log | tail -8
[ $r -ne 0 ] && { rc=$r; tail -40 gpurun_out/r2_tp_tests_${N}.log; }
fi
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N "$@"; }
show() { python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
try:
    d=json.loads([l for l in txt if l.startswith('{')][-1])
    keys=('metric','value','ms_per_step','n_gpus','scaling')
    print({k:d.get(k) for k in keys}, 'tp', d.get('tp'), 'frac', d.get('model_frac_of_peak'), 'roofline', (d.get('roofline') or {}).get('frac'), 'e2e', d['e2e']['value'])
except Exception as e: print('failed', e, txt[-5:])"; }
echo "== TP=$N prefill 3B S=8192 B=1"
NCCL_DEBUG=WARN timeout 300 bash -c "$(declare -f tr); N=$N; tr --tp $N --model 3b --seq 8192 --batch 1 --steps 4 --warmup 2 --no-precise --no-cpu-baseline --event-every 1" 2>gpurun_out/r2_tp_prefill_${N}.err | tee gpurun_out/r2_tp_prefill_${N}.json | show
if [ "$N" = "2" ]; then
echo "== 1 GPU prefill 3B S=8192 B=1 (same work on one GPU)"
timeout 300 python bench.py --model 3b --seq 8192 --batch 1 --steps 4 --warmup 2 --no-precise --no-cpu-baseline --no-decode 2>/dev/null | tee gpurun_out/r2_tp_prefill_1gpu_ref.json | show
fi
echo "== TP=$N decode 1B B=64, NCCL all-reduce only (B2_TP_NO_P2P=1)"
B2_TP_NO_P2P=1 timeout 240 bash -c "$(declare -f tr); N=$N; tr --tp $N --workload decode --steps 3 --warmup 1" 2>/dev/null | tee gpurun_out/r2_tp_decode_${N}_nccl.json | show
echo "== TP=$N decode 1B B=64 (fused all-reduce + RMSNorm over peer memory)"
timeout 240 bash -c "$(declare -f tr); N=$N; tr --tp $N --workload decode --steps 3 --warmup 1" 2>gpurun_out/r2_tp_decode_${N}.err | tee gpurun_out/r2_tp_decode_${N}.json | show
if [ "$N" != "8" ]; then
echo "== DP x$N prefill 1B (default line)"
timeout 300 bash -c "$(declare -f tr); N=$N; tr --steps 6 --warmup 3 --no-precise" 2>gpurun_out/r2_dp_${N}.err | tee gpurun_out/r2_dp_${N}.json | show
fi
if [ "$N" = "8" ]; then
  echo "== DP x8 config #3: 3B S=8192, 2 sequences per GPU"
  timeout 300 bash -c "$(declare -f tr); N=$N; tr --model 3b --seq 8192 --batch 2 --steps 4 --warmup 2 --no-precise --no-decode" 2>gpurun_out/r2_dp8_3b.err | tee gpurun_out/r2_dp8_3b.json | show
fi
tail -5 gpurun_out/r2_tp_prefill_${N}.err gpurun_out/r2_tp_decode_${N}.err 2>/dev/null | cut -c1-300
exit $rc
