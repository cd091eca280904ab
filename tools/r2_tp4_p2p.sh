#!/bin/bash
# 4-GPU check of the peer-memory all-reduce (loads pipelined): TP tests at world 2 and 4, decode TP=4 NCCL vs fused
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tp.py -q -m gpu -s --timeout 350 -p no:cacheprovider > gpurun_out/r2_tp4_tests.log 2>&1
r=$?; echo "== tp tests exit $r"; grep -E "passed|failed|error|parity-tp|b2:" gpurun_out/r2_tp4_tests.log | tail -8
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 4 "$@"; }
for v in nccl p2p; do
  if [ $v = nccl ]; then export B2_TP_NO_P2P=1; else unset B2_TP_NO_P2P; fi
  timeout 200 bash -c "$(declare -f tr); tr --tp 4 --workload decode --steps 3 --warmup 1" 2>gpurun_out/r2_tp4_${v}.err | tee gpurun_out/r2_tp4_decode_${v}.json | python -c "
import json,sys
txt=[l for l in sys.stdin.read().splitlines() if l.startswith('{')]
try:
    d=json.loads(txt[-1]); print('$v', round(d['value']), 'tok/s', round(d['ms_per_token_step'],4), 'ms/token step')
except Exception as e: print('$v failed', e); print(open('gpurun_out/r2_tp4_${v}.err').read()[-1500:])"
done
