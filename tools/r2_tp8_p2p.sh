#!/bin/bash
# 8-GPU confirmation of the peer-memory all-reduce: decode TP=8 with NCCL only, then with the fused kernel
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 8 "$@"; }
for v in nccl p2p; do
  if [ $v = nccl ]; then export B2_TP_NO_P2P=1; else unset B2_TP_NO_P2P; fi
  timeout 200 bash -c "$(declare -f tr); tr --tp 8 --workload decode --steps 3 --warmup 1" 2>gpurun_out/r2_tp8_${v}.err | tee gpurun_out/r2_tp8_decode_${v}.json | python -c "
import json,sys
txt=[l for l in sys.stdin.read().splitlines() if l.startswith('{')]
try:
    d=json.loads(txt[-1]); print('$v', round(d['value']), 'tok/s', round(d['ms_per_token_step'],4), 'ms/token step')
except Exception as e: print('$v failed', e); print(open('gpurun_out/r2_tp8_${v}.err').read()[-1500:])"
done
