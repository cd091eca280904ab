#!/bin/bash
# 2-GPU check of the peer-memory all-reduce path: TP tests, then decode TP=2 with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tp.py -q -m gpu -s --timeout 500 -p no:cacheprovider > gpurun_out/r2_tp_p2p_tests.log 2>&1
r=$?; echo "== tp tests exit $r"; grep -E "passed|failed|error|parity-tp|b2:" gpurun_out/r2_tp_p2p_tests.log | tail -8
[ $r -ne 0 ] && { tail -30 gpurun_out/r2_tp_p2p_tests.log | cut -c1-300; exit $r; }
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 "$@"; }
for v in nccl p2p nccl p2p; do
  if [ $v = nccl ]; then export B2_TP_NO_P2P=1; else unset B2_TP_NO_P2P; fi
  timeout 240 bash -c "$(declare -f tr); tr --tp 2 --workload decode --steps 3 --warmup 1" 2>gpurun_out/r2_tp_p2p_${v}.err | tee gpurun_out/r2_tp2_decode_${v}.json | python -c "
import json,sys
txt=[l for l in sys.stdin.read().splitlines() if l.startswith('{')]
try:
    d=json.loads(txt[-1]); print('$v', round(d['value']), 'tok/s', round(d['ms_per_token_step'],4), 'ms/token step')
except Exception as e: print('$v failed', e); print(open('gpurun_out/r2_tp_p2p_${v}.err').read()[-1200:])"
done
