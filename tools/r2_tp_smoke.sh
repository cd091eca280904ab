#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --tp 2 --model 1b --seq 2048 --batch 2 --steps 3 --warmup 1 --no-precise --no-cpu-baseline --event-every 1 > gpurun_out/r2_tp_smoke.json 2> gpurun_out/r2_tp_smoke.err
echo "rc $?"; tail -c 2500 gpurun_out/r2_tp_smoke.json; grep -v "^  File\|^    " gpurun_out/r2_tp_smoke.err | tail -8 | cut -c1-300
