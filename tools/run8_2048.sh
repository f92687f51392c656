#!/bin/bash
# the north-star configuration on one 8 x B200 box: PPD=2048 qPLT+rescale RVZel, device-resident bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29524 bench.py --gpus 8 --ppd 2048 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_n8_2048.err | tee gpurun_out/bench_n8_ppd2048.json | cut -c1-200,600-1100
tail -n 2 gpurun_out/bench_n8_2048.err
