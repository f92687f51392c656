#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29523 tools/sweep.py --ppd 1024 2048 --steps 4 2>/dev/null | grep "^{" | tee gpurun_out/c22_sweep_n8.jsonl
