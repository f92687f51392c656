#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29523 tools/sweep.py --ppd 1024 --steps 5 2>/dev/null | grep "^{" | tee gpurun_out/c21_sweep_n4.jsonl
