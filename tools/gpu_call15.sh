#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/run_slab.py --ppd 256 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee gpurun_out/c15_slab.log
timeout 600 $TR --master-port 29522 tools/run_slab.py --ppd 1024 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee -a gpurun_out/c15_slab.log
timeout 600 $TR --master-port 29523 tools/sweep.py --ppd 1024 2>/dev/null | grep "^{" | tee gpurun_out/c15_sweep_n2.jsonl
