#!/bin/bash
# 2 real GPUs, per-source receive layout forced (what 4 and 8 ranks use by default): records against one GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ZPLT_B2_LAYOUT=1 timeout 200 $TR --master-port 29521 tools/run_slab.py --ppd 512 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee gpurun_out/c25_slab.log
