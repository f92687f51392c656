#!/bin/bash
# 2-GPU sanity after kernel changes: slab tests on one GPU, slab run vs single GPU, the distributed CLI, and the bench line
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "slab" 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29521 tools/run_slab.py --ppd 256 --p2p 2>&1 | grep -i "slab run\|error\|Traceback" | head
bash tools/test_mgpu_cli.sh 2>&1 | tail -2
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | cut -c1-200,560-1000
