#!/bin/bash
# 4 GPUs (never run by hand before): real-IPC records check + the bench line with parity
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29522 tools/run_slab.py --ppd 1024 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee gpurun_out/c19_slab.log
timeout 600 $TR --master-port 29524 bench.py --gpus 4 --steps 10 --warmup 3 2>gpurun_out/c19_bench_n4.err > gpurun_out/c19_bench_n4.json; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/c19_bench_n4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['all_to_all']['nvlink_gbs_per_gpu'], d['e2e']['value'], d['parity'])"
