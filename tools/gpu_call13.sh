#!/bin/bash
# 8 GPUs: the scaling line with the PPD=2048 sub-record, a records check against one GPU, and the tuning sweep
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nproc > gpurun_out/c13_host.txt; free -g >> gpurun_out/c13_host.txt; nvidia-smi topo -m >> gpurun_out/c13_host.txt 2>&1
timeout 900 $TR --master-port 29524 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/c13_bench_n8.err > gpurun_out/c13_bench_n8.json; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/c13_bench_n8.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['all_to_all']['nvlink_gbs_per_gpu'], d['e2e']['value'], d['parity']['ok']); print(json.dumps(d.get('ppd2048'))[:1500])"
timeout 600 $TR --master-port 29523 tools/sweep.py --ppd 1024 2048 2>/dev/null | grep "^{" | tee gpurun_out/c13_sweep_n8.jsonl
timeout 600 $TR --master-port 29522 tools/run_slab.py --ppd 1024 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee gpurun_out/c13_slab.log
