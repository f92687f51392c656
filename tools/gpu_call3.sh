#!/bin/bash
# 2 GPUs: real NVLink/IPC path — slab run vs single-GPU run (bit-identical records), distributed CLI vs CLI, bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/run_slab.py --ppd 256 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee gpurun_out/c3_slab.log
timeout 600 $TR --master-port 29522 tools/run_slab.py --ppd 1024 --p2p 2>&1 | grep -i "slab run\|error\|Traceback\|assert" | head -5 | tee -a gpurun_out/c3_slab.log
timeout 300 bash tools/test_mgpu_cli.sh 2>&1 | tail -3 | tee -a gpurun_out/c3_slab.log
timeout 900 $TR --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/c3_bench_n2.err > gpurun_out/c3_bench_n2.json; echo "bench rc=$?"
cut -c1-3000 gpurun_out/c3_bench_n2.json
for o in "slab_ring=0" "p2p_ctas=0" "p2p_ctas=64" "p2p_ctas=128" "slab_groups=4" "slab_groups=16"; do
  k=${o%%=*}; v=${o##*=}; K=$(echo $k | tr a-z A-Z)
  env ZPLT_$K=$v timeout 300 $TR --master-port 29524 bench.py --gpus 2 --steps 5 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$o', d['ms_per_step'], d['stage_ms'], d['all_to_all']['nvlink_gbs_per_gpu'])" | tee -a gpurun_out/c3_variants.log
done
