#!/bin/bash
# first GPU call of round 2: the whole GPU test suite, then 2048-kernel timings on one rank, then the bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/c1_gpu.txt; nproc >> gpurun_out/c1_gpu.txt; free -g >> gpurun_out/c1_gpu.txt; df -h /tmp >> gpurun_out/c1_gpu.txt
timeout 2400 python -m pytest tests -q -m gpu -x --durations=25 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -40 gpurun_out/c1_pytest.log
for o in "dit2048=0" "dit2048=1" "dit2048=1 --opt p2p_ctas=0" "dit2048=0 --opt p2p_ctas=0"; do
  timeout 300 python tools/slab1.py --ppd 2048 --ranks 8 --opt $o --tag "$o" 2>/dev/null | tail -1 >> gpurun_out/c1_slab1.jsonl
done
for o in "slab_ring=1" "slab_ring=0" "slab_ring=1 --opt p2p_ctas=0" "slab_ring=1 --opt yring=0"; do
  timeout 300 python tools/slab1.py --ppd 1024 --ranks 8 --opt $o --tag "1024 $o" 2>/dev/null | tail -1 >> gpurun_out/c1_slab1.jsonl
  timeout 300 python tools/slab1.py --ppd 1024 --ranks 2 --opt $o --tag "1024/2 $o" 2>/dev/null | tail -1 >> gpurun_out/c1_slab1.jsonl
done
cat gpurun_out/c1_slab1.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/c1_bench.json
