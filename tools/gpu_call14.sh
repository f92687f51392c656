#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -k "fused_exchange or slab_decomposition or c5_rank" > gpurun_out/c14_pytest.log 2>&1; tail -6 gpurun_out/c14_pytest.log
for o in "p2p_resident=1" "p2p_resident=0" "p2p_resident=1 --opt p2p_ctas=64" "p2p_resident=1 --opt p2p_ctas=56"; do
  timeout 300 python tools/slab1.py --ppd 1024 --ranks 8 --opt $o --tag "1024/8 $o" 2>/dev/null | tail -1 >> gpurun_out/c14_slab1.jsonl
  timeout 300 python tools/slab1.py --ppd 1024 --ranks 2 --opt $o --tag "1024/2 $o" 2>/dev/null | tail -1 >> gpurun_out/c14_slab1.jsonl
done
for o in "p2p_resident=1" "p2p_resident=0" "p2p_resident=1 --opt p2p_ctas=88"; do
  timeout 300 python tools/slab1.py --ppd 2048 --ranks 8 --opt $o --tag "2048/8 $o" 2>/dev/null | tail -1 >> gpurun_out/c14_slab1.jsonl
done
cat gpurun_out/c14_slab1.jsonl
