#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu --durations=30 > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
grep -v "^$" gpurun_out/c2_pytest.log | tail -60
for o in "dit2048=0" "dit2048=1"; do
  timeout 300 python tools/slab1.py --ppd 2048 --ranks 8 --opt $o --tag "$o" 2>/dev/null | tail -1 >> gpurun_out/c2_slab1.jsonl
done
cat gpurun_out/c2_slab1.jsonl
