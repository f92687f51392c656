#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
grep -v "^$" gpurun_out/c4_pytest.log | tail -14
for o in "yring=12" "yring=0" "yring=0 --opt dit2048_emit=1"; do
  timeout 300 python tools/slab1.py --ppd 2048 --ranks 8 --opt $o --tag "2048 $o" 2>/dev/null | tail -1 >> gpurun_out/c4_slab1.jsonl
done
cat gpurun_out/c4_slab1.jsonl
B="python bench.py --steps 5 --warmup 2 --no-e2e --no-parity --no-cpu-baseline --no-files"
timeout 300 $B --icformat RVdoubleZel 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RVdoubleZel', d['ms_per_step'], d['stage_ms'])" | tee gpurun_out/c4_variants.log
timeout 300 $B --za 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZA RVZel', d['ms_per_step'], d['stage_ms'])" | tee -a gpurun_out/c4_variants.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/c4_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['e2e_files'], d['parity']['ok'])"
