#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
grep -v "^$" gpurun_out/c5_pytest.log | tail -12
B="python bench.py --steps 5 --warmup 2 --no-e2e --no-parity --no-cpu-baseline --no-files"
timeout 300 $B --icformat RVdoubleZel 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RVdoubleZel', d['ms_per_step'], d['stage_ms'])" | tee gpurun_out/c5_variants.log
timeout 300 $B --icformat RVdoubleZel --za 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZA RVdoubleZel', d['ms_per_step'], d['stage_ms'])" | tee -a gpurun_out/c5_variants.log
