#!/bin/bash
# one GPU: whole GPU suite, smoke, bench line, ncu launch list + full capture of the three hot kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
grep -v "^$" gpurun_out/c17_pytest.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/c17_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['e2e_files']['seconds'], d['parity']['ok'], d['cpu_baseline']['value'])"
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-parity --no-cpu-baseline --no-files"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_ppd1024.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gen_xfft|fft_tile_ring|fft_emit_ring" -s 3 -c 3 -f -o gpurun_out/prof_r02_1024 $B > /dev/null 2>&1
ls -la gpurun_out/prof_r02_1024.ncu-rep gpurun_out/r02_launches_ppd1024.csv
