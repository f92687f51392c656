#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})'
ZPLT_LIB=$PWD/zeldovich-plt_b200/libzplt_base.so $B 2>/dev/null | python -c "$pick" base
$B 2>/dev/null | python -c "$pick" evict_last
$B --za 2>/dev/null | python -c "$pick" evict_last_za
ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum -k regex:emit_ring -c 1 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e 2>/dev/null | grep -i "emit_ring" | cut -d, -f5,13-15 | cut -c1-300
