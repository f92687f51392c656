#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})'
ZPLT_LIB=$PWD/zeldovich-plt_b200/libzplt_base.so $B 2>/dev/null | python -c "$pick" base
$B 2>/dev/null | python -c "$pick" slices
$B --za 2>/dev/null | python -c "$pick" slices_za
