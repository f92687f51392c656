#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), d["e2e"]["seconds_per_step"], d["e2e"]["value"])'
ZPLT_YRING=0 $B 2>gpurun_out/e4_a.err | python -c "$pick" yring0
$B 2>gpurun_out/e4_b.err | python -c "$pick" default
ZPLT_LIB=$PWD/zeldovich-plt_b200/libzplt_base.so $B 2>gpurun_out/e4_c.err | python -c "$pick" base
python tools/d2h_test.py 2>&1 | tail -3
