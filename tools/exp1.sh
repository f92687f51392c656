#!/bin/bash
# round-1 experiment: run-of-4 generation phase + TMA ring-prefetched z pass, A/B against the previous build
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ZPLT_ZRING=12 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})'
ZPLT_LIB=$PWD/zeldovich-plt_b200/libzplt_base.so $B 2>gpurun_out/e1_base.err | python -c "$pick" base
$B 2>gpurun_out/e1_new.err | python -c "$pick" new
for k in 4 8 12; do ZPLT_ZRING=$k $B 2>gpurun_out/e1_ring$k.err | python -c "$pick" ring$k; done
ZPLT_GENX_T=8 $B 2>gpurun_out/e1_t8.err | python -c "$pick" genxT8
tail -3 gpurun_out/e1_*.err
