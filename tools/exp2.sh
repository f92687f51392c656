#!/bin/bash
# experiment 2: persistent ring-prefetched y pass + emission, A/B
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})'
ZPLT_YRING=0 $B 2>gpurun_out/e2_y0.err | python -c "$pick" yring0
$B 2>gpurun_out/e2_y12.err | python -c "$pick" yring12
ZPLT_YRING=0 $B --icformat RVdoubleZel 2>gpurun_out/e2_dbl0.err | python -c "$pick" dbl_yring0
$B --icformat RVdoubleZel 2>gpurun_out/e2_dbl12.err | python -c "$pick" dbl_yring12
ZPLT_YRING=0 $B --za 2>gpurun_out/e2_za0.err | python -c "$pick" za_yring0
$B --za 2>gpurun_out/e2_za12.err | python -c "$pick" za_yring12
for f in gpurun_out/e2_*.err; do echo $f; tail -n 2 $f; done
