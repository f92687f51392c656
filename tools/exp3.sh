#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})'
$B 2>gpurun_out/e3_a.err | python -c "$pick" qplt
$B --za 2>gpurun_out/e3_b.err | python -c "$pick" za
$B --icformat RVdoubleZel 2>gpurun_out/e3_c.err | python -c "$pick" dbl
