#!/bin/bash
# one GPU, the round's last seconds: PPD=1024 qPLT+rescale RVZel out of core in 4 passes through host memory (no files), then resident
mkdir -p gpurun_out
timeout 75 python tools/ooc_time.py --ppd 1024 --passes 4 0 2> gpurun_out/c27_ooc.err | tee gpurun_out/c27_ooc.jsonl
grep -i "out of core\|Block IO\|took" gpurun_out/c27_ooc.err | head
