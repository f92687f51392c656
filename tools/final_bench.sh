#!/bin/bash
# the round's evidence run on one B200: default bench line, ncu launch list of the same command, one full capture of the three hot kernels
mkdir -p gpurun_out
python bench.py 2>gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_1024_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
head -14 gpurun_out/launches_1024_final.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"gen_xfft|fft_tile_ring|fft_emit_ring" -c 3 -o gpurun_out/prof_r01_final2_1024 -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/prof_final2.log 2>&1
ls -la gpurun_out/prof_r01_final2_1024.ncu-rep
