#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"gen_xfft|fft_tile_ring" -c 2 -o gpurun_out/prof_r01_ring_genrun -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/prof2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
