python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k "regex:fft_tile|fft_emit|gen_xfft" -s 3 -c 3 -o gpurun_out/prof_r01_final_1024 -f python bench.py --ppd 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_under_ncu6.log 2>&1
