ncu --set full --clock-control none --import-source on -k "regex:fft_tile|fft_emit|gen_xfft" -s 3 -c 3 -o gpurun_out/prof_r01_v3_1024 -f python bench.py --ppd 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_under_ncu4.log 2>&1
tail -2 gpurun_out/b_under_ncu4.log | cut -c1-300
