python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k "regex:gen_xfft" -s 1 -c 1 -o gpurun_out/prof_r01_v4_genx -f python bench.py --ppd 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_under_ncu5.log 2>&1
