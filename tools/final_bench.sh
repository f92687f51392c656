python bench.py 2>gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_1024_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
head -12 gpurun_out/launches_1024_final.csv | cut -c1-260
