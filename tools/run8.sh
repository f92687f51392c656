TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 tools/run_slab.py --ppd 256 --p2p 2>&1 | grep "slab run"
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_n8.err | tee gpurun_out/bench_n8_ppd1024.json | cut -c1-1800
timeout 900 $TR --master-port 29524 bench.py --gpus 8 --ppd 2048 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_n8_2048.err | tee gpurun_out/bench_n8_ppd2048.json | cut -c1-1800
tail -2 gpurun_out/bench_n8_2048.err
