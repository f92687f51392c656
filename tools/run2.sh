TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29560
for cfg in "1 0" "4 0" "4 96" "4 64" "8 0" "8 96" "2 0"; do set -- $cfg; port=$((port+1))
echo "GROUPS=$1 P2P_CTAS=$2"; ZPLT_SLAB_GROUPS=$1 ZPLT_P2P_CTAS=$2 $TR --master-port $port bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/err_$1_$2.log | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print(d['ms_per_step'], d['stage_ms'], d['all_to_all']['ms'])
except Exception as e: print('FAILED', t[:200])"
done
tail -5 gpurun_out/err_4_64.log
