for ps in 0 1; do for pf in 0 1; do for c in 0; do
echo "PERSIST=$ps PREFETCH=$pf"; ZPLT_PERSIST=$ps ZPLT_PREFETCH=$pf python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done; done; done
