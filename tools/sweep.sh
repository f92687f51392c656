python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for sc in 0 1 0 1; do
echo "EMIT_SCRATCH=$sc"; ZPLT_EMIT_SCRATCH=$sc python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
