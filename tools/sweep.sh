python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for gt in 4 8; do
echo GENX_T=$gt; ZPLT_GENX_T=$gt python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
