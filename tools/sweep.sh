python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --icformat RVdoubleZel --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RVdouble', d['ms_per_step'], d['stage_ms'])"
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RVZel', d['ms_per_step'], d['stage_ms'])"
