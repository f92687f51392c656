python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['stage_gbs'])"
python bench.py --za --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ZA', d['ms_per_step'], d['stage_ms'], d['stage_gbs'])"
python bench.py --icformat RVdoubleZel --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RVdouble', d['ms_per_step'], d['stage_ms'], d['stage_gbs'])"
