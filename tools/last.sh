#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
pick='import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})'
ZPLT_WIDE_RECORDS=1 $B 2>/dev/null | python -c "$pick" wide
$B 2>/dev/null | python -c "$pick" base
ZPLT_WIDE_RECORDS=1 python -m pytest tests -x -q -m gpu -k "golden or oracle or full_size" 2>&1 | tail -2
