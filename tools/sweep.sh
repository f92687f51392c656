for d in 0 148 296 592 128; do
pf=$(( d*4 + (d>0 ? 1 : 0) ))
echo "z-pass prefetch distance $d (flag $pf)"; ZPLT_PREFETCH=$pf python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
