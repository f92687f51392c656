for pf in 2; do for tt in 8 4; do for ps in 0 1; do
echo "copy-only PREFETCH=$pf TILE_T=$tt PERSIST=$ps"; ZPLT_PREFETCH=$pf ZPLT_TILE_T=$tt ZPLT_PERSIST=$ps python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done; done; done
