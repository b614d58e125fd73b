for f in 0 1 2 4 8 16 32 33 63; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dbg-flags $f 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags', $f, d['roofline']['kernels'])"
done
