for mb in 104 140 208 312 416; do
  python bench.py --steps 10 --no-cpu-baseline --chunk-mb $mb 2>&1 | tail -1 > gpurun_out/sw_$mb.json
  python -c "
import json; d=json.load(open('gpurun_out/sw_$mb.json')); print('chunk_mb', $mb, d['value'], d['e2e']['value'], d['roofline']['kernels'])"
done
