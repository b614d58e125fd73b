# final-build evidence on an 8-GPU box: default bench line (with the band record) at 8 / 4 / 2 / 1 ranks
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29700+n)) bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/r2q_n$n.err | grep '^{' > gpurun_out/r2q_bench_n$n.json
done
python bench.py --steps 10 --warmup 3 2>gpurun_out/r2q_n1.err | grep '^{' > gpurun_out/r2q_bench_n1.json
python -m pytest tests/test_multi_gpu_gpu.py -m gpu -q 2>&1 | tail -2
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    d = json.load(open("gpurun_out/r2q_bench_n%d.json" % n))
    b = d["band"]
    print(n, d["value"], d["e2e"]["value"], d["e2e"]["frac_of_bound"], b.get("device_ms"), b.get("e2e_ms"), b.get("halo_bytes"), b.get("checksum"), b.get("error"))
PY
