set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29600+n)) tools/pcie_ceiling.py 2>/dev/null | grep '^{' >> gpurun_out/r2f_pcie.jsonl
done
$TR --nproc-per-node 8 --master-port 29620 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/r2f_n8.err | grep '^{' > gpurun_out/r2f_bench_n8.json
$TR --nproc-per-node 4 --master-port 29621 bench.py --gpus 4 --steps 10 --warmup 3 2>gpurun_out/r2f_n4.err | grep '^{' > gpurun_out/r2f_bench_n4.json
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29630+n)) bench.py --gpus $n --config 5 --steps 5 --warmup 3 --no-extras 2>/dev/null | grep '^{' >> gpurun_out/r2f_cfg5_sweep.jsonl
done
python -m pytest tests/test_multi_gpu_gpu.py -m gpu -q 2>&1 | tail -3
wc -c gpurun_out/r2f_*
