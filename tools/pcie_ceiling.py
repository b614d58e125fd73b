#!/usr/bin/env python
"""Raw host<->device copy ceiling of the box, in the layout bench.py runs in (one process per GPU under torchrun):
every rank moves the bytes of one bench step -- H2D of the source frames and D2H of the upscaled frames -- with bare
cudaMemcpyAsync from / to pinned memory on two streams, all ranks at once, no kernels.  One JSON line (rank 0).

    python tools/pcie_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

The end-to-end number of bench.py (`e2e`) cannot exceed min(device-resident rate, this ceiling); bench.py measures the
same thing inline (`e2e.copy_ceiling`).  Also reports each direction alone."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h2d, d2h = 32 * 3840 * 2160, 32 * 7680 * 4320          # bench default: 32 frames 4K -> 8K, u8
    hs, hd = torch.empty(h2d, dtype=torch.uint8).pin_memory(), torch.empty(d2h, dtype=torch.uint8).pin_memory()
    ds, dd = torch.empty(h2d, dtype=torch.uint8, device="cuda"), torch.empty(d2h, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up, down, reps=5):
        def once():
            if up:
                with torch.cuda.stream(s1):
                    ds.copy_(hs, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hd.copy_(dd, non_blocking=True)
        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    both, up, down = run(True, True), run(True, False), run(False, True)
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, h2d_bytes_per_rank=h2d, d2h_bytes_per_rank=d2h,
                              both_directions_gbs=round(world * (h2d + d2h) / both / 1e9, 1), both_ms=round(both * 1e3, 3),
                              h2d_alone_gbs=round(world * h2d / up / 1e9, 1), d2h_alone_gbs=round(world * d2h / down / 1e9, 1),
                              as_output_mpix_s=round(world * d2h / both / 1e6, 1),
                              note="aggregate over all ranks, max-over-ranks time; as_output_mpix_s = the bench metric this ceiling allows")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
