"""Row-banded upscale across 2+ GPUs with the halo rows read from the neighbour's memory over
NVLink P2P (CUDA IPC).  Needs >= 2 GPUs; skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sh, sw, s, q):
    import torch.distributed as dist
    from oclcomputervision_b200 import ClRaisr, synth
    from oclcomputervision_b200 import multi_gpu as mg
    from oracle import raisr_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        F = synth.random_filters(s)
        r = ClRaisr(1, device=rank)
        setattr(r, "filters_x%d" % s, F)
        img = synth.synthetic_frame(sh, sw, 5, sigma=2.0)
        up = mg.BandedUpscaler(r, sw, sh, s)
        me = up.me
        out = up.upsample_band(img[me.own_first:me.own_last + 1])
        ref = O.raisr_ref_c(img, F, s, want=("out_u8", "hash"))
        mine = r.debug_hash(img, s)[0]
        same = (mine == ref["hash"])[me.dst_row0:me.dst_row0 + me.dst_rows]
        diff = np.abs(out.astype(int) - ref["out_u8"][me.dst_row0:me.dst_row0 + me.dst_rows].astype(int))
        ok = bool(diff[same].max() <= 1) and same.mean() > 0.999 and (world == 1 or up.halo_bytes > 0)
        # a second call re-uses the windows: the write-after-read hand-shake (done words) must let it through
        out2 = up.upsample_band(img[me.own_first:me.own_last + 1])
        ok = ok and bool(np.array_equal(out, out2))
        dist.barrier()            # nobody frees a window that a neighbour may still be reading
        up.close()
        r.close()
        res = [None] * world
        dist.all_gather_object(res, ok)
        if rank == 0:
            q.put(all(res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("sh,sw,s", [(300, 256, 2), (200, 192, 3)])
def test_banded_upscale_p2p_halo(sh, sw, s):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sh, sw, s, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
