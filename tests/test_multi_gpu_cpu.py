"""Host-side logic of the multi-GPU paths, run on CPU: frame sharding and row-band planning, plus a
world_size-2 gloo run that moves halo rows between ranks exactly as the plan says and checks that
every rank's window reproduces the oracle rows of the whole image."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oclcomputervision_b200 import multi_gpu as mg
from oclcomputervision_b200 import synth
from oracle import raisr_oracle as O


def test_shard_frames_covers_batch_once():
    for n in (1, 7, 64, 512, 2048):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = mg.shard_frames(n, r, world)
                seen += list(range(s, s + c))
            assert seen == list(range(n))
            counts = [mg.shard_frames(n, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


@pytest.mark.parametrize("sh,s,world", [(16384, 3, 8), (2160, 2, 4), (100, 2, 3), (9, 4, 2), (3, 2, 8)])
def test_band_plan(sh, s, world):
    bands = mg.band_plan(sh, s, world)
    rows = sum(b.dst_rows for b in bands)
    assert rows == sh * s
    for b in bands:
        if b.dst_rows == 0:
            continue
        assert b.dst_row0 == b.own_first * s and b.dst_rows == (b.own_last - b.own_first + 1) * s
        assert 0 <= b.src_first <= b.own_first and b.own_last <= b.src_last <= sh - 1
        assert b.halo_above <= 3 and b.halo_below <= 3          # <= 3 halo source rows per side
        got = []
        for peer, lo, hi in mg.halo_sources(bands, b.rank):
            assert peer != b.rank and bands[peer].own_first <= lo <= hi <= bands[peer].own_last
            got += list(range(lo, hi + 1))
        want = list(range(b.src_first, b.own_first)) + list(range(b.own_last + 1, b.src_last + 1))
        assert got == want
    if (sh, s, world) == (16384, 3, 8):
        assert (bands[1].src_first, bands[1].src_last) in ((2046, 4098), (2046, 4097), (2045, 4098))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sh, sw, s, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = synth.synthetic_frame(sh, sw, 5, sigma=2.0)          # every rank can regenerate the truth
        bands = mg.band_plan(sh, s, world)
        me = bands[rank]
        own = torch.from_numpy(img[me.own_first:me.own_last + 1].copy())
        win = torch.zeros((me.src_last - me.src_first + 1, sw), dtype=torch.uint8)
        win[me.own_first - me.src_first: me.own_last - me.src_first + 1] = own
        # serve my rows to the peers that need them, fetch mine: same plan as the P2P path
        reqs = []
        for other in bands:
            if other.rank == rank:
                continue
            for peer, lo, hi in mg.halo_sources(bands, other.rank):
                if peer == rank:
                    reqs.append(dist.isend(own[lo - me.own_first: hi - me.own_first + 1].contiguous(), dst=other.rank))
        for peer, lo, hi in mg.halo_sources(bands, rank):
            buf = torch.zeros((hi - lo + 1, sw), dtype=torch.uint8)
            dist.recv(buf, src=peer)
            win[lo - me.src_first: hi - me.src_first + 1] = buf
        for r in reqs:
            r.wait()
        ok = bool(np.array_equal(win.numpy(), img[me.src_first:me.src_last + 1]))
        # the band's rows of the oracle on the whole image only depend on the window: recompute them
        # from a padded copy where rows outside the window are garbage
        F = synth.random_filters(s)
        full = O.raisr_ref_c(img, F, s, want=("out_u8",))["out_u8"]
        dirty = np.full_like(img, 77)
        dirty[me.src_first:me.src_last + 1] = win.numpy()
        part = O.raisr_ref_c(dirty, F, s, want=("out_u8",))["out_u8"]
        ok = ok and bool(np.array_equal(part[me.dst_row0:me.dst_row0 + me.dst_rows], full[me.dst_row0:me.dst_row0 + me.dst_rows]))
        gathered = [None] * world
        dist.all_gather_object(gathered, ok)
        if rank == 0:
            q.put(all(gathered))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sh,s", [(2, 48, 2), (3, 45, 3)])
def test_halo_exchange_plan_world_size_n_gloo(world, sh, s):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sh, 40, s, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
