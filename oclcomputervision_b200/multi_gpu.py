"""Multi-GPU use of the RAISR path: one process per GPU, no data-path collective.

* batches of frames: every rank upsamples a contiguous slice of the batch (``shard_frames``);
* one very large image: every rank owns a band of source rows, produces the matching band of
  output rows and reads the <=3 halo source rows per side it needs from its neighbours' memory over
  NVLink peer-to-peer (CUDA IPC handles exchanged once through ``torch.distributed``; the copy is a
  peer ``cudaMemcpy2DAsync`` issued by the consumer -- "read once").

The reference is single-device (raisr.py:70-72), so none of this has a counterpart there; the only
reference-derived rule is that the coordinate map of raisr.cl:209 uses the GLOBAL image size.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Tuple

from . import _cabi


def shard_frames(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous (start, count) of the frames rank `rank` processes; counts differ by at most 1."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


@dataclass
class Band:
    rank: int
    own_first: int      # first / last (inclusive) global source row owned by the rank
    own_last: int
    dst_row0: int       # first global output row produced, number of output rows
    dst_rows: int
    src_first: int      # first / last (inclusive) global source row needed (owned + halo)
    src_last: int

    @property
    def halo_above(self) -> int:
        return self.own_first - self.src_first

    @property
    def halo_below(self) -> int:
        return self.src_last - self.own_last


def band_plan(global_sh: int, scale: int, world: int) -> List[Band]:
    """Row bands of a global image with `global_sh` source rows: rank r owns a contiguous run of
    source rows and produces exactly their `scale`x output rows; the halo is what
    raisr_band_src_rows (the coordinate map of raisr.cl:209 on global sizes, +-5 output rows of
    patch margin) says it needs beyond them."""
    lib = _cabi.load()
    bands = []
    for r in range(world):
        start, count = shard_frames(global_sh, r, world)
        if count == 0:
            bands.append(Band(r, start, start - 1, start * scale, 0, 0, -1))
            continue
        first, last = ctypes.c_int(), ctypes.c_int()
        _cabi.check(lib.raisr_band_src_rows(global_sh, scale, start * scale, count * scale,
                                            ctypes.byref(first), ctypes.byref(last)))
        bands.append(Band(r, start, start + count - 1, start * scale, count * scale, first.value, last.value))
    return bands


def halo_sources(bands: List[Band], rank: int) -> List[Tuple[int, int, int]]:
    """(peer_rank, first_row, last_row) global source-row ranges rank `rank` must read from peers."""
    me = bands[rank]
    out = []
    if me.dst_rows == 0:
        return out
    for lo, hi in ((me.src_first, me.own_first - 1), (me.own_last + 1, me.src_last)):
        row = lo
        while row <= hi:
            owner = next(b for b in bands if b.dst_rows and b.own_first <= row <= b.own_last)
            end = min(hi, owner.own_last)
            out.append((owner.rank, row, end))
            row = end + 1
    return out


class BandedUpscaler:
    """Row-banded upscale of one large image across the ranks of a torch.distributed group.

    Each rank calls ``upsample_band(own_rows_u8)`` with its owned source rows (a 2-D uint8 host
    array); it returns that rank's band of output rows as a host array.  Device buffers are plain
    cudaMalloc allocations so they can be shared through CUDA IPC.
    """

    def __init__(self, raisr, sw: int, global_sh: int, scale: int, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.raisr, self.sw, self.global_sh, self.scale = raisr, sw, global_sh, scale
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bands = band_plan(global_sh, scale, self.world)
        self.me = self.bands[self.rank]
        self.lib = _cabi.load()
        self.pitch = (sw + 255) // 256 * 256
        n_win = self.me.src_last - self.me.src_first + 1
        self.win = ctypes.c_void_p()
        _cabi.check(self.lib.raisr_dev_alloc(raisr._h, ctypes.byref(self.win), max(1, n_win) * self.pitch))
        self.out = ctypes.c_void_p()
        self.out_pitch = (sw * scale + 255) // 256 * 256
        _cabi.check(self.lib.raisr_dev_alloc(raisr._h, ctypes.byref(self.out), max(1, self.me.dst_rows) * self.out_pitch))
        # publish the window (owned rows live inside it) to the peers
        handle = (ctypes.c_ubyte * 64)()
        _cabi.check(self.lib.raisr_ipc_export(self.win, handle))
        mine = dict(handle=bytes(handle), first=self.me.src_first, pitch=self.pitch)
        if self.world > 1:
            infos = [None] * self.world
            dist.all_gather_object(infos, mine, group=group)
        else:
            infos = [mine]
        self.peers = {}
        for peer, lo, hi in halo_sources(self.bands, self.rank):
            if peer not in self.peers:
                ptr = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(infos[peer]["handle"])
                _cabi.check(self.lib.raisr_ipc_open(buf, ctypes.byref(ptr)))
                self.peers[peer] = (ptr, infos[peer]["first"], infos[peer]["pitch"])

    def _row_ptr(self, global_row: int) -> int:
        return self.win.value + (global_row - self.me.src_first) * self.pitch

    def upsample_band(self, own_rows):
        import numpy as np
        me = self.me
        assert own_rows.shape == (me.own_last - me.own_first + 1, self.sw) and own_rows.dtype == np.uint8
        own_rows = np.ascontiguousarray(own_rows)
        import torch
        # owned rows: host -> my window (the only H2D of source data)
        t = torch.from_numpy(own_rows).cuda()
        _cabi.check(self.lib.raisr_p2p_copy2d(self.raisr._h, self._row_ptr(me.own_first), self.pitch, t.data_ptr(), self.sw,
                                              self.sw, own_rows.shape[0]))
        self.raisr.sync()
        if self.world > 1:
            self.dist.barrier(group=self.group)     # every rank's owned rows are resident before peers read them
        # halo rows: peer memory -> my window, over NVLink P2P, once
        halo_bytes = 0
        for peer, lo, hi in halo_sources(self.bands, self.rank):
            ptr, pfirst, ppitch = self.peers[peer]
            _cabi.check(self.lib.raisr_p2p_copy2d(self.raisr._h, self._row_ptr(lo), self.pitch,
                                                  ptr.value + (lo - pfirst) * ppitch, ppitch, self.sw, hi - lo + 1))
            halo_bytes += (hi - lo + 1) * self.sw
        self.halo_bytes = halo_bytes
        _cabi.check(self.lib.raisr_upsample_band_u8(self.raisr._h, self.win, self.sw, self.global_sh, self.pitch, me.src_first,
                                                    me.src_last - me.src_first + 1, self.out, self.out_pitch, me.dst_row0,
                                                    me.dst_rows, self.scale))
        self.raisr.sync()
        if self.world > 1:
            self.dist.barrier(group=self.group)     # peers may overwrite their rows only after everyone has read
        out = torch.empty((me.dst_rows, self.sw * self.scale), dtype=torch.uint8, device="cuda")
        _cabi.check(self.lib.raisr_p2p_copy2d(self.raisr._h, out.data_ptr(), self.sw * self.scale, self.out, self.out_pitch,
                                              self.sw * self.scale, me.dst_rows))
        self.raisr.sync()
        return out.cpu().numpy()

    def close(self):
        for ptr, _, _ in self.peers.values():
            self.lib.raisr_ipc_close(ptr)
        self.peers = {}
        if self.win:
            self.lib.raisr_dev_free(self.raisr._h, self.win)
            self.lib.raisr_dev_free(self.raisr._h, self.out)
            self.win = self.out = None
