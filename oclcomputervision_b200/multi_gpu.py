"""Multi-GPU use of the RAISR path: one process per GPU, no data-path collective.

* batches of frames: every rank upsamples a contiguous slice of the batch (``shard_frames``);
* one very large image: every rank owns a band of source rows, produces the matching band of
  output rows and reads the <=3 halo source rows per side it needs from its neighbours' memory over
  NVLink peer-to-peer (CUDA IPC handles exchanged once; the copy is a peer ``cudaMemcpy2DAsync`` issued
  by the consumer -- "read once"; ranks are ordered by sequence words in device memory, not by barriers).

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
    """Row-banded upscale of one large image across the ranks of a job, on the C-ABI only.

    Every rank owns a band of source rows in a plain ``cudaMalloc`` window that its neighbours map through CUDA IPC
    (handles exchanged ONCE at construction -- through ``exchange``, by default ``torch.distributed.all_gather_object``;
    nothing of torch touches the data path).  One call then enqueues, all on the handle's stream and without any
    host synchronisation or collective:

        H2D of the owned rows (optional: they may already be resident)          raisr_copy2d
        publish sequence number k in this rank's ``ready`` word                 raisr_flag_set
        wait until both neighbours have published k                             raisr_flag_wait (polls peer memory)
        read the <=3 halo rows per side from the neighbours' windows over NVLink  raisr_p2p_copy2d  ("read once")
        publish k in this rank's ``done`` word (my reads of the peers are over)  raisr_flag_set
        upscale + hash + filter of the band                                     raisr_upsample_band_u8
        D2H of the band (optional)                                              raisr_copy2d

    Before call k+1 overwrites the owned rows it waits for the neighbours' ``done`` >= k.  Device time is taken with
    CUDA events on the same stream (``last_ms``).  Every rank must make the same sequence of calls, and ``close()`` may
    only be called once all ranks have finished their last call (it frees memory the neighbours have mapped).  The coordinate map uses the GLOBAL image size (raisr.cl:209).
    """

    FLAG_BYTES = 256          # [0] ready, [1] done; the rows start behind them
    TIMEOUT_MS = 20000

    def __init__(self, raisr, sw: int, global_sh: int, scale: int, group=None, exchange=None, rank=None, world=None):
        self.raisr, self.sw, self.global_sh, self.scale = raisr, sw, global_sh, scale
        if exchange is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)

                def exchange(obj):
                    out = [None] * world
                    dist.all_gather_object(out, obj, group=group)
                    return out
            else:
                rank, world = 0, 1
                exchange = lambda obj: [obj]   # noqa: E731
        self.rank, self.world = rank, world
        self.bands = band_plan(global_sh, scale, self.world)
        self.me = self.bands[self.rank]
        self.lib = _cabi.load()
        self.pitch = (sw + 255) // 256 * 256
        self.n_win = max(1, self.me.src_last - self.me.src_first + 1)
        self.base = ctypes.c_void_p()
        _cabi.check(self.lib.raisr_dev_alloc(raisr._h, ctypes.byref(self.base), self.FLAG_BYTES + self.n_win * self.pitch))
        zero = (ctypes.c_ubyte * self.FLAG_BYTES)()
        _cabi.check(self.lib.raisr_copy2d(raisr._h, self.base, self.FLAG_BYTES, zero, self.FLAG_BYTES, self.FLAG_BYTES, 1, 0))
        raisr.sync()
        self.win = self.base.value + self.FLAG_BYTES
        self.out = ctypes.c_void_p()
        self.out_pitch = (sw * scale + 255) // 256 * 256
        _cabi.check(self.lib.raisr_dev_alloc(raisr._h, ctypes.byref(self.out), max(1, self.me.dst_rows) * self.out_pitch))
        handle = (ctypes.c_ubyte * 64)()
        _cabi.check(self.lib.raisr_ipc_export(self.base, handle))
        infos = exchange(dict(handle=bytes(handle), first=self.me.src_first, pitch=self.pitch))
        self.halo = halo_sources(self.bands, self.rank)
        # the neighbours that read MY rows (their halo lists name me): they must be done before I overwrite my rows
        self.readers = sorted({b.rank for b in self.bands if b.rank != self.rank and any(p == self.rank for p, _, _ in halo_sources(self.bands, b.rank))})
        self.peers = {}
        for peer in sorted({p for p, _, _ in self.halo} | set(self.readers)):
            ptr = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(infos[peer]["handle"])
            _cabi.check(self.lib.raisr_ipc_open(buf, ctypes.byref(ptr)))
            self.peers[peer] = (ptr, infos[peer]["first"], infos[peer]["pitch"])
        self.halo_bytes = sum((hi - lo + 1) * sw for _, lo, hi in self.halo)
        self.seq = 0
        self.last_ms = None

    def _row_ptr(self, global_row: int) -> int:
        return self.win + (global_row - self.me.src_first) * self.pitch

    def enqueue(self, own_rows_host_ptr: int = 0, own_rows_pitch: int = 0, out_host_ptr: int = 0, out_host_pitch: int = 0) -> None:
        """Enqueue one banded upscale on the handle's stream (see the class docstring).  own_rows_host_ptr = 0: the
        owned rows are already in the window (``own_rows_device_ptr``); out_host_ptr = 0: the band stays on the device
        (``out`` / ``out_pitch``).  Returns without waiting; ``finish()`` blocks and returns the device time."""
        lib, h, me = self.lib, self.raisr._h, self.me
        self.seq += 1
        k = self.seq
        _cabi.check(lib.raisr_timer_mark(h, 0))
        if own_rows_host_ptr:
            for peer in self.readers:      # write-after-read: the neighbours have finished reading my rows of call k-1
                _cabi.check(lib.raisr_flag_wait(h, self.peers[peer][0].value + 4, k - 1, self.TIMEOUT_MS))
            if me.dst_rows:
                _cabi.check(lib.raisr_copy2d(h, self._row_ptr(me.own_first), self.pitch, own_rows_host_ptr, own_rows_pitch or self.sw,
                                             self.sw, me.own_last - me.own_first + 1, 0))
        _cabi.check(lib.raisr_timer_mark(h, 1))
        _cabi.check(lib.raisr_flag_set(h, self.base.value, k))                          # my rows of call k are resident
        for peer in sorted({p for p, _, _ in self.halo}):
            _cabi.check(lib.raisr_flag_wait(h, self.peers[peer][0].value, k, self.TIMEOUT_MS))
        for peer, lo, hi in self.halo:                                                   # NVLink P2P, each halo row read once
            ptr, pfirst, ppitch = self.peers[peer]
            _cabi.check(lib.raisr_p2p_copy2d(h, self._row_ptr(lo), self.pitch, ptr.value + self.FLAG_BYTES + (lo - pfirst) * ppitch,
                                             ppitch, self.sw, hi - lo + 1))
        _cabi.check(lib.raisr_flag_set(h, self.base.value + 4, k))                      # I no longer read the peers' rows of call k
        _cabi.check(lib.raisr_timer_mark(h, 2))
        if me.dst_rows:
            _cabi.check(lib.raisr_upsample_band_u8(h, self.win, self.sw, self.global_sh, self.pitch, me.src_first,
                                                   me.src_last - me.src_first + 1, self.out, self.out_pitch, me.dst_row0,
                                                   me.dst_rows, self.scale))
        _cabi.check(lib.raisr_timer_mark(h, 3))
        if out_host_ptr and me.dst_rows:
            _cabi.check(lib.raisr_copy2d(h, out_host_ptr, out_host_pitch or self.sw * self.scale, self.out, self.out_pitch,
                                         self.sw * self.scale, me.dst_rows, 1))
        _cabi.check(lib.raisr_timer_mark(h, 4))

    def finish(self):
        """Block until the enqueued call is complete; returns dict(h2d, halo, kernels, d2h, total) in device ms."""
        ms = ctypes.c_float()
        out = {}
        for name, a, b in (("h2d", 0, 1), ("halo", 1, 2), ("kernels", 2, 3), ("d2h", 3, 4), ("total", 0, 4)):
            _cabi.check(self.lib.raisr_timer_elapsed_ms(self.raisr._h, a, b, ctypes.byref(ms)))
            out[name] = float(ms.value)
        self.raisr.sync()
        self.last_ms = out
        return out

    @property
    def own_rows_device_ptr(self) -> int:
        return self._row_ptr(self.me.own_first)

    def upsample_band(self, own_rows):
        """Host arrays in and out: this rank's owned source rows (2-D uint8) -> its band of output rows."""
        import numpy as np
        me = self.me
        assert own_rows.shape == (me.own_last - me.own_first + 1, self.sw) and own_rows.dtype == np.uint8
        own_rows = np.ascontiguousarray(own_rows)
        out = np.empty((me.dst_rows, self.sw * self.scale), np.uint8)
        self.enqueue(own_rows.ctypes.data if own_rows.size else 0, self.sw, out.ctypes.data if out.size else 0, self.sw * self.scale)
        self.finish()
        return out

    def close(self):
        if getattr(self, "base", None):
            # The neighbours read their halo rows out of MY window and I poll words in THEIRS: the caller must make sure
            # that every rank has finished its last call before any rank closes (a barrier or any collective; bench.py's
            # all_reduce of the timings, the tests' barrier).
            try:
                self.raisr.sync()
            except Exception:
                pass
            for ptr, _, _ in self.peers.values():
                self.lib.raisr_ipc_close(ptr)
            self.peers = {}
            self.lib.raisr_dev_free(self.raisr._h, self.base)
            self.lib.raisr_dev_free(self.raisr._h, self.out)
            self.base = self.out = None
