"""Histogram equalisation on the B200 -- host-side mirror of the reference's `histeq` package
(SURVEY.md 8(f) row N4).

Same names and argument meaning as the reference:

    clHistEq.getInstance()                          /root/reference/histeq/eq_opencl.py:8-35
        .histGrid(gray)            -> (hist, ms)    eq_opencl.py:37-51   uint32 (h//32, w//256, 256)
        .histeqGlobal(gray, map)   -> (img, ms)     eq_opencl.py:53-68
        .histeqLocalBlock(gray, maps, blockshape)   eq_opencl.py:70-89
    calc_transfer_func(hist, alpha, punch, clip)    eq_global.py:10-39   (host, float64, 256 entries)
    histeq_global(gray, alpha, punch, clip)         eq_global.py:41-65
    histeq_local_block(gray, alpha, punch, clip, blockshape)   eq_local_block.py:10-79

The three image passes run as CUDA kernels (csrc/histeq.cuh) behind the C-ABI `ocv_*` entry points; the
256-entry transfer functions stay on the host like the reference's (a few microseconds of numpy on
kilobytes).  The reference's `use_gpu=False` numpy branch is not reproduced: this package has no CPU
path, asking for one raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _cabi

HIST_BINS = 256          # eq_opencl.py:13
HIST_THREAD_NUM = 32     # eq_opencl.py:14: rows per histogram tile


def _as_gray(gray) -> np.ndarray:
    g = np.asarray(gray)
    if g.ndim != 2 or g.dtype != np.uint8:
        raise ValueError("expected an 8-bit single-channel image (h, w) uint8")
    return np.ascontiguousarray(g)


class clHistEq:
    """Device context for the three histeq passes (singleton like the reference's, eq_opencl.py:27-35)."""

    _instance: Optional["clHistEq"] = None

    @staticmethod
    def getInstance(device: int = 0) -> "clHistEq":
        if clHistEq._instance is None:
            clHistEq._instance = clHistEq(device)
        return clHistEq._instance

    def __init__(self, device: int = 0):
        self._lib = _cabi.load()
        h = ctypes.c_void_p()
        _cabi.check(self._lib.raisr_create(ctypes.byref(h), int(device), 24, 3, 3, 11))
        self._h = h
        self.HIST_BINS = HIST_BINS
        self.HIST_THREAD_NUM = HIST_THREAD_NUM

    def close(self):
        if getattr(self, "_h", None):
            self._lib.raisr_destroy(self._h)
            self._h = None
            if clHistEq._instance is self:
                clHistEq._instance = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def histGrid(self, npInGray) -> Tuple[np.ndarray, float]:
        g = _as_gray(npInGray)
        h, w = g.shape
        out = np.empty((h // HIST_THREAD_NUM, w // HIST_BINS, HIST_BINS), np.uint32)
        ms = (ctypes.c_float * 3)()
        _cabi.check(self._lib.ocv_hist_grid_u8(self._h, g.ctypes.data, w, h, g.strides[0], out.ctypes.data,
                                               _cabi.RAISR_HOST, ms))
        return out, float(ms[1])

    def histeqGlobal(self, npInGray, npMapping) -> Tuple[np.ndarray, float]:
        g = _as_gray(npInGray)
        m = np.ascontiguousarray(npMapping)
        if m.dtype != np.uint8 or m.size != HIST_BINS:
            raise ValueError("mapping must be 256 uint8 entries")
        h, w = g.shape
        out = np.empty_like(g)
        ms = (ctypes.c_float * 3)()
        _cabi.check(self._lib.ocv_histeq_global_u8(self._h, g.ctypes.data, w, h, g.strides[0], out.ctypes.data,
                                                   out.strides[0], m.ctypes.data, _cabi.RAISR_HOST, ms))
        return out, float(ms[1])

    def histeqLocalBlock(self, npInGray, npMappings, blockshape) -> Tuple[np.ndarray, float]:
        g = _as_gray(npInGray)
        m = np.ascontiguousarray(npMappings, dtype=np.float32)
        if m.ndim != 3 or m.shape[2] != HIST_BINS:
            raise ValueError("mappings must be float32 (ny, nx, 256)")
        h, w = g.shape
        bh, bw = int(blockshape[0]), int(blockshape[1])
        if m.shape[0] != max(h // bh, 0) or m.shape[1] != max(w // bw, 0) or m.shape[0] < 1 or m.shape[1] < 1:
            raise ValueError("mappings grid %s does not match image %s / block %s" % (m.shape[:2], g.shape, (bh, bw)))
        out = np.empty_like(g)
        ms = (ctypes.c_float * 3)()
        _cabi.check(self._lib.ocv_histeq_local_block_u8(self._h, g.ctypes.data, w, h, g.strides[0], out.ctypes.data,
                                                        out.strides[0], m.ctypes.data, m.shape[1], m.shape[0], bw, bh,
                                                        _cabi.RAISR_HOST, ms))
        return out, float(ms[1])


def calc_transfer_func(hist, alpha, punch, clip) -> np.ndarray:
    """Grey-level mapping from a histogram (eq_global.py:10-39), float64[len(hist)].

    The CDF is rebuilt over the `punch`..`1-punch` mass only (levels below map to 0, above to full scale),
    blended with the identity by `alpha`, clipped to 0..255 and gain-limited to [level/clip, level*clip].
    """
    hist = np.asarray(hist)
    total = np.sum(hist)
    if hist.ndim != 1 or total == 0:
        raise ValueError("histogram must be 1-D and non-empty")
    level = np.arange(hist.size)
    cdf = np.cumsum(hist) / total
    lo = int(np.flatnonzero(cdf >= punch)[0])
    hi = int(np.flatnonzero(cdf >= 1 - punch)[0])
    body = hist[lo:hi]
    cdf[:lo] = 0
    cdf[hi:] = 1
    cdf[lo:hi] = np.cumsum(body) / np.sum(body)
    curve = np.clip(alpha * cdf * 255 + (1 - alpha) * level, 0, 255)
    return np.clip(curve, level / clip, level * clip)


def _no_cpu_path(use_gpu):
    if not use_gpu:
        raise NotImplementedError("oclcomputervision_b200 has no CPU path (the reference's use_gpu=False numpy branch "
                                  "is not reproduced)")


def histeq_global(gray, alpha=1, punch=0.05, clip=2, use_gpu=True) -> np.ndarray:
    """eq_global.py:41-65: tile histograms on the device, one transfer function, LUT pass on the device."""
    _no_cpu_path(use_gpu)
    cleq = clHistEq.getInstance()
    grid, _ = cleq.histGrid(gray)
    hist = grid.sum(axis=0).sum(axis=0)
    mapping = calc_transfer_func(hist, alpha, punch, clip).astype(np.uint8)
    out, _ = cleq.histeqGlobal(gray, mapping)
    return out


def histeq_local_block(gray, alpha=0.5, punch=0.05, clip=3, blockshape=(256, 256), use_gpu=True) -> np.ndarray:
    """eq_local_block.py:10-79: one transfer function per block, blended bilinearly between block centres.

    The reference's OpenCL branch hard-codes 8 tile rows and one tile column per block (eq_local_block.py:
    22-27), i.e. 256x256 blocks; here any block whose height is a multiple of 32 and width a multiple of 256
    is merged from the tile grid (same result as histogramming the block directly).
    """
    _no_cpu_path(use_gpu)
    g = _as_gray(gray)
    bh, bw = int(blockshape[0]), int(blockshape[1])
    if bh % HIST_THREAD_NUM or bw % HIST_BINS or bh < 1 or bw < 1:
        raise ValueError("block shape must be a multiple of the 32x256 histogram tile")
    cleq = clHistEq.getInstance()
    grid, _ = cleq.histGrid(g)
    ny, nx = g.shape[0] // bh, g.shape[1] // bw
    if ny < 1 or nx < 1:
        raise ValueError("image smaller than one block")
    ry, rx = bh // HIST_THREAD_NUM, bw // HIST_BINS
    merged = grid[:ny * ry, :nx * rx].reshape(ny, ry, nx, rx, HIST_BINS).sum(axis=(1, 3))
    mappings = np.zeros((ny, nx, HIST_BINS), np.float32)
    for i in range(ny):
        for j in range(nx):
            mappings[i, j] = calc_transfer_func(merged[i, j], alpha, punch, clip).astype(np.float32)
    out, _ = cleq.histeqLocalBlock(g, mappings, (bh, bw))
    return out
