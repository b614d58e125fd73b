"""Host-side mirror of the reference's interpolation wrapper.

``clUtility`` keeps the four entry points of /root/reference/basic/interpolation.py:16-107 --
``bilinear``, ``bilinear_lds``, ``bicubic``, ``bicubic_lds``, each ``(src, dst) -> [h2d, kernel, d2h] ms``
with the output size taken from ``dst.shape`` -- on top of ``raisr_resize_u8`` of the C-ABI.  The
reference uses BGRA images (``(h, w, 4)`` uint8, interpolation.py:43,61,79,97); 2-D gray arrays are
accepted too.  ``bicubic`` and ``bicubic_lds`` compute the same function in the reference (the _lds
variant only stages the source through local memory), so they map to one kernel here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from .raisr import get_elapsed_ms

_MODES = {"bilinear": 2, "bilinear_lds": 0, "bicubic": 1, "bicubic_lds": 1}


class clUtility:
    def __init__(self, device: int = 0):
        self._lib = _cabi.load()
        self._h = ctypes.c_void_p()
        _cabi.check(self._lib.raisr_create(ctypes.byref(self._h), device, 24, 3, 3, 11))

    def _run(self, src, dst, mode):
        if not isinstance(src, np.ndarray) or not isinstance(dst, np.ndarray) or src.dtype != np.uint8 or dst.dtype != np.uint8:
            raise TypeError("src and dst must be uint8 numpy arrays")
        if src.ndim != dst.ndim or src.ndim not in (2, 3) or (src.ndim == 3 and (src.shape[2] != 4 or dst.shape[2] != 4)):
            raise ValueError("expected (h, w) gray or (h, w, 4) BGRA arrays")
        if src.strides[-1] != 1 or dst.strides[-1] != 1 or (src.ndim == 3 and (src.strides[1] != 4 or dst.strides[1] != 4)):
            raise ValueError("rows must be contiguous")
        ch = 1 if src.ndim == 2 else 4
        ms = (ctypes.c_float * 3)()
        _cabi.check(self._lib.raisr_resize_u8(self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], ch,
                                              dst.ctypes.data, dst.shape[1], dst.shape[0], dst.strides[0], _MODES[mode], 1,
                                              _cabi.RAISR_HOST, ms))
        return get_elapsed_ms(ms)

    def bilinear(self, src, dst):
        """interpolation.py:37-53 / interpolation.cl:3-15 (normalised-coordinate linear sampler)."""
        return self._run(src, dst, "bilinear")

    def bilinear_lds(self, src, dst):
        """interpolation.py:73-89 / interpolation.cl:17-71 (align-corners; the mapping RAISR uses)."""
        return self._run(src, dst, "bilinear_lds")

    def bicubic(self, src, dst):
        """interpolation.py:55-71 / interpolation.cl:79-130 (Catmull-Rom a=-0.5)."""
        return self._run(src, dst, "bicubic")

    def bicubic_lds(self, src, dst):
        """interpolation.py:91-107 / interpolation.cl:132-211."""
        return self._run(src, dst, "bicubic_lds")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.raisr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def psnr(a: np.ndarray, b: np.ndarray, data_range: float = 255.0) -> float:
    """peak_signal_noise_ratio as the reference scripts use it (skimage is not installed here)."""
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(data_range * data_range / mse)
