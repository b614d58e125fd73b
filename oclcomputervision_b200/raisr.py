"""Host-side mirror of the reference's RAISR entry point.

``ClRaisr`` keeps the constructor, the public attributes and the ``upsample(src, dst, scale_factor)``
call of /root/reference/super_resolution/raisr.py:18-135, but the pyopencl context / queue / Image /
Buffer plumbing (raisr.py:62-77,96-133) is replaced by one handle of the C-ABI library
``libraisr_b200.so`` (hand-written sm_100a CUDA, see csrc/).  Differences, all additive:

* ``filters=`` may be passed instead of ``filter.p`` (the pretrained pickle needs a download,
  super_resolution/download-pre-trained-weights.txt:1); ``filters_x2`` stays settable and keeps the
  reference layout ``(24, 3, 3, 4, 121)`` float32 (raisr.py:77-78, raisr.cl:316-317);
* ``filters_x3`` / ``filters_x4`` enable the scales the reference's host gate rejects (raisr.py:90-94);
* ``upsample_batch`` and device-pointer calls for batches; ``debug_hash`` for the parity tests;
* the intended hash semantics of SURVEY.md 8(a) are computed -- the shipped OpenCL kernel returns
  after the bilinear stage (raisr.cl:219-230); ``bilinear_only`` reproduces exactly that.

Both modes of the reference are built: gray (``grayMode == 1``, raisr.py:97-100, the contract path of
the benchmark) and BGRA colour (``grayMode == 0``, raisr.py:101-104, what its ``__main__`` runs).
"""
from __future__ import annotations

import ctypes
import os
import pickle
import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi


def get_elapsed_ms(ms3) -> List[float]:
    """Same shape as the reference's helper (raisr.py:12-16): [h2d_ms, kernel_ms, d2h_ms]."""
    return [float(ms3[0]), float(ms3[1]), float(ms3[2])]


class ClRaisr:
    workGroupSize = (16, 16)  # kept for API compatibility (raisr.py:19); CUDA tiling is internal
    cscRgb2yuv = np.array([0.299, 0.587, 0.114, 0, -0.14713, -0.28886, 0.436, 0,
                           0.615, -0.51499, -0.10001, 0, 0, 0, 0, 1], dtype=np.float32)
    cscYuv2rgb = np.array([1, 0, 1.13983, 0, 1, -0.39465, -0.58060, 0,
                           1, 2.03211, 0, 0, 0, 0, 0, 1], dtype=np.float32)
    cscYuv2yuv = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1], dtype=np.float32)
    sobelX = np.array([-1, 0, 1, -2, 0, 2, -1, 0, 1], dtype=np.float32)
    sobelY = np.array([-1, -2, -1, 0, 0, 0, 1, 2, 1], dtype=np.float32)

    _TAPS = {"fp32": 0, "fp16": 1, "b24": 2, "auto": 3}

    def gaussian2d(self, shape=(3, 3), sigma=0.5):
        """Normalised 2-D Gaussian mask, MATLAB fspecial('gaussian') convention (same result as raisr.py:48-60)."""
        half_r, half_c = (float(shape[0]) - 1.0) / 2.0, (float(shape[1]) - 1.0) / 2.0
        rows = np.arange(-half_r, half_r + 1.0)[:, None]
        cols = np.arange(-half_c, half_c + 1.0)[None, :]
        mask = np.exp(-(cols * cols + rows * rows) / (2.0 * sigma * sigma))
        mask[mask < np.finfo(mask.dtype).eps * mask.max()] = 0     # fspecial's cut-off (a no-op for 9x9, sigma 2)
        total = mask.sum()
        return mask / total if total != 0 else mask

    def __init__(self, grayMode, filters: Optional[np.ndarray] = None, device: int = 0,
                 n_angle: int = 24, n_strength: int = 3, n_coherence: int = 3,
                 filter_path: Optional[str] = None, quirks: str = "intended", taps: str = "auto",
                 upscaler: str = "bilinear"):
        """quirks="as_written" reproduces the three slips of the kernel text (raisr.cl:271,310,316).
        taps = precision of the taps held in shared memory (arithmetic is fp32 in every mode): "fp32"; "fp16" =
        rounded to half precision like the reference's `(half)pf[...]` (raisr.cl:328); "b24" = sign, exponent and 15
        mantissa bits; "auto" (default) = b24 when that provably keeps every output within 5e-5 of the fp32-tap
        result (`effective_filters()` reports the bound), else fp32."""
        if grayMode not in (0, 1):
            raise ValueError("grayMode must be 1 (gray, raisr.py:97-100) or 0 (BGRA, raisr.py:101-104)")
        if quirks not in ("intended", "as_written") or taps not in self._TAPS or upscaler not in ("bilinear", "bicubic"):
            raise ValueError("quirks must be 'intended' or 'as_written', taps one of %s, upscaler 'bilinear' or 'bicubic'" % sorted(self._TAPS))
        self.grayMode = grayMode
        self.n_angle, self.n_strength, self.n_coherence = n_angle, n_strength, n_coherence
        self._lib = _cabi.load()
        self._h = ctypes.c_void_p()
        _cabi.check(self._lib.raisr_create(ctypes.byref(self._h), device, n_angle, n_strength, n_coherence, 11))
        self._filters = {}
        self.quirks, self.taps = quirks, taps
        if quirks == "as_written":
            self.set_option("quirks", 1)
        if taps != "auto":
            self.set_option("taps", self._TAPS[taps])
        if upscaler == "bicubic":   # the reference's unused cubic_sample as stage 1 (raisr.cl:63-106)
            self.set_option("cheap_upscaler", 1)
        # raisr.py:80-82
        g = self.gaussian2d([9, 9], 2)
        g = np.diag(g.ravel()).astype(np.float32)
        self.gaussian = np.diag(g).copy()
        if filters is None:
            # raisr.py:74-78 unpickles `filter.p` from the module's own directory; so does this class (never the
            # process's working directory).  A pickle executes code when loaded: only point filter_path at files
            # you trust -- pass `filters=` (a plain array) otherwise.
            path = filter_path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "filter.p")
            if os.path.exists(path):
                with open(path, "rb") as fp:
                    filters = pickle.load(fp)
            elif filter_path is not None:
                self.close()
                raise FileNotFoundError(filter_path)
        if filters is not None:
            self.filters_x2 = filters

    # ---- filter tables (reference attribute: filters_x2, raisr.py:78)
    def _set_filters(self, scale: int, table) -> None:
        t = np.ascontiguousarray(np.asarray(table).astype(np.float32))
        want = (self.n_angle, self.n_strength, self.n_coherence, scale * scale, 121)
        if t.shape != want:
            raise ValueError("filters for scale %d must have shape %s, got %s" % (scale, want, t.shape))
        _cabi.check(self._lib.raisr_set_filters(self._h, scale, t.ctypes.data, t.size))
        self._filters[scale] = t

    filters_x2 = property(lambda self: self._filters.get(2), lambda self, t: self._set_filters(2, t))
    filters_x3 = property(lambda self: self._filters.get(3), lambda self, t: self._set_filters(3, t))
    filters_x4 = property(lambda self: self._filters.get(4), lambda self, t: self._set_filters(4, t))

    def effective_filters(self, scale: int):
        """(table, tap_format, b24_bound): the taps the gray filter kernel actually multiplies by (reference layout),
        the format in use ("fp32" | "fp16" | "b24") and the bound on |output - fp32-tap output| of the b24 records."""
        t = np.empty_like(self._filters[scale])
        fmt, bound = ctypes.c_int(), ctypes.c_float()
        _cabi.check(self._lib.raisr_get_effective_filters(self._h, scale, t.ctypes.data, t.size, ctypes.byref(fmt), ctypes.byref(bound)))
        return t, ("fp32", "fp16", "b24")[fmt.value], float(bound.value)

    def set_quantizers(self, strength_q: Sequence[float], coherence_q: Sequence[float]) -> None:
        """Thresholds of raisr.py:112-115."""
        sq = np.ascontiguousarray(strength_q, dtype=np.float32)
        cq = np.ascontiguousarray(coherence_q, dtype=np.float32)
        _cabi.check(self._lib.raisr_set_quantizers(self._h, sq.ctypes.data, sq.size, cq.ctypes.data, cq.size))

    def set_option(self, key: str, value: int) -> None:
        _cabi.check(self._lib.raisr_set_option(self._h, key.encode(), int(value)))

    # ---- the reference call (raisr.py:85-135)
    def upsample(self, src, dst, scale_factor):
        if scale_factor not in self._filters:
            # reference behaviour for an untrained scale (raisr.py:93-94)
            print('Fatal. not trained for scale factor {}'.format(scale_factor))
            return
        ms = (ctypes.c_float * 3)()
        if self.grayMode == 0:   # BGRA images (raisr.py:101-104,163-164)
            src, dst = self._check_pair(src, dst, np.uint8, channels=4)
            _cabi.check(self._lib.raisr_upsample_bgra_u8(
                self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                dst.ctypes.data, dst.shape[1], dst.shape[0], dst.strides[0],
                int(scale_factor), 1, _cabi.RAISR_HOST, ms))
            return get_elapsed_ms(ms)
        src, dst = self._check_pair(src, dst, np.uint8)
        _cabi.check(self._lib.raisr_upsample_u8(
            self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
            dst.ctypes.data, dst.shape[1], dst.shape[0], dst.strides[0],
            int(scale_factor), 1, _cabi.RAISR_HOST, ms))
        return get_elapsed_ms(ms)

    upscale = upsample  # name used by BASELINE.json's north_star

    def upsample_batch(self, src, dst, scale_factor):
        """Host arrays, frames pipelined (H2D / kernels / D2H on three streams).
        gray:   src (N, sh, sw) u8        -> dst (N, s*sh, s*sw) u8 or float32
        colour: src (N, sh, sw, 4) u8 BGRA -> dst (N, s*sh, s*sw, 4) u8 or float32"""
        if scale_factor not in self._filters:
            print('Fatal. not trained for scale factor {}'.format(scale_factor))
            return
        nd = 4 if self.grayMode == 0 else 3
        if src.ndim != nd or dst.ndim != nd or src.shape[0] != dst.shape[0] or (nd == 4 and (src.shape[3] != 4 or dst.shape[3] != 4)):
            raise ValueError("upsample_batch expects (N, h, w) arrays in gray mode and (N, h, w, 4) BGRA arrays in colour mode")
        if src.dtype != np.uint8 or not src.flags.c_contiguous or not dst.flags.c_contiguous:
            raise ValueError("src must be C-contiguous uint8 and dst C-contiguous")
        if self.grayMode == 0:
            fn = {np.dtype(np.uint8): self._lib.raisr_upsample_bgra_u8, np.dtype(np.float32): self._lib.raisr_upsample_bgra_f32}.get(dst.dtype)
        else:
            fn = {np.dtype(np.uint8): self._lib.raisr_upsample_u8, np.dtype(np.float32): self._lib.raisr_upsample_f32}.get(dst.dtype)
        if fn is None:
            raise ValueError("dst must be uint8 or float32")
        ms = (ctypes.c_float * 3)()
        n, sh, sw = src.shape[:3]
        _cabi.check(fn(self._h, src.ctypes.data, sw, sh, src.strides[1], dst.ctypes.data, dst.shape[2], dst.shape[1],
                       dst.strides[1], int(scale_factor), n, _cabi.RAISR_HOST, ms))
        return get_elapsed_ms(ms)

    def upsample_f32(self, src, scale_factor) -> np.ndarray:
        """Float [0,1] output of one frame (the value the reference's write_imagef would quantise)."""
        src = np.ascontiguousarray(src, dtype=np.uint8)
        if scale_factor not in self._filters:
            raise ValueError("not trained for scale factor %d" % scale_factor)
        if self.grayMode == 0:
            dst = np.empty((src.shape[0] * scale_factor, src.shape[1] * scale_factor, 4), np.float32)
            _cabi.check(self._lib.raisr_upsample_bgra_f32(self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                                                          dst.ctypes.data, dst.shape[1], dst.shape[0], dst.strides[0],
                                                          int(scale_factor), 1, _cabi.RAISR_HOST, None))
            return dst
        dst = np.empty((src.shape[0] * scale_factor, src.shape[1] * scale_factor), np.float32)
        _cabi.check(self._lib.raisr_upsample_f32(self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                                                 dst.ctypes.data, dst.shape[1], dst.shape[0], dst.strides[0],
                                                 int(scale_factor), 1, _cabi.RAISR_HOST, None))
        return dst

    def upsample_device(self, src_ptr: int, sw: int, sh: int, src_pitch: int, dst_ptr: int, dst_pitch: int,
                        scale_factor: int, n_frames: int = 1, out_dtype=np.uint8, timed: bool = False):
        """Device-pointer form (e.g. torch tensors' data_ptr()); enqueues on the handle's stream."""
        fn = self._lib.raisr_upsample_u8 if np.dtype(out_dtype) == np.uint8 else self._lib.raisr_upsample_f32
        ms = (ctypes.c_float * 3)() if timed else None
        _cabi.check(fn(self._h, src_ptr, sw, sh, src_pitch, dst_ptr, sw * scale_factor, sh * scale_factor, dst_pitch,
                       int(scale_factor), n_frames, _cabi.RAISR_DEVICE, ms))
        return get_elapsed_ms(ms) if timed else None

    def bilinear_only(self, src, dst, scale_factor):
        """Literal behaviour of the shipped kernel (raisr.cl:219-230) / bilinear_lds (interpolation.cl:17-71)."""
        src, dst = self._check_pair(src, dst, np.uint8)
        ms = (ctypes.c_float * 3)()
        _cabi.check(self._lib.raisr_bilinear_u8(
            self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
            dst.ctypes.data, dst.shape[1], dst.shape[0], dst.strides[0], int(scale_factor), 1, _cabi.RAISR_HOST, ms))
        return get_elapsed_ms(ms)

    def debug_hash(self, src, scale_factor) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(hash int32, angle, L1, coherence, U) per output pixel (raisr.cl:278-317)."""
        src = np.ascontiguousarray(src, dtype=np.uint8)
        dh, dw = src.shape[0] * scale_factor, src.shape[1] * scale_factor
        h = np.empty((dh, dw), np.int32)
        a, l1, co, u = (np.empty((dh, dw), np.float32) for _ in range(4))
        if self.grayMode == 0:   # colour: the quantities come from the Y plane; there is no single upscaled image
            if src.ndim != 3 or src.shape[2] != 4:
                raise ValueError("colour mode expects (h, w, 4) BGRA arrays")
            _cabi.check(self._lib.raisr_debug_hash_bgra(self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                                                        int(scale_factor), h.ctypes.data, a.ctypes.data, l1.ctypes.data,
                                                        co.ctypes.data, _cabi.RAISR_HOST))
            return h, a, l1, co, None
        _cabi.check(self._lib.raisr_debug_hash(self._h, src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                                               int(scale_factor), h.ctypes.data, a.ctypes.data, l1.ctypes.data,
                                               co.ctypes.data, u.ctypes.data, _cabi.RAISR_HOST))
        return h, a, l1, co, u

    def set_stream(self, cuda_stream: int) -> None:
        _cabi.check(self._lib.raisr_set_stream(self._h, cuda_stream))

    def sync(self) -> None:
        _cabi.check(self._lib.raisr_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.raisr_launch_count(self._h))

    def last_kernel_ms(self) -> Tuple[float, float]:
        a, b = ctypes.c_float(), ctypes.c_float()
        _cabi.check(self._lib.raisr_last_kernel_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def device_info(self):
        sm, khz = ctypes.c_int(), ctypes.c_int()
        name = ctypes.create_string_buffer(128)
        _cabi.check(self._lib.raisr_device_info(self._h, ctypes.byref(sm), ctypes.byref(khz), name, 128))
        return dict(sm_count=sm.value, sm_clock_khz=khz.value, name=name.value.decode())

    def measure_ffma_tflops(self) -> float:
        t = ctypes.c_float()
        _cabi.check(self._lib.raisr_measure_ffma_tflops(self._h, ctypes.byref(t)))
        return t.value

    @staticmethod
    def _check_pair(src, dst, dtype, channels=1):
        if not isinstance(src, np.ndarray) or not isinstance(dst, np.ndarray):
            raise TypeError("src and dst must be numpy arrays")
        if channels == 1:
            if src.ndim != 2 or dst.ndim != 2:
                raise ValueError("gray mode expects 2-D arrays (raisr.py:98)")
        elif src.ndim != 3 or dst.ndim != 3 or src.shape[2] != channels or dst.shape[2] != channels:
            raise ValueError("colour mode expects (h, w, 4) BGRA arrays (raisr.py:102,163-164)")
        if src.dtype != dtype or dst.dtype != dtype:
            raise ValueError("src and dst must be %s" % np.dtype(dtype).name)
        if src.strides[-1] != 1 or dst.strides[-1] != 1 or (channels > 1 and (src.strides[1] != channels or dst.strides[1] != channels)):
            raise ValueError("rows must be contiguous")
        return src, dst

    # ---- page-locking of caller-owned arrays (no reference counterpart; see raisr_host_register)
    _pinned = {}

    @classmethod
    def pin(cls, arr: np.ndarray) -> np.ndarray:
        """Page-lock the memory of a C-contiguous numpy array that will be passed to `upsample` repeatedly (the
        reference's own loop, raisr.py:166-182, re-uses one src and one dst): the HOST path then copies at full
        PCIe speed instead of through the driver's pageable staging.  The lock is released when the array is
        garbage-collected or `unpin` is called.  Returns `arr`."""
        if not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous or arr.nbytes == 0:
            raise ValueError("pin() needs a non-empty C-contiguous numpy array")
        key = arr.ctypes.data
        if key in cls._pinned:
            return arr
        lib = _cabi.load()
        _cabi.check(lib.raisr_host_register(key, arr.nbytes))
        owner = arr if arr.base is None else arr.base
        fin = None
        try:
            fin = weakref.finalize(owner, cls._unregister, key)
        except TypeError:       # the owner of the memory is not weak-referenceable: the caller must unpin
            pass
        cls._pinned[key] = fin
        return arr

    @classmethod
    def _unregister(cls, key: int) -> None:
        if cls._pinned.pop(key, "absent") != "absent":
            try:
                _cabi.load().raisr_host_unregister(key)
            except Exception:
                pass

    @classmethod
    def unpin(cls, arr: np.ndarray) -> None:
        key = arr.ctypes.data
        fin = cls._pinned.get(key)
        if fin is not None:
            fin.detach()
        cls._unregister(key)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.raisr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
