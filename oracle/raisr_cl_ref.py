"""ctypes front end of oracle/_ref/lib*_ref_*.so -- the reference's own raisr.cl and interpolation.cl run on the CPU.
TEST INFRASTRUCTURE.

`run()` does what ClRaisr.upsample does around the kernel (raisr.py:86-133): CL_R or CL_BGRA UNORM_INT8 images, the
Sobel / colour / Gaussian / quantiser buffers of raisr.py:19-50,82-84,111-114, one work-item per destination pixel in
16 x 16 groups.  `interp()` does what clUtility.bilinear / bilinear_lds / bicubic / bicubic_lds do
(basic/interpolation.py:37-117).  The libraries only exist where /root/reference does (oracle/build_ref.py); everywhere else the
committed outputs in tests/golden/ref_cl*.npz stand in for them (oracle/make_golden_ref_cl.py).
"""
import ctypes
import os

import numpy as np

from . import build_ref

F32 = np.float32
# raisr.py:19-50 (host-side constants handed to the kernel)
CSC_RGB2YUV = np.array([0.299, 0.587, 0.114, 0, -0.14713, -0.28886, 0.436, 0, 0.615, -0.51499, -0.10001, 0, 0, 0, 0, 1], F32)
CSC_YUV2RGB = np.array([1, 0, 1.13983, 0, 1, -0.39465, -0.58060, 0, 1, 2.03211, 0, 0, 0, 0, 0, 1], F32)
CSC_IDENTITY = np.eye(4, dtype=F32).ravel()
SOBEL_X = np.array([-1, 0, 1, -2, 0, 2, -1, 0, 1], F32)
SOBEL_Y = np.array([-1, -2, -1, 0, 0, 0, 1, 2, 1], F32)
STRENGTH_Q = np.array([0.0001, 0.001], F32)      # raisr.py:111
COHERENCE_Q = np.array([0.25, 0.5], F32)         # raisr.py:113

_libs = {}


def available() -> bool:
    return os.path.exists(build_ref.REF_CL) or all(os.path.exists(p) for p in build_ref.lib_paths())


def _open(name: str):
    if name not in _libs:
        if os.path.exists(build_ref.REF_CL):
            build_ref.build()
        _libs[name] = ctypes.CDLL(os.path.join(build_ref.OUT, name))
    return _libs[name]


def _lib(kind: str, prec: str):
    lib = _open("libraisr_ref_%s_%s.so" % (kind, prec))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.raisr_cl_run.restype = ci
    lib.raisr_cl_run.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, ci, vp]
    return lib


def gaussian81() -> np.ndarray:
    """raisr.py:47-59,82-84: fspecial('gaussian', [9, 9], 2), flattened."""
    y, x = np.ogrid[-4.0:5.0, -4.0:5.0]
    h = np.exp(-(x * x + y * y) / (2.0 * 2 * 2))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    h /= h.sum()
    return np.ascontiguousarray(h.ravel(), dtype=F32)


def run(src: np.ndarray, filters: np.ndarray, scale: int = 2, *, kind: str = "shipped", prec: str = "f16") -> np.ndarray:
    """src: (h, w) u8 gray or (h, w, 4) u8 BGRA; returns the destination image the kernel writes.
    kind: "shipped" (early return active, as the file lies) | "full" (early return compiled out) | "intended" (full, with
    the three slips of raisr.cl:271,310,316 corrected in the text: build_ref.INTENDED_FIXES) | "cubic_intended" (the same with
    stage 1 calling the file's cubic_sample instead of linear_sample: build_ref.CUBIC_SWITCH; prec "f32" only).
    prec: "f16" (half is binary16) | "f32" (half kept in binary32)."""
    assert kind in ("shipped", "full", "intended", "cubic_intended") and prec in ("f16", "f32")
    src = np.ascontiguousarray(src, dtype=np.uint8)
    gray = src.ndim == 2
    sh, sw = src.shape[:2]
    dh, dw = sh * scale, sw * scale
    dst = np.zeros((dh, dw) if gray else (dh, dw, 4), np.uint8)
    flt = np.ascontiguousarray(filters, dtype=F32)
    to_yuv, from_yuv = (CSC_IDENTITY, CSC_IDENTITY) if gray else (CSC_RGB2YUV, CSC_YUV2RGB)     # raisr.py:99-105
    g = gaussian81()
    rc = _lib(kind, prec).raisr_cl_run(src.ctypes.data, sw, sh, src.strides[0], 1 if gray else 4, dst.ctypes.data, dw, dh, dst.strides[0],
                                       SOBEL_X.ctypes.data, SOBEL_Y.ctypes.data, to_yuv.ctypes.data, from_yuv.ctypes.data, g.ctypes.data,
                                       STRENGTH_Q.ctypes.data, COHERENCE_Q.ctypes.data, int(scale), flt.ctypes.data)
    if rc != 0:
        raise ValueError("raisr_cl_run: destination must be a multiple of the 16 x 16 work-group (raisr.py:129)")
    return dst


INTERP_KERNELS = {"bilinear": 0, "bilinear_lds": 1, "bicubic": 2, "bicubic_lds": 3}     # clUtility method -> kernel (interpolation.py:30-33)


def interp(src: np.ndarray, out_hw, method: str, *, prec: str = "f16") -> np.ndarray:
    """clUtility.<method>(src, dst) (basic/interpolation.py:37-117) with the reference's own kernels.
    src: (h, w, 4) u8 BGRA as the reference uses, or (h, w) u8 (a CL_R image; the kernels do not care)."""
    assert prec in ("f16", "f32")
    src = np.ascontiguousarray(src, dtype=np.uint8)
    gray = src.ndim == 2
    dh, dw = out_hw
    dst = np.zeros((dh, dw) if gray else (dh, dw, 4), np.uint8)
    lib = _open("libinterp_ref_%s.so" % prec)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.interp_cl_run.restype = ci
    lib.interp_cl_run.argtypes = [ci, vp, ci, ci, ci, ci, vp, ci, ci, ci]
    rc = lib.interp_cl_run(INTERP_KERNELS[method], src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], 1 if gray else 4,
                           dst.ctypes.data, dw, dh, dst.strides[0])
    if rc != 0:
        raise ValueError("interp_cl_run: the LDS kernels need a destination that is a multiple of the 16 x 16 work-group")
    return dst


# ---------------------------------------------------------------- histeq/hist.cl as clHistEq launches it (eq_opencl.py:37-89)
def _hist_lib():
    lib = _open("libhist_ref.so")
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.hist_cl_run.restype = ci
    lib.hist_cl_run.argtypes = [vp, ci, ci, ci, vp]
    lib.histeq_global_cl_run.restype = ci
    lib.histeq_global_cl_run.argtypes = [vp, ci, ci, ci, vp, ci, vp]
    lib.histeq_local_block_cl_run.restype = ci
    lib.histeq_local_block_cl_run.argtypes = [vp, ci, ci, ci, vp, ci, vp, ci, ci, ci, ci]
    return lib


def hist_grid(gray: np.ndarray) -> np.ndarray:
    """clHistEq.histGrid (eq_opencl.py:37-51): (h // 32, w // 256, 256) uint32."""
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    h, w = gray.shape
    out = np.zeros((h // 32, w // 256, 256), np.uint32)
    if _hist_lib().hist_cl_run(gray.ctypes.data, w, h, gray.strides[0], out.ctypes.data) != 0:
        raise ValueError("hist: the image must be a multiple of the 256 x 32 tile")
    return out


def histeq_global(gray: np.ndarray, mapping: np.ndarray) -> np.ndarray:
    """clHistEq.histeqGlobal (eq_opencl.py:53-68)."""
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    mapping = np.ascontiguousarray(mapping, dtype=np.uint8)
    h, w = gray.shape
    out = np.zeros_like(gray)
    if _hist_lib().histeq_global_cl_run(gray.ctypes.data, w, h, gray.strides[0], out.ctypes.data, out.strides[0], mapping.ctypes.data) != 0:
        raise ValueError("histeq_global: the image must be a multiple of the 16 x 16 work-group")
    return out


def histeq_local_block(gray: np.ndarray, mappings: np.ndarray, blockshape) -> np.ndarray:
    """clHistEq.histeqLocalBlock (eq_opencl.py:70-89): mappings (ny, nx, 256), cast to float32 as the reference does."""
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    grid = np.ascontiguousarray(mappings, dtype=F32)
    h, w = gray.shape
    out = np.zeros_like(gray)
    rc = _hist_lib().histeq_local_block_cl_run(gray.ctypes.data, w, h, gray.strides[0], out.ctypes.data, out.strides[0], grid.ctypes.data,
                                               int(blockshape[1]), int(blockshape[0]), grid.shape[1], grid.shape[0])
    if rc != 0:
        raise ValueError("histeq_local_block: the image must be a multiple of the 16 x 16 work-group")
    return out
