"""CPU oracle for the RAISR hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oclcomputervision_b200/`` may import this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``) use it, and only
as the checker or the timed CPU baseline.

PARITY PIN: the reference (/root/reference/super_resolution/raisr.py + raisr.cl) has no tests, no golden
vectors, no CPU path, and its shipped kernel returns right after the bilinear upscale (raisr.cl:219-230).
The pin is made here instead: oracle/build_ref.py compiles the reference's own kernel source against an
OpenCL-C shim (oracle/_ref/, see raisr_cl_ref.py) and tests/test_ref_pin.py holds both restatements below to
its outputs (tests/golden/ref_cl.npz) -- the shipped kernel bit for bit, the full text (``quirks="as_written"``)
within 1 LSB except hashes that rounding decides, and the INTENDED semantics (the default) the same way against the
reference's text with its three slips corrected, one token each (build_ref.INTENDED_FIXES).  True binary16 arithmetic
is compared statistically only.  Two independent restatements live here and are also
checked against each other, against closed-form cases and against the committed fixtures in tests/golden/:

* ``raisr_ref``      numpy, written from the OpenCL text stage by stage (this file)
* ``raisr_ref_c``    plain C (oracle/raisr_oracle.c) loaded through ctypes; also the timed CPU baseline

Evaluation order (identical in both; see the header of raisr_oracle.c):
  coordinate map divide-then-multiply (raisr.cl:209); bilinear in the expression order of
  raisr.cl:60; Sobel as a flipped 3x3 convolution (raisr.cl:43-46,235-253); separable 9-tap
  Gaussian, horizontal then vertical, ascending index with fused multiply-add (raisr.cl:258-276,
  raisr.py:48-60); eigen/hash per raisr.cl:278-317 with the three intended-semantics fixes of
  SURVEY.md section 8(a); dot i-outer j-inner with fused multiply-add (raisr.cl:322-330).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np

FLEN = 11
MARGIN = 5
GRAD = 4
F32 = np.float32
PI_F = F32(np.pi)

DEFAULT_STRENGTH_Q = np.array([0.0001, 0.001], dtype=F32)  # raisr.py:112
DEFAULT_COHERENCE_Q = np.array([0.25, 0.5], dtype=F32)  # raisr.py:114


def gaussian2d(shape=(3, 3), sigma=0.5) -> np.ndarray:
    """The reference's Gaussian mask (raisr.py:48-60, MATLAB fspecial('gaussian')): exp(-(x^2+y^2)/(2 sigma^2)) on the
    centred integer grid, entries below eps*max zeroed, normalised to sum 1 (float64)."""
    half_r, half_c = (float(shape[0]) - 1.0) / 2.0, (float(shape[1]) - 1.0) / 2.0
    rows = np.arange(-half_r, half_r + 1.0)[:, None]
    cols = np.arange(-half_c, half_c + 1.0)[None, :]
    mask = np.exp(-(cols * cols + rows * rows) / (2.0 * sigma * sigma))
    mask[mask < np.finfo(mask.dtype).eps * mask.max()] = 0
    total = mask.sum()
    return mask / total if total != 0 else mask


def reference_gaussian81() -> np.ndarray:
    """The (81,) float32 weight vector of raisr.py:80-82 (diag / diag round trip included)."""
    g = gaussian2d([9, 9], 2)
    g = np.diag(g.ravel()).astype(F32)
    return np.diag(g).copy()


def gauss1d() -> np.ndarray:
    """Separable factor: g1[k] = exp(-k^2/8)/sum, k=-4..4, float32 (9,)."""
    k = np.arange(-GRAD, GRAD + 1, dtype=np.float64)
    g = np.exp(-k * k / 8.0)
    return (g / g.sum()).astype(F32)


def _fma(a, b, c):
    """float32 fused multiply-add emulated through float64 (a*b is exact in float64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


CUBIC_MATRIX = np.array([[0.0, -0.5, 1.0, -0.5], [1.0, 0.0, -2.5, 1.5], [0.0, 0.5, 2.0, -1.5], [0.0, 0.0, -0.5, 0.5]], F32)  # raisr.cl:63-68


def _cubic_weights(frac: np.ndarray) -> np.ndarray:
    """(n, 4) weights of raisr.cl:79-88,90-98: w_k = dot((1, u, u^2, u^3), cubic_matrix[k]), left to right."""
    u = frac.astype(F32)
    u2 = (u * u).astype(F32)
    u3 = (u2 * u).astype(F32)
    w = []
    for k in range(4):
        m = CUBIC_MATRIX[k]
        acc = (F32(1.0) * m[0] + u * m[1]).astype(F32)
        acc = (acc + u2 * m[2]).astype(F32)
        acc = (acc + u3 * m[3]).astype(F32)
        w.append(acc)
    return np.stack(w, axis=1)


def upscale_ext_cubic(src_u8: np.ndarray, s: int) -> np.ndarray:
    """Stage 1 with the reference's alternative cheap upscaler `cubic_sample` (raisr.cl:63-106, never called by
    the shipped kernel): 4x4 taps around floor(coord), weights from cubic_matrix,
    acc += (pix * xweight[j]) * yweight[i] with i outer / j inner, clamp to [0,1].  Same coordinates and
    CLAMP_TO_EDGE reads as the bilinear stage."""
    sh, sw = src_u8.shape
    dw, dh = sw * s, sh * s

    def axis(n_dst, n_src):
        pos = np.arange(-MARGIN, n_dst + MARGIN, dtype=np.int64).astype(F32)
        f = (pos / F32(n_dst - 1)) * F32(n_src - 1)
        fl = np.floor(f)
        i0 = fl.astype(np.int64)
        idx = np.stack([np.clip(i0 - 1 + k, 0, n_src - 1) for k in range(4)], axis=1)
        return idx, _cubic_weights((f - fl).astype(F32))

    xi, xw = axis(dw, sw)
    yi, yw = axis(dh, sh)
    p = src_u8.astype(F32) / F32(255.0)
    acc = np.zeros((yi.shape[0], xi.shape[0]), F32)
    for i in range(4):
        for j in range(4):
            pix = p[np.ix_(yi[:, i], xi[:, j])]
            acc = (acc + ((pix * xw[None, :, j]).astype(F32) * yw[:, i, None]).astype(F32)).astype(F32)
    return np.clip(acc, F32(0), F32(1)).astype(F32)


def upscale_ext(src_u8: np.ndarray, s: int) -> np.ndarray:
    """Stage 1 on the extended (dh+10, dw+10) domain (raisr.cl:48-61,171-190,198-217)."""
    sh, sw = src_u8.shape
    dw, dh = sw * s, sh * s

    def axis(n_dst, n_src):
        pos = np.arange(-MARGIN, n_dst + MARGIN, dtype=np.int64).astype(F32)
        f = (pos / F32(n_dst - 1)) * F32(n_src - 1)  # raisr.cl:209: divide, then multiply
        fl = np.floor(f)
        frac = (f - fl).astype(F32)
        i0 = fl.astype(np.int64)
        return np.clip(i0, 0, n_src - 1), np.clip(i0 + 1, 0, n_src - 1), frac

    x0, x1, u = axis(dw, sw)
    y0, y1, v = axis(dh, sh)
    p = src_u8.astype(F32) / F32(255.0)  # UNORM_INT8 decode (raisr.py:98, read_imagef)
    p00 = p[np.ix_(y0, x0)]
    p01 = p[np.ix_(y0, x1)]
    p10 = p[np.ix_(y1, x0)]
    p11 = p[np.ix_(y1, x1)]
    one = F32(1.0)
    omu, omv = (one - u)[None, :], (one - v)[:, None]
    uu, vv = u[None, :], v[:, None]
    acc = (omu * omv) * p00  # raisr.cl:60, left to right
    acc = acc + (uu * omv) * p01
    acc = acc + (omu * vv) * p10
    acc = acc + (uu * vv) * p11
    return acc.astype(F32)


def tensor(Uext: np.ndarray):
    """Stages 3-4: Sobel (flipped conv, raisr.cl:43-46) and the Gaussian-weighted tensor."""
    two = F32(2.0)
    a, b, c = Uext[:-2], Uext[1:-1], Uext[2:]
    d0 = a[:, :-2] - a[:, 2:]
    d1 = b[:, :-2] - b[:, 2:]
    d2 = c[:, :-2] - c[:, 2:]
    gx = (d0 + two * d1) + d2
    s0 = (a[:, :-2] + two * a[:, 1:-1]) + a[:, 2:]
    s2 = (c[:, :-2] + two * c[:, 1:-1]) + c[:, 2:]
    gy = s0 - s2
    g = gauss1d()
    planes = []
    for prod in (gx * gx, gx * gy, gy * gy):
        prod = prod.astype(F32)  # (dh+8, dw+8)
        w = prod.shape[1] - 2 * GRAD
        acc = g[0] * prod[:, 0:w]
        for k in range(1, 2 * GRAD + 1):
            acc = _fma(np.broadcast_to(g[k], acc.shape), prod[:, k : k + w], acc)
        hgt = acc.shape[0] - 2 * GRAD
        out = g[0] * acc[0:hgt]
        for k in range(1, 2 * GRAD + 1):
            out = _fma(np.broadcast_to(g[k], out.shape), acc[k : k + hgt], out)
        planes.append(out)
    return planes  # ma, mb, md each (dh, dw)


def eigen_hash(ma, mb, md, s, n_angle=24, n_strength=3, n_coherence=3,
               strength_q=DEFAULT_STRENGTH_Q, coherence_q=DEFAULT_COHERENCE_Q, quirks="intended"):
    """Stages 5-6 (raisr.cl:278-317).  quirks="intended" (default) or "as_written": the literal kernel
    text, i.e. ma = sum w*gx*gy (raisr.cl:271), coherence bucket compares L1 (raisr.cl:310) and strength
    is left out of the hash (raisr.cl:316)."""
    assert quirks in ("intended", "as_written")
    literal = quirks == "as_written"
    if literal:
        ma = mb
    with np.errstate(invalid="ignore", divide="ignore"):
        T = ma + md
        D = ma * md - mb * mb
        rad = (T * T) * F32(0.25) - D
        rad = np.where(rad > 0, rad, F32(0)).astype(F32)
        sq = np.sqrt(rad)
        ht = T * F32(0.5)
        L1 = (ht + sq).astype(F32)
        L2 = (ht - sq).astype(F32)
        L2 = np.where(L2 > 0, L2, F32(0)).astype(F32)
        theta = np.arctan2(mb, (L1 - md).astype(F32)).astype(F32)
        theta = np.where(theta < 0, theta + PI_F, theta).astype(F32)
        s1, s2 = np.sqrt(L1), np.sqrt(L2)
        den = (s1 + s2).astype(F32)
        coh = np.where(den != 0, (s1 - s2) / np.where(den != 0, den, F32(1)), F32(0)).astype(F32)
    a = ((theta / PI_F) * F32(n_angle)).astype(np.int32)
    a = np.clip(a, 0, n_angle - 1)
    sq_ = np.asarray(strength_q, dtype=F32)
    cq_ = np.asarray(coherence_q, dtype=F32)
    si = np.full(a.shape, n_strength - 1, np.int32)
    for i in range(n_strength - 2, -1, -1):
        si = np.where(L1 < sq_[i], i, si)
    ci = np.full(a.shape, n_coherence - 1, np.int32)
    for i in range(n_coherence - 2, -1, -1):
        ci = np.where((L1 if literal else coh) < cq_[i], i, ci)
    if literal:
        si = np.zeros_like(si)
    dh, dw = a.shape
    yy, xx = np.mgrid[0:dh, 0:dw]
    ptype = (yy % s) * s + (xx % s)
    h = ((a * n_strength + si) * n_coherence + ci) * (s * s) + ptype
    return theta, L1, coh, h.astype(np.int32)


def gather_dot(Uext, hash_, filters):
    """Stage 7 (raisr.cl:322-330): acc = fma(patch[i][j], tap[i*11+j], acc), i outer, j inner."""
    dh, dw = hash_.shape
    flt = np.ascontiguousarray(filters, dtype=F32).reshape(-1, FLEN * FLEN)
    taps = flt[hash_]  # (dh, dw, 121)
    acc = np.zeros((dh, dw), F32)
    for i in range(FLEN):
        for j in range(FLEN):
            acc = _fma(Uext[i : i + dh, j : j + dw], taps[:, :, i * FLEN + j], acc)
    return acc


def raisr_ref(src_u8: np.ndarray, filters: Optional[np.ndarray], s: int = 2, *,
              n_angle=24, n_strength=3, n_coherence=3,
              strength_q=DEFAULT_STRENGTH_Q, coherence_q=DEFAULT_COHERENCE_Q,
              quirks: str = "intended", taps: str = "fp32", upscaler: str = "bilinear") -> Dict[str, np.ndarray]:
    """numpy restatement.  Returns U, Uext, angle, L1, coherence, hash, out_f32, out_u8.

    upscaler: "bilinear" (what raisr.cl:198-217 calls) | "bicubic" (its unused cubic_sample, raisr.cl:63-106).
    quirks: "intended" | "as_written" (see eigen_hash).  taps: "fp32" | "fp16" -- "fp16" rounds every tap
    to half precision first, as the reference's `(half)pf[i*FILTER_LEN+j]` does (raisr.cl:328); the
    arithmetic stays fp32 either way."""
    assert taps in ("fp32", "fp16") and upscaler in ("bilinear", "bicubic")
    src_u8 = np.ascontiguousarray(src_u8, dtype=np.uint8)
    Uext = upscale_ext(src_u8, s) if upscaler == "bilinear" else upscale_ext_cubic(src_u8, s)
    ma, mb, md = tensor(Uext)
    theta, L1, coh, h = eigen_hash(ma, mb, md, s, n_angle, n_strength, n_coherence,
                                   strength_q, coherence_q, quirks)
    if filters is not None and taps == "fp16":
        filters = np.asarray(filters, dtype=F32).astype(np.float16).astype(F32)
    res = dict(Uext=Uext, U=Uext[MARGIN:-MARGIN, MARGIN:-MARGIN].copy(), angle=theta, L1=L1,
               coherence=coh, hash=h, ma=ma, mb=mb, md=md)
    if filters is not None:
        acc = gather_dot(Uext, h, filters)
        out = np.clip(acc, F32(0), F32(1)).astype(F32)
        res["out_f32"] = out
        res["out_u8"] = np.rint(out * F32(255.0)).astype(np.uint8)
    return res


def edge_distance(res: Dict[str, np.ndarray], n_angle=24,
                  strength_q=DEFAULT_STRENGTH_Q, coherence_q=DEFAULT_COHERENCE_Q, quirks="intended") -> np.ndarray:
    """Per-pixel distance of the float angle*n/pi, L1 and coherence to their nearest bin edge
    (the quantity north_star excuses below 1e-5).  With quirks="as_written" the only thresholds that reach
    the hash are the angle bins and L1 against the *coherence* quantisers."""
    a = (res["angle"].astype(np.float64) / np.pi) * n_angle
    da = np.abs(a - np.rint(a))
    l1 = res["L1"].astype(np.float64)
    ds = np.min([np.abs(l1 - float(q)) for q in strength_q], axis=0)
    co = res["coherence"].astype(np.float64)
    dc = np.min([np.abs(co - float(q)) for q in coherence_q], axis=0)
    if quirks == "as_written":
        return np.minimum(da, np.min([np.abs(l1 - float(q)) for q in coherence_q], axis=0))
    return np.minimum(np.minimum(da, ds), dc)


# --------------------------------------------------------------------------------------------
# C twin
# --------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraisr_oracle.so")
_lib = None


def build_c(force: bool = False) -> str:
    src = os.path.join(_HERE, "raisr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def _load():
    global _lib
    if _lib is None:
        src_c = os.path.join(_HERE, "raisr_oracle.c")
        if not os.path.exists(_SO) or (os.path.exists(src_c) and os.path.getmtime(src_c) > os.path.getmtime(_SO)):
            build_c(force=True)
        lib = ctypes.CDLL(_SO)
        vp = ctypes.c_void_p
        lib.raisr_oracle_run.restype = ctypes.c_int
        lib.raisr_oracle_run.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int,
                                         vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                         vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int]
        lib.raisr_oracle_run_ex.restype = ctypes.c_int
        lib.raisr_oracle_run_ex.argtypes = list(lib.raisr_oracle_run.argtypes) + [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.raisr_oracle_bilinear_u8.restype = ctypes.c_int
        lib.raisr_oracle_bilinear_u8.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                                 ctypes.c_int, vp]
        lib.raisr_oracle_max_threads.restype = ctypes.c_int
        lib.raisr_oracle_run_bgra.restype = ctypes.c_int
        lib.raisr_oracle_run_bgra.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, vp, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int]
        lib.raisr_oracle_run_bgra_ex.restype = ctypes.c_int
        lib.raisr_oracle_run_bgra_ex.argtypes = list(lib.raisr_oracle_run_bgra.argtypes) + [vp, vp, vp, ctypes.c_int]
        lib.raisr_oracle_run_bgra_q.restype = ctypes.c_int
        lib.raisr_oracle_run_bgra_q.argtypes = list(lib.raisr_oracle_run_bgra_ex.argtypes) + [ctypes.c_int, ctypes.c_int]
        lib.raisr_oracle_resize_u8.restype = ctypes.c_int
        lib.raisr_oracle_resize_u8.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, vp,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
        _lib = lib
    return _lib


def c_max_threads() -> int:
    return int(_load().raisr_oracle_max_threads())


def raisr_ref_c(src_u8: np.ndarray, filters: Optional[np.ndarray], s: int = 2, *,
                n_angle=24, n_strength=3, n_coherence=3,
                strength_q=DEFAULT_STRENGTH_Q, coherence_q=DEFAULT_COHERENCE_Q,
                nthreads: int = 0, want=("U", "Uext", "angle", "L1", "coherence", "hash",
                                         "out_f32", "out_u8"),
                quirks: str = "intended", taps: str = "fp32", upscaler: str = "bilinear") -> Dict[str, np.ndarray]:
    """C restatement (oracle/raisr_oracle.c).  Same keys and options as raisr_ref."""
    assert quirks in ("intended", "as_written") and taps in ("fp32", "fp16") and upscaler in ("bilinear", "bicubic")
    lib = _load()
    src_u8 = np.ascontiguousarray(src_u8, dtype=np.uint8)
    sh, sw = src_u8.shape
    dh, dw = sh * s, sw * s
    sq = np.ascontiguousarray(strength_q, dtype=F32)
    cq = np.ascontiguousarray(coherence_q, dtype=F32)
    flt = None
    if filters is not None:
        flt = np.ascontiguousarray(filters, dtype=F32)
        assert flt.size == n_angle * n_strength * n_coherence * s * s * FLEN * FLEN
    res: Dict[str, np.ndarray] = {}
    shapes = dict(U=((dh, dw), F32), Uext=((dh + 2 * MARGIN, dw + 2 * MARGIN), F32),
                  angle=((dh, dw), F32), L1=((dh, dw), F32), coherence=((dh, dw), F32),
                  hash=((dh, dw), np.int32), out_f32=((dh, dw), F32), out_u8=((dh, dw), np.uint8))
    for k in want:
        if k in ("out_f32", "out_u8") and flt is None:
            continue
        shp, dt = shapes[k]
        res[k] = np.empty(shp, dt)

    def ptr(k):
        return res[k].ctypes.data if k in res else None

    rc = lib.raisr_oracle_run_ex(src_u8.ctypes.data, sw, sh, src_u8.strides[0], s,
                                 flt.ctypes.data if flt is not None else None,
                                 n_angle, n_strength, n_coherence, sq.ctypes.data, cq.ctypes.data,
                                 ptr("U"), ptr("Uext"), ptr("angle"), ptr("L1"), ptr("coherence"),
                                 ptr("hash"), ptr("out_f32"), ptr("out_u8"), int(nthreads),
                                 int(quirks == "as_written"), int(taps == "fp16"), int(upscaler == "bicubic"))
    if rc != 0:
        raise RuntimeError("raisr_oracle_run failed (%d)" % rc)
    return res


def bilinear_u8_c(src_u8: np.ndarray, s: int) -> np.ndarray:
    lib = _load()
    src_u8 = np.ascontiguousarray(src_u8, dtype=np.uint8)
    sh, sw = src_u8.shape
    dst = np.empty((sh * s, sw * s), np.uint8)
    rc = lib.raisr_oracle_bilinear_u8(src_u8.ctypes.data, sw, sh, src_u8.strides[0], s,
                                      dst.ctypes.data)
    if rc != 0:
        raise RuntimeError("raisr_oracle_bilinear_u8 failed")
    return dst


RESIZE_MODES = {"bilinear_lds": 0, "bicubic": 1, "bicubic_lds": 1, "bilinear": 2}


def resize_u8_c(src: np.ndarray, out_hw, mode: str) -> np.ndarray:
    """C restatement of basic/interpolation.cl (see raisr_oracle.c: raisr_oracle_resize_u8)."""
    lib = _load()
    src = np.ascontiguousarray(src, dtype=np.uint8)
    ch = 1 if src.ndim == 2 else src.shape[2]
    dh, dw = out_hw
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.uint8)
    rc = lib.raisr_oracle_resize_u8(src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], ch, dst.ctypes.data,
                                    dw, dh, dst.strides[0], RESIZE_MODES[mode])
    if rc != 0:
        raise RuntimeError("raisr_oracle_resize_u8 failed")
    return dst


def raisr_ref_bgra_c(src_bgra: np.ndarray, filters: np.ndarray, s: int = 2, *, n_angle=24, n_strength=3, n_coherence=3,
                     strength_q=DEFAULT_STRENGTH_Q, coherence_q=DEFAULT_COHERENCE_Q, nthreads: int = 0,
                     upscaler: str = "bilinear", quirks: str = "intended", taps: str = "fp32") -> Dict[str, np.ndarray]:
    """C restatement of the colour (BGRA) path (raisr_oracle.c: raisr_oracle_run_bgra).  upscaler="bicubic": stage 1 is
    the reference's cubic_sample (raisr.cl:63-106) on each of the four channels."""
    assert upscaler in ("bilinear", "bicubic") and quirks in ("intended", "as_written") and taps in ("fp32", "fp16")
    lib = _load()
    src = np.ascontiguousarray(src_bgra, dtype=np.uint8)
    assert src.ndim == 3 and src.shape[2] == 4
    sh, sw = src.shape[:2]
    dh, dw = sh * s, sw * s
    flt = np.ascontiguousarray(filters, dtype=F32)
    sq = np.ascontiguousarray(strength_q, dtype=F32)
    cq = np.ascontiguousarray(coherence_q, dtype=F32)
    res = dict(hash=np.empty((dh, dw), np.int32), out_f32=np.empty((dh, dw, 4), F32), out_u8=np.empty((dh, dw, 4), np.uint8),
               angle=np.empty((dh, dw), F32), L1=np.empty((dh, dw), F32), coherence=np.empty((dh, dw), F32))
    rc = lib.raisr_oracle_run_bgra_q(src.ctypes.data, sw, sh, src.strides[0], s, flt.ctypes.data, n_angle, n_strength, n_coherence,
                                      sq.ctypes.data, cq.ctypes.data, res["hash"].ctypes.data, res["out_f32"].ctypes.data,
                                      res["out_u8"].ctypes.data, int(nthreads), res["angle"].ctypes.data, res["L1"].ctypes.data,
                                     res["coherence"].ctypes.data, int(upscaler == "bicubic"), int(quirks == "as_written"),
                                     int(taps == "fp16"))
    if rc != 0:
        raise RuntimeError("raisr_oracle_run_bgra failed (%d)" % rc)
    return res
