"""Stand-alone interpolation: oracle sanity on CPU, CUDA parity on GPU.
The reference's own check only prints PSNR against scipy.interp2d (basic/interpolation.py:121-184)."""
import numpy as np
import pytest

from oracle import raisr_oracle as O
from oclcomputervision_b200 import synth


def _ramp(h, w, ch):
    x = np.linspace(10, 240, w)[None, :] * np.ones((h, 1))
    img = np.rint(x).astype(np.uint8)
    return img if ch == 1 else np.stack([img, img[::-1], img, np.full_like(img, 255)], axis=2).copy()


@pytest.mark.parametrize("mode", ["bilinear_lds", "bicubic"])
def test_oracle_reproduces_a_linear_ramp(mode):
    # bilinear and Catmull-Rom both have linear precision under the align-corners map
    src = np.rint(np.linspace(0, 200, 41))[None, :].repeat(9, 0).astype(np.uint8)   # exact steps of 5
    out = O.resize_u8_c(src, (18, 81), mode)
    want = np.rint(np.linspace(0, 200, 81))[None, :].repeat(18, 0)
    assert np.abs(out.astype(int) - want.astype(int)).max() <= 1


def test_oracle_identity_size_is_identity():
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (17, 23, 4), dtype=np.uint8)
    for mode in ("bilinear_lds", "bicubic"):
        assert np.array_equal(O.resize_u8_c(src, (17, 23), mode), src)


def test_oracle_gray_bilinear_lds_is_raisr_stage1():
    src = synth.synthetic_frame(33, 47, 3, sigma=1.5)
    assert np.array_equal(O.resize_u8_c(src, (66, 94), "bilinear_lds"), O.bilinear_u8_c(src, 2))


def test_oracle_bilinear_simple_uses_pixel_centres_shifted_map():
    # CLK_NORMALIZED_COORDS_TRUE: x_out = 0 samples texel position -0.5 -> clamps to texel 0;
    # the last output column samples position w_in - 0.5 -> texel w_in - 1
    src = _ramp(6, 12, 1)
    out = O.resize_u8_c(src, (6, 30), "bilinear")
    assert out[0, 0] == src[0, 0] and out[0, -1] == src[0, -1]
    lds = O.resize_u8_c(src, (6, 30), "bilinear_lds")
    assert not np.array_equal(out, lds)          # not the same mapping as bilinear_lds


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bilinear", "bilinear_lds", "bicubic", "bicubic_lds"])
@pytest.mark.parametrize("shape,out,ch", [((40, 52), (80, 104), 4), ((33, 47), (90, 61), 4), ((64, 64), (128, 128), 1),
                                          ((5, 7), (2, 2), 4), ((128, 96), (100, 300), 1)])
def test_gpu_matches_oracle_bit_exactly(mode, shape, out, ch):
    from oclcomputervision_b200.interpolation import clUtility
    rng = np.random.default_rng(shape[0] + out[1])
    src = rng.integers(0, 256, shape + ((4,) if ch == 4 else ()), dtype=np.uint8)
    dst = np.zeros(out + ((4,) if ch == 4 else ()), np.uint8)
    util = clUtility()
    ms = getattr(util, mode)(src, dst)
    assert len(ms) == 3
    assert np.array_equal(dst, O.resize_u8_c(src, out, mode))
    util.close()


@pytest.mark.gpu
def test_gpu_reference_demo_shapes_and_psnr():
    # the reference's __main__: 1024x1024 BGRA -> 2048x2048, PSNR of each kernel against a CPU interpolation
    from oclcomputervision_b200.interpolation import clUtility, psnr
    img = synth.synthetic_frame(256, 256, 11, sigma=3.0)
    bgra = np.stack([img, img[::-1], img.T.copy(), np.full_like(img, 255)], axis=2).copy()
    util = clUtility()
    outs = {}
    for mode in ("bilinear_lds", "bicubic_lds"):
        dst = np.zeros((512, 512, 4), np.uint8)
        getattr(util, mode)(bgra, dst)
        outs[mode] = dst
    assert psnr(outs["bilinear_lds"], outs["bicubic_lds"]) > 35.0
    assert psnr(outs["bilinear_lds"], outs["bilinear_lds"]) == float("inf")
    util.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,out", [((270, 480), (540, 960)), ((100, 300), (200, 600)), ((64, 100), (192, 300)), ((37, 53), (74, 106)),
                                       ((50, 77), (45, 61)), ((300, 700), (301, 1203)), ((9, 16), (36, 64)), ((1, 8), (2, 16)),
                                       ((120, 2000), (130, 2100)), ((700, 64), (90, 40))])
@pytest.mark.parametrize("mode,ch", [("bilinear_lds", 1), ("bilinear", 1), ("bilinear_lds", 4), ("bilinear", 4), ("bicubic", 4)])
def test_gpu_fast_gray_bilinear_kernel_bit_exact(shape, out, mode, ch):
    """The fast resize kernels (csrc/raisr_resize.cuh: bilinear with four pixels per thread for gray -- also
    ClRaisr.bilinear_only, the shipped raisr kernel's output --, bilinear and bicubic with one BGRA pixel per thread)
    against the oracle and against the generic kernel, on up-scales, down-scales, ragged widths (stores of 1-3 leftover bytes) and tiles whose window is larger
    than the fast path allows (falls back)."""
    import ctypes
    from oclcomputervision_b200 import _cabi
    from oclcomputervision_b200.interpolation import clUtility
    rng = np.random.default_rng(out[1])
    src = rng.integers(0, 256, shape + ((4,) if ch == 4 else ()), dtype=np.uint8)
    want = O.resize_u8_c(src, out, mode)
    util = clUtility()
    lib = _cabi.load()
    res = []
    for fast in (1, 0):
        _cabi.check(lib.raisr_set_option(util._h, b"resize_fast", fast))
        pitch = (out[1] + 3) // 4 * 4                      # 4-byte aligned rows qualify for the fast kernels
        if ch == 4:
            canvas = np.full((out[0] + 1, out[1] + 2, 4), 173, np.uint8)
            dst = canvas[:out[0], :out[1]]
        else:
            canvas = np.full((out[0] + 1, pitch + 4), 173, np.uint8)
            dst = canvas[:out[0], :out[1]]
        getattr(util, mode)(src, dst)
        assert np.array_equal(dst, want), (mode, fast)
        assert (canvas[:, out[1]:] == 173).all() and (canvas[-1] == 173).all()      # nothing written beyond the image
        res.append(dst.copy())
    assert np.array_equal(res[0], res[1])
    util.close()
