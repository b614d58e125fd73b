"""Histogram-equalisation path: oracle vs the reference's own CPU outputs
(tests/golden/histeq_ref.npz, made by oracle/make_golden_histeq.py from /root/reference/histeq), and the
CUDA path vs both.  Integer / byte work: every comparison is bit-exact."""
import os

import numpy as np
import pytest

from oracle import histeq_oracle as ho
from oracle.make_golden_histeq import test_image as golden_image

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "histeq_ref.npz"))


def _gold_gray():
    h, w, seed = (int(v) for v in GOLD["img_shape_seed"])
    g = golden_image(h, w, seed)
    crc = [int(g.astype(np.uint64).sum()), int((g.astype(np.uint64) * (np.arange(g.size).reshape(g.shape) % 251)).sum())]
    assert crc == [int(v) for v in GOLD["img_crc"]], "golden input image no longer reproduces"
    return g


def _tf_cases():
    for i in range(len(GOLD["tf_hist"])):
        a, p, c = GOLD["tf_params"][i]
        yield GOLD["tf_hist"][i], (int(a) if a == int(a) else float(a)), float(p), float(c), GOLD["tf_curve"][i]


# ---------------------------------------------------------------- CPU: oracle + host logic vs reference outputs
def test_oracle_transfer_func_matches_reference():
    for hist, a, p, c, want in _tf_cases():
        assert np.array_equal(ho.transfer_func(hist, a, p, c), want)


def test_product_transfer_func_matches_reference():
    from oclcomputervision_b200.histeq import calc_transfer_func
    for hist, a, p, c, want in _tf_cases():
        got = calc_transfer_func(hist, a, p, c)
        assert got.dtype == np.float64 and np.array_equal(got, want)
    with pytest.raises(ValueError):
        calc_transfer_func(np.zeros(256, np.uint32), 1, 0.05, 2)


def test_oracle_images_match_reference():
    g = _gold_gray()
    assert np.array_equal(ho.histeq_global(g), GOLD["global_default"])
    assert np.array_equal(ho.histeq_global(g, 0.6, 0.02, 3)[::16], GOLD["global_a06_rows"])
    assert np.array_equal(ho.histeq_local_block(g), GOLD["local_default"])
    assert np.array_equal(ho.histeq_local_block(g, 0.7, 0.03, 2.5, (128, 256)), GOLD["local_128x256"])


def test_oracle_hist_grid_properties():
    g = golden_image(96, 600, 3)
    grid = ho.hist_grid(g)
    assert grid.shape == (3, 2, 256) and grid.dtype == np.uint32
    assert (grid.sum(axis=2) == 256 * 32).all()
    assert np.array_equal(grid.sum(axis=(0, 1)), np.bincount(g[:96, :512].ravel(), minlength=256))


def test_no_cpu_path():
    from oclcomputervision_b200 import histeq
    g = np.zeros((256, 256), np.uint8)
    with pytest.raises(NotImplementedError):
        histeq.histeq_global(g, use_gpu=False)
    with pytest.raises(NotImplementedError):
        histeq.histeq_local_block(g, use_gpu=False)
    src = open(histeq.__file__).read()
    assert "oracle" not in src.replace("oclcomputervision", "")


# ---------------------------------------------------------------- GPU: the CUDA path through the C-ABI
@pytest.fixture(scope="module")
def cleq():
    from oclcomputervision_b200.histeq import clHistEq
    return clHistEq.getInstance()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(32, 256), (64, 512), (100, 700), (512, 512), (1080, 1920), (2160, 3840)])
def test_gpu_hist_grid_bit_exact(cleq, shape):
    g = golden_image(shape[0], shape[1], 5)
    got, ms = cleq.histGrid(g)
    want = ho.hist_grid(g)
    assert got.shape == want.shape and got.dtype == np.uint32 and ms >= 0
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_hist_grid_extremes_and_views(cleq):
    for fill in (0, 255):
        g = np.full((64, 512), fill, np.uint8)
        got, _ = cleq.histGrid(g)
        assert (got[..., fill] == 8192).all() and got.sum() == g.size
    big = golden_image(200, 1100, 9)
    view = big[3:3 + 96, 5:5 + 777]           # unaligned base: the wrapper makes it contiguous, rows 777 bytes
    got, _ = cleq.histGrid(view)
    assert np.array_equal(got, ho.hist_grid(np.ascontiguousarray(view)))


@pytest.mark.gpu
def test_gpu_global_and_local_match_reference_golden(cleq):
    from oclcomputervision_b200 import histeq
    g = _gold_gray()
    assert np.array_equal(histeq.histeq_global(g), GOLD["global_default"])
    assert np.array_equal(histeq.histeq_global(g, alpha=0.6, punch=0.02, clip=3)[::16], GOLD["global_a06_rows"])
    assert np.array_equal(histeq.histeq_local_block(g), GOLD["local_default"])
    assert np.array_equal(histeq.histeq_local_block(g, alpha=0.7, punch=0.03, clip=2.5, blockshape=(128, 256)),
                          GOLD["local_128x256"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,block", [((1080, 1920), (256, 256)), ((2160, 3840), (256, 256)),
                                         ((777, 1031), (128, 256)), ((2160, 3840), (96, 512)),
                                         ((300, 700), (128, 256)), ((256, 256), (256, 256)), ((511, 1023), (224, 512))])
def test_gpu_local_block_vs_oracle(cleq, shape, block):
    from oclcomputervision_b200 import histeq
    g = golden_image(shape[0], shape[1], 21)
    got = histeq.histeq_local_block(g, blockshape=block)
    want = ho.histeq_local_block(g, blockshape=block)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_lut_passes_vs_oracle(cleq):
    rng = np.random.default_rng(0)
    for shape in [(1, 1), (7, 33), (1080, 1920), (333, 4099)]:
        g = rng.integers(0, 256, shape, dtype=np.uint8)
        m = rng.permutation(256).astype(np.uint8)
        got, _ = cleq.histeqGlobal(g, m)
        assert np.array_equal(got, m[g])
    g = rng.integers(0, 256, (300, 520), dtype=np.uint8)
    maps = (rng.random((2, 2, 256)) * 300 - 20).astype(np.float32)     # exercises both clamps
    got, _ = cleq.histeqLocalBlock(g, maps, (150, 260))                 # non power-of-two blocks: inexact weights
    assert np.array_equal(got, ho.local_block_apply(g, maps, (150, 260)))


@pytest.mark.gpu
def test_gpu_histeq_errors(cleq):
    from oclcomputervision_b200._cabi import RaisrError
    with pytest.raises(RaisrError):
        cleq.histGrid(np.zeros((16, 256), np.uint8))            # smaller than one tile
    with pytest.raises(ValueError):
        cleq.histGrid(np.zeros((64, 512, 3), np.uint8))
    with pytest.raises(ValueError):
        cleq.histeqGlobal(np.zeros((64, 512), np.uint8), np.zeros(100, np.uint8))
    with pytest.raises(ValueError):
        cleq.histeqLocalBlock(np.zeros((512, 512), np.uint8), np.zeros((3, 3, 256), np.float32), (256, 256))
