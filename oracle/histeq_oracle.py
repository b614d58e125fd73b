"""CPU restatement of the reference's histogram-equalisation path.

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and tools/bench_histeq.py's cpu leg may import
this; the product (oclcomputervision_b200/histeq.py) never does.

PARITY PINNED for everything the reference can run on a CPU: /root/reference/histeq/eq_global.py and
eq_local_block.py carry a `use_gpu=False` numpy path; oracle/make_golden_histeq.py imports them (with
`pyopencl`/`matplotlib` stubbed, they are only needed by the OpenCL branch) and stores their outputs in
tests/golden/histeq_ref.npz, which tests/test_histeq.py checks this file against -- bit-exact, including
the local-block blend: the OpenCL kernel (hist.cl:139-144) works in fp32, and under numpy 2 scalar
promotion (python-float weight x np.float32 table entry -> float32) so does the reference's Python loop
(eq_local_block.py:66-77), so `local_block_apply` below restates both.  The weights s, t and their
products are exact in fp32 for power-of-two block sizes, which is what makes the two agree to the bit.
The OpenCL kernels themselves (hist.cl, run on the CPU through oracle/build_ref.py) are a second pin:
tests/test_ref_pin.py::test_hist_kernels_match_the_histeq_oracle, also with block sizes that are not powers of two.
"""
from __future__ import annotations

import numpy as np

HIST_BINS = 256   # eq_opencl.py:13
TILE_ROWS = 32    # eq_opencl.py:14


def hist_grid(gray: np.ndarray) -> np.ndarray:
    """hist.cl:41-90 as launched by eq_opencl.py:37-51: one 256-bin histogram per 256x32 tile."""
    h, w = gray.shape
    ty, tx = h // TILE_ROWS, w // HIST_BINS
    out = np.zeros((ty, tx, HIST_BINS), np.uint32)
    for i in range(ty):
        for j in range(tx):
            tile = gray[i * TILE_ROWS:(i + 1) * TILE_ROWS, j * HIST_BINS:(j + 1) * HIST_BINS]
            out[i, j] = np.bincount(tile.ravel(), minlength=HIST_BINS)
    return out


def transfer_func(hist, alpha, punch, clip) -> np.ndarray:
    """eq_global.py:10-39 (float64): punched CDF, alpha blend with identity, gain limit."""
    hist = np.asarray(hist)
    level = np.arange(hist.size)
    cdf = np.cumsum(hist) / np.sum(hist)
    lo = int(np.nonzero(cdf >= punch)[0][0])
    hi = int(np.nonzero(cdf >= 1 - punch)[0][0])
    mid = hist[lo:hi]
    cdf[:lo] = 0
    cdf[hi:] = 1
    cdf[lo:hi] = np.cumsum(mid) / np.sum(mid)
    curve = np.clip(alpha * cdf * 255 + (1 - alpha) * level, 0, 255)
    return np.clip(curve, level / clip, level * clip)


def global_apply(gray: np.ndarray, mapping_u8: np.ndarray) -> np.ndarray:
    """hist.cl:92-102."""
    return mapping_u8[gray]


def block_mappings(gray, alpha, punch, clip, blockshape) -> np.ndarray:
    """eq_local_block.py:14-35: one transfer function per block (float32 grid)."""
    bh, bw = blockshape
    ny, nx = gray.shape[0] // bh, gray.shape[1] // bw
    maps = np.zeros((ny, nx, HIST_BINS), np.float32)
    for i in range(ny):
        for j in range(nx):
            hist = np.bincount(gray[i * bh:(i + 1) * bh, j * bw:(j + 1) * bw].ravel(), minlength=HIST_BINS)
            maps[i, j] = transfer_func(hist, alpha, punch, clip).astype(np.float32)
    return maps


def _block_geometry(n, block, nblocks):
    pos = np.arange(n)
    num = pos - block // 2
    b0 = np.where(num >= 0, num // block, -((-num) // block))     # C division truncates toward zero
    # an image wider than nblocks*block + block/2 makes the reference index one block past its tables
    # (undefined there); here such pixels stay in the last block, with the weight saturating at 1
    b0 = np.minimum(b0, nblocks - 1)
    centre = b0 * block + block // 2
    b1 = np.minimum(b0 + 1, nblocks - 1)
    return pos, b0, b1, centre


def local_block_apply(gray: np.ndarray, maps: np.ndarray, blockshape) -> np.ndarray:
    """hist.cl:104-147 in fp32, one rounding per operation, evaluated left to right."""
    bh, bw = blockshape
    ny, nx = maps.shape[:2]
    h, w = gray.shape
    f32 = np.float32
    x, bx0, bx1, cx = _block_geometry(w, bw, nx)
    y, by0, by1, cy = _block_geometry(h, bh, ny)
    s = np.clip((x - cx).astype(f32) / f32(bw), f32(0), f32(1))[None, :]
    t = np.clip((y - cy).astype(f32) / f32(bh), f32(0), f32(1))[:, None]
    v = gray.astype(np.intp)
    f00 = maps[by0[:, None], bx0[None, :], v]
    f01 = maps[by0[:, None], bx1[None, :], v]
    f10 = maps[by1[:, None], bx0[None, :], v]
    f11 = maps[by1[:, None], bx1[None, :], v]
    one = f32(1)
    acc = ((one - s) * (one - t)) * f00
    acc = acc + (s * (one - t)) * f01
    acc = acc + ((one - s) * t) * f10
    acc = acc + (s * t) * f11
    assert acc.dtype == np.float32
    return np.clip(acc, f32(0), f32(255)).astype(np.uint8)


def histeq_global(gray, alpha=1, punch=0.05, clip=2) -> np.ndarray:
    """eq_global.py:41-65 through the tile-histogram route of its OpenCL branch."""
    hist = hist_grid(gray).sum(axis=0).sum(axis=0)
    return global_apply(gray, transfer_func(hist, alpha, punch, clip).astype(np.uint8))


def histeq_local_block(gray, alpha=0.5, punch=0.05, clip=3, blockshape=(256, 256)) -> np.ndarray:
    return local_block_apply(gray, block_mappings(gray, alpha, punch, clip, blockshape), blockshape)
