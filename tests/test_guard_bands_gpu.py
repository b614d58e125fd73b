"""Out-of-bounds WRITE detection without a sanitizer: every device-pointer entry point writes into the
interior of a larger, pattern-filled allocation (odd pitches, unaligned starts, ragged sizes) and the
guard bands must come back untouched while the interior matches the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import histeq_oracle as ho
from oracle import raisr_oracle as O
from oclcomputervision_b200 import ClRaisr, _cabi, synth

torch = pytest.importorskip("torch")
PAT = 0xA5


def _framed(h, w, dtype, pad_rows=3, pad_left=16, pad_right=24):
    """(full tensor, view of the h x w interior, pitch in bytes)."""
    full = torch.full((h + 2 * pad_rows, pad_left + w + pad_right), PAT if dtype == torch.uint8 else -777.0, dtype=dtype, device="cuda")
    view = full[pad_rows:pad_rows + h, pad_left:pad_left + w]
    return full, view, full.stride(0) * full.element_size()


def _guards_intact(full, view_shape, dtype, pad_rows=3, pad_left=16):
    h, w = view_shape
    f = full.clone()
    f[pad_rows:pad_rows + h, pad_left:pad_left + w] = PAT if dtype == torch.uint8 else -777.0
    return bool((f == (PAT if dtype == torch.uint8 else -777.0)).all())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,s", [((37, 53), 2), ((40, 56), 2), ((24, 33), 3), ((19, 21), 4), ((131, 70), 2)])
@pytest.mark.parametrize("out", ["u8", "f32"])
def test_raisr_device_entry_respects_pitch_and_bounds(shape, s, out):
    sh, sw = shape
    src = synth.synthetic_frame(sh, sw, seed=77)
    flt = synth.random_filters(s, seed=9)
    r = ClRaisr(1, device=0)
    setattr(r, "filters_x%d" % s, flt)
    dsrc_full, dsrc, spitch = _framed(sh, sw, torch.uint8, pad_left=5, pad_right=9)
    dsrc.copy_(torch.from_numpy(src).cuda())
    dt = torch.uint8 if out == "u8" else torch.float32
    full, view, dpitch = _framed(sh * s, sw * s, dt)
    torch.cuda.synchronize()
    r.upsample_device(dsrc.data_ptr(), sw, sh, spitch, view.data_ptr(), dpitch, s, 1, np.uint8 if out == "u8" else np.float32)
    r.sync()
    assert _guards_intact(full, (sh * s, sw * s), dt)
    want = O.raisr_ref_c(src, flt, s)
    got = view.cpu().numpy()
    ok = r.debug_hash(src, s)[0] == want["hash"]
    if out == "u8":
        assert np.abs(got.astype(int) - want["out_u8"].astype(int))[ok].max() <= 1
    else:
        assert np.abs(got - want["out_f32"])[ok].max() <= 1e-4
    r.close()


@pytest.mark.gpu
def test_raisr_batch_device_entry_guards():
    """Three frames back to back (frame stride = pitch * height) inside one guarded allocation."""
    sh, sw, s, n = 45, 61, 2, 3
    flt = synth.random_filters(s, seed=2)
    r = ClRaisr(1, device=0)
    r.filters_x2 = flt
    frames = np.stack([synthetic for synthetic in (synth.synthetic_frame(sh, sw, seed=k) for k in range(n))])
    dsrc = torch.from_numpy(frames).cuda()
    dh, dw = sh * s, sw * s
    pitch = dw + 40
    full = torch.full((n * dh + 6, pitch), PAT, dtype=torch.uint8, device="cuda")
    base = full[3:3 + n * dh]
    torch.cuda.synchronize()
    r.upsample_device(dsrc.data_ptr(), sw, sh, sw, base.data_ptr(), pitch, s, n, np.uint8)
    r.sync()
    got = base.cpu().numpy()
    assert (full[:3] == PAT).all() and (full[3 + n * dh:] == PAT).all() and (got[:, dw:] == PAT).all()
    for k in range(n):
        want = O.raisr_ref_c(frames[k], flt, s)
        ok = r.debug_hash(frames[k], s)[0] == want["hash"]
        assert np.abs(got[k * dh:(k + 1) * dh, :dw].astype(int) - want["out_u8"].astype(int))[ok].max() <= 1
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 512), (97, 777), (33, 300)])
def test_histeq_device_entries_respect_pitch_and_bounds(shape):
    lib = _cabi.load()
    h = ctypes.c_void_p()
    _cabi.check(lib.raisr_create(ctypes.byref(h), 0, 24, 3, 3, 11))
    hgt, w = shape
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    sfull, sview, spitch = _framed(hgt, w, torch.uint8, pad_left=7, pad_right=13)
    sview.copy_(torch.from_numpy(img).cuda())
    # LUT pass
    lut = rng.permutation(256).astype(np.uint8)
    dlut = torch.from_numpy(lut).cuda()
    full, view, dpitch = _framed(hgt, w, torch.uint8, pad_left=3, pad_right=29)
    torch.cuda.synchronize()
    _cabi.check(lib.ocv_histeq_global_u8(h, ctypes.c_void_p(sview.data_ptr()), w, hgt, spitch, ctypes.c_void_p(view.data_ptr()), dpitch,
                                         ctypes.c_void_p(dlut.data_ptr()), _cabi.RAISR_DEVICE, None))
    _cabi.check(lib.raisr_sync(h))
    assert _guards_intact(full, shape, torch.uint8, pad_left=3) and np.array_equal(view.cpu().numpy(), lut[img])
    # block blend
    bh, bw = 32, 128
    ny, nx = max(hgt // bh, 1), max(w // bw, 1)
    maps = (rng.random((ny, nx, 256)) * 280 - 10).astype(np.float32)
    dmaps = torch.from_numpy(maps).cuda()
    full, view, dpitch = _framed(hgt, w, torch.uint8, pad_left=8, pad_right=8)
    torch.cuda.synchronize()
    _cabi.check(lib.ocv_histeq_local_block_u8(h, ctypes.c_void_p(sview.data_ptr()), w, hgt, spitch, ctypes.c_void_p(view.data_ptr()), dpitch,
                                              ctypes.c_void_p(dmaps.data_ptr()), nx, ny, bw, bh, _cabi.RAISR_DEVICE, None))
    _cabi.check(lib.raisr_sync(h))
    assert _guards_intact(full, shape, torch.uint8, pad_left=8)
    assert np.array_equal(view.cpu().numpy(), ho.local_block_apply(img, maps, (bh, bw)))
    # tile histograms (only for images that hold at least one tile)
    if hgt >= 32 and w >= 256:
        ty, tx = hgt // 32, w // 256
        hfull = torch.full((ty * tx * 256 + 64,), 0xDEADBEEF - (1 << 32), dtype=torch.int32, device="cuda")
        out = hfull[32:32 + ty * tx * 256]
        torch.cuda.synchronize()
        _cabi.check(lib.ocv_hist_grid_u8(h, ctypes.c_void_p(sview.data_ptr()), w, hgt, spitch, ctypes.c_void_p(out.data_ptr()),
                                         _cabi.RAISR_DEVICE, None))
        _cabi.check(lib.raisr_sync(h))
        assert (hfull[:32] == 0xDEADBEEF - (1 << 32)).all() and (hfull[32 + ty * tx * 256:] == 0xDEADBEEF - (1 << 32)).all()
        assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(ty, tx, 256), ho.hist_grid(img))
    lib.raisr_destroy(h)
