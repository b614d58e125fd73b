"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on the same inputs.

Acceptance (BASELINE.json north_star / SURVEY.md 8(c)):
  * upscaled image U              bit-exact (every op is a correctly-rounded fp32 op in a fixed order)
  * hash                          identical except where the oracle's float angle*24/pi, L1 or coherence
                                  lies within 1e-5 of a bin edge; both counts are reported
  * out_f32                       within 1e-4 absolute
  * out_u8                        within 1 LSB
"""
import os

import numpy as np
import pytest

from oracle import raisr_oracle as O
from oclcomputervision_b200 import ClRaisr, synth, _cabi

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4
EDGE_EPS = 1e-5


@pytest.fixture(scope="module")
def raisr():
    r = ClRaisr(1)
    r.filters_x2 = synth.random_filters(2)
    r.filters_x3 = synth.random_filters(3)
    r.filters_x4 = synth.random_filters(4)
    yield r
    r.close()


TOL_EFF = 5e-6   # against the oracle run on the taps the kernel actually holds (summation order differs, nothing else)


def check_against_oracle(raisr, src, s, F, impl=None, label=""):
    """Full-frame comparison with the C oracle.  Returns dict(excused, unexcused, err_fp32, err_eff, tap_format, bound).
    Hash mismatches are classified by the oracle's distance to the nearest bin edge (< 1e-5 = excused); pixels are
    compared (i) with the fp32-tap oracle at the north_star tolerance of 1e-4 / 1 LSB and (ii) with the oracle run
    on `effective_filters()` -- the 24-bit taps when the table qualifies -- at 5e-6."""
    ref = O.raisr_ref_c(src, F, s)
    if impl is not None:
        raisr.set_option("filter_impl", impl)
    h, ang, l1, coh, U = raisr.debug_hash(src, s)
    assert np.array_equal(U, ref["U"]), "stage-1 upscale is not bit-exact"
    assert np.array_equal(l1, ref["L1"]) and np.array_equal(coh, ref["coherence"])
    assert np.abs(ang - ref["angle"]).max() < 2e-6
    diff = h != ref["hash"]
    excused = diff & (O.edge_distance(ref) < EDGE_EPS)
    unexcused = int((diff & ~excused).sum())
    print("%shash mismatches: %d excused (bin edge), %d unexcused, of %d" % (label, int(excused.sum()), unexcused, h.size))
    assert unexcused == 0
    out = raisr.upsample_f32(src, s)
    dst = np.zeros((src.shape[0] * s, src.shape[1] * s), np.uint8)
    ms = raisr.upsample(src, dst, s)
    assert len(ms) == 3 and all(m >= 0 for m in ms)
    ok = ~diff  # pixels that hashed to a different (edge) bucket legitimately use another filter
    err = np.abs(out - ref["out_f32"])
    assert err[ok].max() < TOL_F32, err[ok].max()
    lsb = np.abs(dst.astype(int) - ref["out_u8"].astype(int))
    assert lsb[ok].max() <= 1
    Feff, fmt, bound = raisr.effective_filters(s)
    err_eff = float(err[ok].max())
    if fmt != "fp32" and impl != 0:      # the block kernel (impl 0) always holds fp32 taps
        assert bound <= 5e-5 or raisr.taps != "auto"
        assert np.abs(Feff - F).max() <= np.abs(F).max() * 2.0 ** -16
        ref_eff = O.raisr_ref_c(src, Feff, s, want=("out_f32",))
        err_eff = float(np.abs(out - ref_eff["out_f32"])[ok].max())
        assert err_eff < TOL_EFF, err_eff
        assert err[ok].max() <= bound + TOL_EFF
    print("%smax |out - oracle(fp32 taps)| = %.3g, vs oracle(%s taps) = %.3g, b24 bound %.3g, u8 > 0 LSB on %d px" %
          (label, float(err[ok].max()), fmt, err_eff, bound, int((lsb[ok] > 0).sum())))
    return dict(excused=int(excused.sum()), unexcused=unexcused, err_fp32=float(err[ok].max()), err_eff=err_eff,
                tap_format=fmt, bound=bound)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("shape,s", [((40, 56), 2), ((37, 53), 2), ((128, 192), 2), ((24, 32), 3), ((50, 70), 3),
                                     ((16, 24), 4), ((1, 9), 2), ((9, 1), 2), ((2, 2), 3), ((130, 67), 2)])
def test_parity_small(raisr, shape, s, impl):
    src = synth.synthetic_frame(shape[0], shape[1], 21 + shape[0], sigma=2.0)
    F = {2: raisr.filters_x2, 3: raisr.filters_x3, 4: raisr.filters_x4}[s]
    check_against_oracle(raisr, src, s, F, impl)


def test_golden_small_cases(raisr, golden_dir):
    c = np.load(os.path.join(golden_dir, "small_cases.npz"))
    saved = {2: raisr.filters_x2, 3: raisr.filters_x3, 4: raisr.filters_x4}
    try:
        for name in ("a_x2", "b_x3", "c_x2_ragged", "d_x4"):
            s = int(c[name + "_scale"])
            F = synth.random_filters(s, seed=int(c[name + "_fseed"]))
            setattr(raisr, "filters_x%d" % s, F)
            src = c[name + "_src"]
            h, ang, l1, coh, U = raisr.debug_hash(src, s)
            assert np.array_equal(U, c[name + "_U"]) and np.array_equal(l1, c[name + "_L1"])
            ok = h == c[name + "_hash"]
            assert ok.mean() > 0.999
            out = raisr.upsample_f32(src, s)
            assert np.abs(out - c[name + "_out_f32"])[ok].max() < TOL_F32
            dst = np.zeros_like(c[name + "_out_u8"])
            raisr.upsample(src, dst, s)
            assert np.abs(dst.astype(int) - c[name + "_out_u8"].astype(int))[ok].max() <= 1
    finally:
        for s, F in saved.items():
            setattr(raisr, "filters_x%d" % s, F)


def test_lenna_config1(raisr, golden_dir):
    # BASELINE.json configs[0]: 512x512 luma of images/lenna.png, 2x, 24x3x3x4 buckets
    g = np.load(os.path.join(golden_dir, "lenna_x2.npz"))
    src = g["src"]
    err = check_against_oracle(raisr, src, 2, raisr.filters_x2, 1)["err_fp32"]
    h = raisr.debug_hash(src, 2)[0]
    hist = np.bincount(h.ravel(), minlength=864)
    assert np.abs(hist - g["hash_hist"]).sum() <= 40         # only bin-edge pixels may move (2 entries each)
    dst = np.zeros((1024, 1024), np.uint8)
    raisr.upsample(src, dst, 2)
    crop_ok = h[448:512, 448:512] == g["crop_hash"]
    assert np.abs(dst[448:512, 448:512].astype(int) - g["crop_out_u8"].astype(int))[crop_ok].max() <= 1
    print("lenna max |out-oracle| = %.3g" % err)


def test_shipped_kernel_behaviour_bilinear_only(raisr, golden_dir):
    # raisr.cl:219-230 returns after stage 1; interpolation.cl:17-71 is the same mapping
    src = synth.synthetic_frame(75, 99, 4, sigma=1.5)
    for s in (2, 3):
        dst = np.zeros((75 * s, 99 * s), np.uint8)
        raisr.bilinear_only(src, dst, s)
        assert np.array_equal(dst, O.bilinear_u8_c(src, s))


def test_batch_equals_single_frames(raisr):
    frames = synth.synthetic_batch(5, 72, 104, pool=5, seed=50)
    out = np.zeros((5, 144, 208), np.uint8)
    raisr.set_option("chunk_budget_bytes", 1 << 20)   # force several chunks through the pipeline
    try:
        ms = raisr.upsample_batch(frames, out, 2)
    finally:
        raisr.set_option("chunk_budget_bytes", 208 << 20)
    assert len(ms) == 3
    for k in range(5):
        one = np.zeros((144, 208), np.uint8)
        raisr.upsample(frames[k], one, 2)
        assert np.array_equal(one, out[k])
    outf = np.zeros((5, 144, 208), np.float32)
    raisr.upsample_batch(frames, outf, 2)
    assert np.abs(np.rint(outf * 255) - out).max() <= 1


def test_overlapped_pipeline_matches_serial(raisr):
    # prep of chunk c+1 co-resident with the filter of chunk c (two streams, two scratch sets)
    frames = synth.synthetic_batch(7, 200, 328, pool=7, seed=90)
    serial = np.zeros((7, 400, 656), np.uint8)
    raisr.set_option("chunk_budget_bytes", 2 << 20)
    raisr.set_option("taps", 0)      # the overlapped pipeline's single-buffered kernel holds fp32 records: compare like with like
    try:
        raisr.upsample_batch(frames, serial, 2)
        raisr.set_option("overlap", 1)
        over = np.zeros_like(serial)
        raisr.upsample_batch(frames, over, 2)
        import torch
        t_src = torch.from_numpy(frames).cuda()
        t_dst = torch.zeros((7, 400, 656), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        raisr.upsample_device(t_src.data_ptr(), 328, 200, 328, t_dst.data_ptr(), 656, 2, 7, np.uint8, timed=True)
        dev = t_dst.cpu().numpy()
    finally:
        raisr.set_option("overlap", 0)
        raisr.set_option("taps", 3)
        raisr.set_option("chunk_budget_bytes", 208 << 20)
    assert np.array_equal(serial, over)
    assert np.array_equal(serial, dev)


def test_full_size_config2_frame(raisr):
    # BASELINE.json configs[1] frame size: 1080p -> 4K, compared in full against the C oracle
    src = synth.synthetic_frame(1080, 1920, 1000)
    check_against_oracle(raisr, src, 2, raisr.filters_x2, 1)


def test_size_independent_properties_at_8k(raisr):
    # BASELINE.json configs[2] frame size 4K -> 8K.  (i) an all-zero frame stays zero and a frame of
    # 255 hits strength bin 0; (ii) every 64-row band of a frame, run through the row-band entry
    # point, reproduces the rows of the whole-frame result; (iii) batch == singles is covered above.
    import torch
    src = synth.synthetic_frame(2160, 3840, 77)
    dst = np.zeros((4320, 7680), np.uint8)
    raisr.upsample(src, dst, 2)
    z = np.zeros((64, 3840), np.uint8)
    dz = np.ones((128, 7680), np.uint8)
    raisr.upsample(z, dz, 2)
    assert (dz == 0).all()
    # down-sampling check: RAISR filters are identity+noise, so the result stays close to bilinear
    bl = np.zeros_like(dst)
    raisr.bilinear_only(src, bl, 2)
    assert np.abs(dst.astype(int) - bl.astype(int)).mean() < 8
    # row-band path on device pointers against the whole-frame result
    lib = _cabi.load()
    tsrc = torch.from_numpy(src).cuda()
    import ctypes
    for (row0, rows) in ((0, 256), (2048, 512), (4320 - 128, 128)):
        first, last = ctypes.c_int(), ctypes.c_int()
        _cabi.check(lib.raisr_band_src_rows(2160, 2, row0, rows, ctypes.byref(first), ctypes.byref(last)))
        win = tsrc[first.value:last.value + 1].contiguous()
        tout = torch.zeros((rows, 7680), dtype=torch.uint8, device="cuda")
        _cabi.check(lib.raisr_upsample_band_u8(raisr._h, win.data_ptr(), 3840, 2160, 3840, first.value,
                                               last.value - first.value + 1, tout.data_ptr(), 7680, row0, rows, 2))
        raisr.sync()
        assert np.array_equal(tout.cpu().numpy(), dst[row0:row0 + rows]), (row0, rows)


def test_error_behaviour(raisr):
    src = np.zeros((8, 8), np.uint8)
    # untrained scale: the reference prints and returns None (raisr.py:93-94)
    assert raisr.upsample(src, np.zeros((40, 40), np.uint8), 5) is None
    with pytest.raises(_cabi.RaisrError):
        raisr.upsample(src, np.zeros((17, 16), np.uint8), 2)      # dst is not scale x src
    with pytest.raises(ValueError):
        raisr.upsample(src.astype(np.float32), np.zeros((16, 16), np.uint8), 2)
    with pytest.raises(ValueError):
        raisr.filters_x2 = np.zeros((24, 3, 3, 4, 120), np.float32)
    with pytest.raises(ValueError):
        ClRaisr(2)


@pytest.mark.parametrize("shape,s", [((40, 56), 2), ((37, 53), 2), ((24, 32), 3), ((96, 130), 2)])
def test_colour_bgra_path_against_oracle(shape, s):
    # grayMode = 0 is what the reference's __main__ runs (raisr.py:139,163-164)
    rng = np.random.default_rng(shape[1])
    planes = [synth.synthetic_frame(shape[0], shape[1], 300 + k, sigma=2.0) for k in range(3)]
    alpha = rng.integers(200, 256, shape, dtype=np.uint8)
    src = np.stack(planes + [alpha], axis=2).copy()
    F = synth.random_filters(s)
    r = ClRaisr(0)
    setattr(r, "filters_x%d" % s, F)
    ref = O.raisr_ref_bgra_c(src, F, s)
    out = r.upsample_f32(src, s)
    dst = np.zeros((shape[0] * s, shape[1] * s, 4), np.uint8)
    ms = r.upsample(src, dst, s)
    assert len(ms) == 3
    err = np.abs(out - ref["out_f32"]).max(axis=2)
    bad = err >= TOL_F32
    # a pixel whose hash sits on a bin edge may legitimately use a neighbouring filter: count, bound
    print("colour: max err %.3g, pixels over tolerance %d of %d" % (err[~bad].max(), int(bad.sum()), bad.size))
    assert bad.mean() < 2e-4
    assert (np.abs(dst.astype(int) - ref["out_u8"].astype(int)).max(axis=2)[~bad] <= 1).all()
    r.close()


# ---------------------------------------------------------------- semantics switches of SURVEY.md 8(c)
@pytest.mark.gpu
@pytest.mark.parametrize("quirks,taps", [("as_written", "fp32"), ("intended", "fp16"), ("as_written", "fp16")])
@pytest.mark.parametrize("scale", [2, 3])
def test_quirks_and_fp16_taps_against_oracle(quirks, taps, scale):
    """`quirks="as_written"` (raisr.cl:271,310,316) and `taps="fp16"` (raisr.cl:328) follow the oracle's
    switches of the same name: hash identical away from bin edges, pixels within 1e-4 / 1 LSB."""
    from oclcomputervision_b200.synth import random_filters, synthetic_frame
    ro = O
    src = synthetic_frame(120, 168, seed=31)
    flt = random_filters(scale, seed=5)
    want = ro.raisr_ref_c(src, flt, scale, quirks=quirks, taps=taps)
    base = ro.raisr_ref_c(src, flt, scale)
    assert (want["hash"] != base["hash"]).mean() > 0.2 if quirks == "as_written" else True
    assert not np.array_equal(want["out_f32"], base["out_f32"])          # the switch does something
    r = ClRaisr(1, device=0, quirks=quirks, taps=taps)
    setattr(r, "filters_x%d" % scale, flt)
    got_hash = r.debug_hash(src, scale)[0]
    bad = got_hash != want["hash"]
    near = ro.edge_distance(want, quirks=quirks) < 1e-5
    assert not (bad & ~near).any(), "%d unexcused hash mismatches" % int((bad & ~near).sum())
    out_f32 = r.upsample_f32(src, scale)
    ok = ~bad
    assert np.abs(out_f32 - want["out_f32"])[ok].max() <= 1e-4
    dst = np.empty((src.shape[0] * scale, src.shape[1] * scale), np.uint8)
    r.upsample(src, dst, scale)
    assert np.abs(dst.astype(int) - want["out_u8"].astype(int))[ok].max() <= 1
    # switching the tap precision back re-packs the table
    r.set_option("taps", 0)
    r.set_option("quirks", 0)
    assert np.abs(r.upsample_f32(src, scale) - base["out_f32"])[r.debug_hash(src, scale)[0] == base["hash"]].max() <= 1e-4


# ---------------------------------------------------------------- degenerate shapes
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (5, 1), (2, 2), (3, 200), (200, 3), (1, 300), (63, 65)])
@pytest.mark.parametrize("s", [2, 3, 4])
def test_tiny_and_sliver_shapes(shape, s):
    """Sources of one row / one column / one pixel: every clamp of the bilinear map, the tile halo and the
    TMA box is exercised at once.  Both prep kernels must agree with the oracle and with each other."""
    src = synth.synthetic_frame(max(shape[0], 8), max(shape[1], 8), seed=17)[:shape[0], :shape[1]].copy()
    flt = synth.random_filters(s, seed=4)
    want = O.raisr_ref_c(src, flt, s)
    outs = []
    for prep_impl in (2, 1):
        r = ClRaisr(1, device=0)
        setattr(r, "filters_x%d" % s, flt)
        r.set_option("prep_impl", prep_impl)
        h, ang, l1, coh, u = r.debug_hash(src, s)
        assert np.array_equal(u, want["U"]) and np.array_equal(l1, want["L1"]) and np.array_equal(coh, want["coherence"])
        bad = h != want["hash"]
        assert not (bad & ~(O.edge_distance(want) < 1e-5)).any()
        out = r.upsample_f32(src, s)
        assert np.abs(out - want["out_f32"])[~bad].max() <= 1e-4
        dst = np.empty((shape[0] * s, shape[1] * s), np.uint8)
        r.upsample(src, dst, s)
        assert np.abs(dst.astype(int) - want["out_u8"].astype(int))[~bad].max() <= 1
        outs.append((h, out, dst))
        r.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.gpu
def test_pinned_caller_arrays_same_result_and_released():
    """`ClRaisr.pin` page-locks caller-owned arrays for a reference-style loop; results are unchanged and the lock
    goes away with the array (or on unpin)."""
    import gc
    s = 2
    src = synth.synthetic_frame(270, 480, seed=8)
    flt = synth.random_filters(s, seed=6)
    r = ClRaisr(1, filters=flt, device=0)
    dst0 = np.empty((540, 960), np.uint8)
    r.upsample(src, dst0, s)
    psrc, pdst = ClRaisr.pin(src.copy()), ClRaisr.pin(np.empty((540, 960), np.uint8))
    assert ClRaisr.pin(psrc) is psrc                      # idempotent
    for _ in range(3):
        ms = r.upsample(psrc, pdst, s)
    assert np.array_equal(pdst, dst0) and len(ms) == 3
    key = pdst.ctypes.data
    assert key in ClRaisr._pinned
    ClRaisr.unpin(psrc)
    assert psrc.ctypes.data not in ClRaisr._pinned
    del pdst
    gc.collect()
    assert key not in ClRaisr._pinned
    with pytest.raises(ValueError):
        ClRaisr.pin(np.empty((4, 4), np.uint8)[:, ::2])
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(37, 53), (96, 130), (200, 64)])
def test_colour_two_plane_filter_equals_per_plane_launches(shape):
    """The two-planes-per-CTA colour kernel (raisr_octet2.cuh) keeps the per-plane accumulation order:
    its BGRA output is bit-identical to four single-plane launches."""
    rng = np.random.default_rng(3)
    bgra = rng.integers(0, 256, shape + (4,), dtype=np.uint8)
    for c in range(3):
        bgra[..., c] = synth.synthetic_frame(shape[0], shape[1], seed=40 + c)
    flt = synth.random_filters(2, seed=12)
    outs = []
    for impl in (2, 1):
        r = ClRaisr(0, filters=flt, device=0)
        r.set_option("color_filter_impl", impl)
        dst = np.empty((shape[0] * 2, shape[1] * 2, 4), np.uint8)
        r.upsample(bgra, dst, 2)
        outs.append((dst, r.upsample_f32(bgra, 2)))
        r.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,s", [((40, 56), 2), ((37, 53), 2), ((24, 33), 3), ((19, 21), 4), ((130, 150), 2)])
def test_bicubic_cheap_upscaler_against_oracle(shape, s):
    """Stage 1 = the reference's (unused) cubic_sample, raisr.cl:63-106: U bit-exact, the rest to the usual bar."""
    src = synth.synthetic_frame(shape[0], shape[1], seed=23)
    flt = synth.random_filters(s, seed=8)
    want = O.raisr_ref_c(src, flt, s, upscaler="bicubic")
    assert not np.array_equal(want["U"], O.raisr_ref_c(src, flt, s)["U"])
    r = ClRaisr(1, device=0, upscaler="bicubic")
    setattr(r, "filters_x%d" % s, flt)
    h, ang, l1, coh, u = r.debug_hash(src, s)
    assert np.array_equal(u, want["U"]) and np.array_equal(l1, want["L1"]) and np.array_equal(coh, want["coherence"])
    bad = h != want["hash"]
    assert not (bad & ~(O.edge_distance(want) < 1e-5)).any()
    assert np.abs(r.upsample_f32(src, s) - want["out_f32"])[~bad].max() <= 1e-4
    dst = np.empty((shape[0] * s, shape[1] * s), np.uint8)
    r.upsample(src, dst, s)
    assert np.abs(dst.astype(int) - want["out_u8"].astype(int))[~bad].max() <= 1
    r.set_option("prep_impl", 1)
    with pytest.raises(_cabi.RaisrError):
        r.upsample(src, dst, s)
    r.close()
    # colour mode: cubic_sample is a half4 routine in the reference (raisr.cl:63-106), all four channels go through it
    rng = np.random.default_rng(shape[0])
    bgra = np.stack([synth.synthetic_frame(shape[0], shape[1], seed=60 + k) for k in range(3)] +
                    [rng.integers(180, 256, shape, dtype=np.uint8)], axis=2).copy()
    cw = O.raisr_ref_bgra_c(bgra, flt, s, upscaler="bicubic")
    assert not np.array_equal(cw["out_u8"], O.raisr_ref_bgra_c(bgra, flt, s)["out_u8"])
    c = ClRaisr(0, device=0, upscaler="bicubic")
    setattr(c, "filters_x%d" % s, flt)
    ch, cang, cl1, ccoh, _ = c.debug_hash(bgra, s)
    assert np.array_equal(cl1, cw["L1"]) and np.array_equal(ccoh, cw["coherence"])
    cbad = ch != cw["hash"]
    assert not (cbad & ~(O.edge_distance(cw) < 1e-5)).any()
    assert np.abs(c.upsample_f32(bgra, s) - cw["out_f32"]).max(axis=2)[~cbad].max() <= 1e-4
    cdst = np.empty((shape[0] * s, shape[1] * s, 4), np.uint8)
    c.upsample(bgra, cdst, s)
    assert np.abs(cdst.astype(int) - cw["out_u8"].astype(int)).max(axis=2)[~cbad].max() <= 1
    c.close()


@pytest.mark.gpu
def test_host_pipeline_ramped_chunks_match_single_frames():
    """The HOST pipeline runs short chunks first and last (1, 2, chunk..., 2, 1); with a tiny scratch budget a
    20-frame batch goes through that schedule and must equal frame-by-frame calls."""
    s, n = 2, 20
    frames = np.stack([synth.synthetic_frame(120, 200, seed=60 + k) for k in range(n)])
    flt = synth.random_filters(s, seed=3)
    r = ClRaisr(1, filters=flt, device=0)
    per_frame = (240 + 18 + 3) // 4 * 4 * (400 + 10) * 4
    r.set_option("chunk_budget_bytes", max(4 * per_frame + 64, 1 << 20))
    dst = np.empty((n, 240, 400), np.uint8)
    r.upsample_batch(frames, dst, s)
    one = np.empty((240, 400), np.uint8)
    for k in range(n):
        r.upsample(frames[k], one, s)
        assert np.array_equal(dst[k], one), k
    r.close()


@pytest.mark.gpu
def test_filter_pipe_option_is_result_neutral():
    """`filter_pipe` (barrier-free tile pipeline) and `prep_impl` only change scheduling: outputs are bit-identical."""
    s = 2
    frames = np.stack([synth.synthetic_frame(300, 420, seed=70 + k) for k in range(3)])
    flt = synth.random_filters(s, seed=1)
    outs = []
    for pipe, prep in ((1, 2), (0, 2), (1, 1), (0, 1)):
        r = ClRaisr(1, filters=flt, device=0)
        r.set_option("filter_pipe", pipe)
        r.set_option("prep_impl", prep)
        dst = np.empty((3, 600, 840), np.uint8)
        r.upsample_batch(frames, dst, s)
        outs.append(dst)
        r.close()
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


@pytest.mark.gpu
def test_colour_batch_equals_single_frames():
    """`upsample_batch` in colour mode pipelines BGRA frames through the host path; each frame equals a single call."""
    rng = np.random.default_rng(5)
    n, sh, sw = 5, 72, 100
    batch = rng.integers(0, 256, (n, sh, sw, 4), dtype=np.uint8)
    for k in range(n):
        for c in range(3):
            batch[k, :, :, c] = synth.synthetic_frame(sh, sw, seed=90 + 3 * k + c)
    r = ClRaisr(0, filters=synth.random_filters(2, seed=7), device=0)
    dst = np.empty((n, 2 * sh, 2 * sw, 4), np.uint8)
    ms = r.upsample_batch(batch, dst, 2)
    assert len(ms) == 3
    one = np.empty((2 * sh, 2 * sw, 4), np.uint8)
    for k in range(n):
        r.upsample(batch[k], one, 2)
        assert np.array_equal(dst[k], one), k
    dstf = np.empty((n, 2 * sh, 2 * sw, 4), np.float32)
    r.upsample_batch(batch, dstf, 2)
    assert np.abs(dstf * 255 - dst).max() <= 0.5 + 1e-3
    with pytest.raises(ValueError):
        r.upsample_batch(batch[..., 0], dst[..., 0], 2)
    r.close()
