"""GPU parity at the sizes BASELINE.json names (configs 3, 4-like, 5), colour-path hash classification, host arrays
with row strides, the tap-format switch, and SURVEY.md 8(f) N3 (filter.p loader + the reference's demo flow).

Every comparison goes through the C-ABI (ClRaisr -> ctypes -> libraisr_b200.so) against the C oracle; full frames,
with hash mismatches classified as excused (oracle value within 1e-5 of a bin edge) or unexcused.
"""
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import raisr_oracle as O
from oclcomputervision_b200 import ClRaisr, clUtility, synth, _cabi
from tests.conftest import ROOT
from tests.test_parity_gpu import check_against_oracle, TOL_F32, EDGE_EPS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def raisr():
    r = ClRaisr(1)
    r.filters_x2 = synth.random_filters(2)
    r.filters_x3 = synth.random_filters(3)
    yield r
    r.close()


def test_config3_full_frame_4k_to_8k(raisr):
    """BASELINE.json configs[2] frame size (north_star's target): 3840x2160 -> 7680x4320, every one of the 33 M
    output pixels against the C oracle."""
    src = synth.synthetic_frame(2160, 3840, 1000)
    res = check_against_oracle(raisr, src, 2, raisr.filters_x2, 1, label="config 3 (4K->8K): ")
    assert res["unexcused"] == 0 and res["err_fp32"] < TOL_F32


def test_config5_full_frame_720p(raisr):
    """BASELINE.json configs[4] frame size: 1280x720 -> 2560x1440."""
    src = synth.synthetic_frame(720, 1280, 1003)
    res = check_against_oracle(raisr, src, 2, raisr.filters_x2, 1, label="config 5 (720p->1440p): ")
    assert res["unexcused"] == 0 and res["err_fp32"] < TOL_F32


def test_x3_output_above_4096_squared(raisr):
    """The 3x path of config 4 (nine pixel types, 940 KB table) at 1376x1400 -> 4128x4200 in full against the oracle."""
    src = synth.synthetic_frame(1376, 1400, 1004)
    res = check_against_oracle(raisr, src, 3, raisr.filters_x3, 1, label="x3 4128x4200: ")
    assert res["unexcused"] == 0 and res["err_fp32"] < TOL_F32


@pytest.mark.parametrize("shape,s", [((270, 480), 2), ((96, 130), 3)])
def test_colour_hash_and_pixels_classified(shape, s):
    """Colour path (raisr.py:101-104): the hash derived from the Y plane equals the oracle's except on bin edges
    (classified with edge_distance like the gray path), and away from those pixels all four components are within
    1e-4 / 1 LSB -- no unclassified over-tolerance pixels."""
    rng = np.random.default_rng(shape[1])
    planes = [synth.synthetic_frame(shape[0], shape[1], 500 + k, sigma=2.0) for k in range(3)]
    alpha = rng.integers(200, 256, shape, dtype=np.uint8)
    src = np.stack(planes + [alpha], axis=2).copy()
    F = synth.random_filters(s)
    r = ClRaisr(0)
    setattr(r, "filters_x%d" % s, F)
    ref = O.raisr_ref_bgra_c(src, F, s)
    h, ang, l1, coh, _ = r.debug_hash(src, s)
    assert np.array_equal(l1, ref["L1"]) and np.array_equal(coh, ref["coherence"])
    assert np.abs(ang - ref["angle"]).max() < 2e-6
    diff = h != ref["hash"]
    excused = diff & (O.edge_distance(ref) < EDGE_EPS)
    unexcused = int((diff & ~excused).sum())
    out = r.upsample_f32(src, s)
    dst = np.zeros((shape[0] * s, shape[1] * s, 4), np.uint8)
    r.upsample(src, dst, s)
    err = np.abs(out - ref["out_f32"]).max(axis=2)
    lsb = np.abs(dst.astype(int) - ref["out_u8"].astype(int)).max(axis=2)
    print("colour %s x%d: hash mismatches %d excused, %d unexcused of %d; max err %.3g on equal-hash pixels" %
          (shape, s, int(excused.sum()), unexcused, h.size, float(err[~diff].max())))
    assert unexcused == 0
    assert err[~diff].max() < TOL_F32 and lsb[~diff].max() <= 1
    r.close()


def test_host_arrays_with_row_strides_are_not_overrun():
    """ClRaisr / clUtility accept row-strided numpy views (only the last stride must be 1).  The host<->device copies
    must move image bytes only: the bytes between the rows of a view keep their value, nothing is written past the
    parent array (the flat-copy bug would overwrite both), and the result equals the dense call."""
    s = 2
    flt = synth.random_filters(s, seed=2)
    r = ClRaisr(1, filters=flt, device=0)
    sh, sw, off = 60, 100, 24
    src_canvas = np.full((sh, sw + 2 * off), 7, np.uint8)
    dense_src = synth.synthetic_frame(sh, sw, seed=12)
    src_canvas[:, off:off + sw] = dense_src
    dst_canvas = np.full((sh * s + 1, sw * s + 2 * off), 201, np.uint8)   # one guard row below, guard columns beside
    dst_view = dst_canvas[:sh * s, off:off + sw * s]
    want = np.empty((sh * s, sw * s), np.uint8)
    r.upsample(dense_src, want, s)
    r.upsample(src_canvas[:, off:off + sw], dst_view, s)
    assert np.array_equal(dst_view, want)
    assert (dst_canvas[:, :off] == 201).all() and (dst_canvas[:, off + sw * s:] == 201).all() and (dst_canvas[-1] == 201).all()
    # bilinear_only and the float output through the same kind of view
    dst_canvas[:] = 201
    r.bilinear_only(src_canvas[:, off:off + sw], dst_view, s)
    assert np.array_equal(dst_view, O.bilinear_u8_c(dense_src, s))
    assert (dst_canvas[:, :off] == 201).all() and (dst_canvas[:, off + sw * s:] == 201).all() and (dst_canvas[-1] == 201).all()
    r.close()
    # colour mode
    c = ClRaisr(0, filters=flt, device=0)
    bsrc_canvas = np.full((sh, sw + 16, 4), 9, np.uint8)
    bgra = np.stack([synth.synthetic_frame(sh, sw, seed=20 + k) for k in range(4)], axis=2)
    bsrc_canvas[:, 8:8 + sw] = bgra
    bdst_canvas = np.full((sh * s + 1, sw * s + 16, 4), 55, np.uint8)
    bview = bdst_canvas[:sh * s, 8:8 + sw * s]
    bwant = np.empty((sh * s, sw * s, 4), np.uint8)
    c.upsample(np.ascontiguousarray(bgra), bwant, s)
    c.upsample(bsrc_canvas[:, 8:8 + sw], bview, s)
    assert np.array_equal(bview, bwant)
    assert (bdst_canvas[:, :8] == 55).all() and (bdst_canvas[:, 8 + sw * s:] == 55).all() and (bdst_canvas[-1] == 55).all()
    c.close()
    # stand-alone resizer: strided views and a dense gray destination whose width is not a multiple of 4
    u = clUtility()
    g = synth.synthetic_frame(50, 77, seed=3)
    for mode in ("bilinear_lds", "bicubic", "bilinear"):
        d61 = np.zeros((45, 61), np.uint8)
        getattr(u, mode)(g, d61)
        assert np.array_equal(d61, O.resize_u8_c(g, (45, 61), mode)), mode
        canvas = np.full((46, 61 + 10), 99, np.uint8)
        getattr(u, mode)(g, canvas[:45, 5:66])
        assert np.array_equal(canvas[:45, 5:66], d61)
        assert (canvas[:, :5] == 99).all() and (canvas[:, 66:] == 99).all() and (canvas[-1] == 99).all()
    u.close()


@pytest.mark.parametrize("s", [2, 3])
def test_tap_formats_fp32_b24_auto(s):
    """taps="auto" picks the 24-bit records for the bench table (bound <= 5e-5) and falls back to fp32 records for
    a table whose bound is larger; "b24" and "fp32" force either.  Outputs of b24 stay within the reported bound of
    the fp32-tap output, and the oracle run on the effective taps reproduces them to 5e-6."""
    src = synth.synthetic_frame(150, 210, seed=44)
    F = synth.random_filters(s)        # the bench table of this scale: its b24 bound is 4.4e-5 / 4.5e-5
    outs = {}
    for taps in ("fp32", "b24", "auto"):
        r = ClRaisr(1, device=0, taps=taps)
        setattr(r, "filters_x%d" % s, F)
        Feff, fmt, bound = r.effective_filters(s)
        assert fmt == ("fp32" if taps == "fp32" else "b24")
        assert (taps == "fp32") == np.array_equal(Feff, F)
        outs[taps] = (r.upsample_f32(src, s), bound, Feff)
        r.close()
    assert np.array_equal(outs["b24"][0], outs["auto"][0])
    bound = outs["b24"][1]
    assert 0 < bound <= 5e-5
    assert np.abs(outs["b24"][0] - outs["fp32"][0]).max() <= bound + 2e-6
    ref_eff = O.raisr_ref_c(src, outs["b24"][2], s, want=("out_f32", "hash"))
    assert np.abs(outs["b24"][0] - ref_eff["out_f32"]).max() < 5e-6
    # a table with large taps: the bound fails, auto keeps fp32 records
    big = (F * 40.0).astype(np.float32)
    r = ClRaisr(1, device=0)
    setattr(r, "filters_x%d" % s, big)
    _, fmt, bound_big = r.effective_filters(s)
    assert fmt == "fp32" and bound_big > 5e-5
    r.close()


def test_filter_pickle_loader_float64(tmp_path):
    """the upstream filter.p is a pickled float64 (24,3,3,4,121) array (raisr.py:77-78).  Loading
    it through filter_path= gives the same result as passing the float32 array through filters=."""
    F = synth.random_filters(2, seed=21)
    path = tmp_path / "filter.p"
    with open(path, "wb") as fp:
        pickle.dump(F.astype(np.float64), fp)
    src = synth.synthetic_frame(90, 120, seed=5)
    a = ClRaisr(1, filter_path=str(path))
    b = ClRaisr(1, filters=F)
    assert a.filters_x2.dtype == np.float32 and np.array_equal(a.filters_x2, F)
    da, db = np.zeros((180, 240), np.uint8), np.zeros((180, 240), np.uint8)
    a.upsample(src, da, 2)
    b.upsample(src, db, 2)
    assert np.array_equal(da, db)
    a.close(); b.close()
    with pytest.raises(FileNotFoundError):
        ClRaisr(1, filter_path=str(tmp_path / "missing.p"))
    # without any table the reference call prints and returns None (raisr.py:93-94 style)
    c = ClRaisr(1)
    assert c.upsample(src, da, 2) is None
    c.close()


@pytest.mark.parametrize("img_gray", [0, 1])
def test_reference_demo_flow_both_modes(tmp_path, img_gray):
    """examples/raisr_demo.py follows raisr.py:137-186 in colour (imgGray = 0, what the reference runs) and gray
    mode: 20 upsample calls, mean [h2d, kernel, d2h] ms, PSNR of bilinear and RAISR against the HR image."""
    import cv2
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import raisr_demo
    hr = np.stack([synth.synthetic_frame(192, 256, seed=70 + k, sigma=3.0) for k in range(3)], axis=2)
    hr_path = str(tmp_path / "hr.png")
    cv2.imwrite(hr_path, hr)
    F = synth.random_filters(2, seed=33)
    fpath = str(tmp_path / "filter.p")
    with open(fpath, "wb") as fp:
        pickle.dump(F.astype(np.float64), fp)
    res = raisr_demo.run_demo(hr_path, None, img_gray, fpath, loopcount=20, out_dir=str(tmp_path))
    assert len(res["elapsed"]) == 3 and all(t >= 0 for t in res["elapsed"])
    assert os.path.exists(tmp_path / "raisr-out.png") and os.path.exists(tmp_path / "raisr-ref-upsample.png")
    print("demo imgGray=%d: elapsed %.3f + %.3f + %.3f ms, PSNR cubic %.3f raisr %.3f" %
          ((img_gray,) + tuple(res["elapsed"]) + (res["psnr_cubic"], res["psnr_raisr"])))
    assert res["out"].shape[:2] == (192, 256) and res["out"].shape[2] == (3 if img_gray else 4)
    # identity-plus-noise filters (random-init, sigma 0.02 on 121 taps) keep the result in the neighbourhood of the bilinear one
    assert res["psnr_raisr"] > 30.0 and res["psnr_cubic"] > 30.0
    # the demo's RAISR image is what the class computes: re-run the core call and compare
    bgr = cv2.resize(hr, (128, 96))
    r = ClRaisr(img_gray, filters=F)
    if img_gray:
        ycrcb = cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb)
        dst = np.zeros((192, 256), np.uint8)
        r.upsample(ycrcb[:, :, 0].copy(), dst, 2)
        up = cv2.resize(ycrcb, (256, 192), interpolation=cv2.INTER_LINEAR)
        up[:, :, 0] = dst
        assert np.array_equal(cv2.cvtColor(up, cv2.COLOR_YCrCb2BGR), res["out"])
    else:
        dst = np.zeros((192, 256, 4), np.uint8)
        r.upsample(cv2.cvtColor(bgr, cv2.COLOR_BGR2BGRA), dst, 2)
        assert np.array_equal(dst, res["out"])
    r.close()


def test_device_then_host_calls_share_scratch_safely():
    """An un-synchronised DEVICE call on the caller's stream followed at once by a HOST call (which runs on the
    handle's own stream) reuse the same uext / hash scratch: the handle orders them."""
    import torch
    s = 2
    flt = synth.random_filters(s, seed=14)
    r = ClRaisr(1, filters=flt, device=0)
    frames = synth.synthetic_batch(6, 540, 960, pool=6, seed=300)
    small = synth.synthetic_frame(64, 96, seed=301)
    want_small = np.zeros((128, 192), np.uint8)
    r.upsample(small, want_small, s)
    want_big = np.zeros((6, 1080, 1920), np.uint8)
    r.upsample_batch(frames, want_big, s)
    st = torch.cuda.Stream()
    r.set_stream(st.cuda_stream)
    tsrc = torch.from_numpy(frames).cuda()
    tdst = torch.zeros((6, 1080, 1920), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        r.upsample_device(tsrc.data_ptr(), 960, 540, 960, tdst.data_ptr(), 1920, s, 6, np.uint8, timed=False)   # async
        got_small = np.zeros((128, 192), np.uint8)
        r.upsample(small, got_small, s)                                                                      # host call right behind it
        st.synchronize()
        assert np.array_equal(got_small, want_small)
        assert np.array_equal(tdst.cpu().numpy(), want_big)
        tdst.zero_()
    r.set_stream(0)
    r.close()


@pytest.mark.parametrize("shape", [(1, 1), (3, 200), (37, 53), (130, 67), (270, 480), (64, 65)])
def test_duo_kernel_equals_one_type_kernel(shape):
    """s = 2 with 24-bit records runs the two-types-per-CTA kernel (csrc/raisr_duo.cuh, `filter_duo` = 1, default); it
    keeps the per-pixel accumulation order of filter_octet_kernel, so u8 and float outputs are bit-identical to the
    one-type kernel (`filter_duo` = 0) -- single frames, ragged sizes and batches that span several tiles per worker."""
    flt = synth.random_filters(2, seed=17)
    n = 3
    frames = np.stack([synth.synthetic_frame(max(shape[0], 8), max(shape[1], 8), seed=400 + k)[:shape[0], :shape[1]] for k in range(n)]).copy()
    outs = []
    for duo in (1, 0):
        r = ClRaisr(1, filters=flt, device=0)
        assert r.effective_filters(2)[1] == "b24"
        r.set_option("filter_duo", duo)
        dst = np.empty((n, 2 * shape[0], 2 * shape[1]), np.uint8)
        r.upsample_batch(frames, dst, 2)
        dstf = np.empty((n, 2 * shape[0], 2 * shape[1]), np.float32)
        r.upsample_batch(frames, dstf, 2)
        outs.append((dst, dstf))
        r.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("s", [2, 3])
def test_near_flat_frames_tiny_tensors_and_exact_zeros(raisr, s):
    """Flat frames with isolated 1-LSB bumps and a 1-LSB step: the structure tensor runs from ~1e-5 down through values
    below 1e-10 to exact zeros -- the operands on which prep2_kernel's inlined sqrt / divide fast paths rely on their
    range guards and branch-free zero handling.  L1 / coherence stay bit-exact and the hash matches the oracle
    everywhere off the bin edges."""
    rng = np.random.default_rng(7)
    src = np.full((96, 160), 128, np.uint8)
    ys, xs = rng.integers(4, 92, 12), rng.integers(4, 156, 12)
    src[ys, xs] += 1
    src[:, 120:] += 1
    src[40:, :30] -= 1
    F = raisr.filters_x2 if s == 2 else raisr.filters_x3
    ref = O.raisr_ref_c(src, F, s)
    T = ref["L1"]      # L1 <= T <= 2 L1: a proxy for the trace
    assert ((T > 0) & (T < 5e-11)).any() and (T == 0).any() and (T > 1e-8).any()     # all three regimes are present
    res = check_against_oracle(raisr, src, s, F, 1, label="near-flat x%d: " % s)
    assert res["unexcused"] == 0


@pytest.mark.parametrize("shape", [(1, 1), (3, 200), (37, 53), (130, 67), (270, 480)])
def test_eigen_in_filter_equals_two_kernel_split(shape):
    """`eigen_in_filter` = 1: prep2_kernel stores the structure tensor and the s = 2 b24 filter kernel does the
    eigen-solve / hash itself (eigen_bucket, the scalar sequence that is bit-identical to the packed one).  Outputs
    are bit-identical to the default split, for single frames, ragged sizes and multi-chunk batches."""
    flt = synth.random_filters(2, seed=19)
    n = 4
    frames = np.stack([synth.synthetic_frame(max(shape[0], 8), max(shape[1], 8), seed=500 + k)[:shape[0], :shape[1]] for k in range(n)]).copy()
    outs = []
    for eig in (1, 0):
        r = ClRaisr(1, filters=flt, device=0)
        r.set_option("eigen_in_filter", eig)
        r.set_option("filter_duo", 0)
        dst = np.empty((n, 2 * shape[0], 2 * shape[1]), np.uint8)
        r.upsample_batch(frames, dst, 2)
        dstf = np.empty((n, 2 * shape[0], 2 * shape[1]), np.float32)
        r.upsample_batch(frames, dstf, 2)
        outs.append((dst, dstf))
        r.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("n_angle,n_strength,n_coherence,quirks", [(24, 3, 3, "as_written"), (8, 2, 2, "intended"), (28, 3, 3, "intended"),
                                                                   (16, 4, 4, "intended")])
def test_default_kernels_with_other_bucket_counts_and_quirks(n_angle, n_strength, n_coherence, quirks):
    """The default s = 2 path (24-bit records, two-types-per-CTA kernel) with bucket counts other than 24x3x3 --
    32 buckets; 252 buckets, where the two table slices no longer fit shared memory and the one-type kernel takes
    over; 4 x 4 quantisers (general-quantiser prep instantiation) -- and with the as-written quirks."""
    rng = np.random.default_rng(n_angle)
    F = rng.normal(0, 0.02, (n_angle, n_strength, n_coherence, 4, 121))
    F[..., 60] += 1
    F = np.ascontiguousarray((F / F.sum(-1, keepdims=True)).astype(np.float32))
    sq = [1e-4, 1e-3, 1e-2][:n_strength - 1]
    cq = [0.25, 0.5, 0.75][:n_coherence - 1]
    src = synth.synthetic_frame(140, 200, seed=77)
    r = ClRaisr(1, filters=F, device=0, n_angle=n_angle, n_strength=n_strength, n_coherence=n_coherence, quirks=quirks)
    r.set_quantizers(sq, cq)
    ref = O.raisr_ref_c(src, F, 2, n_angle=n_angle, n_strength=n_strength, n_coherence=n_coherence, strength_q=sq, coherence_q=cq, quirks=quirks)
    h = r.debug_hash(src, 2)[0]
    bad = h != ref["hash"]
    near = O.edge_distance(ref, n_angle=n_angle, strength_q=sq, coherence_q=cq, quirks=quirks) < EDGE_EPS
    assert not (bad & ~near).any()
    Feff, fmt, bound = r.effective_filters(2)
    out = r.upsample_f32(src, 2)
    assert np.abs(out - ref["out_f32"])[~bad].max() < TOL_F32
    if fmt == "b24":
        ref_eff = O.raisr_ref_c(src, Feff, 2, n_angle=n_angle, n_strength=n_strength, n_coherence=n_coherence, strength_q=sq, coherence_q=cq,
                                quirks=quirks, want=("out_f32",))
        assert np.abs(out - ref_eff["out_f32"])[~bad].max() < 5e-6
    dst = np.empty((280, 400), np.uint8)
    r.upsample(src, dst, 2)
    assert np.abs(dst.astype(int) - ref["out_u8"].astype(int))[~bad].max() <= 1
    r.close()
