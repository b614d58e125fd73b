"""CPU tests of the oracle (oracle/): two independent restatements, closed-form answers, fixtures.

The reference has no tests or golden vectors for this path; the pin against outputs
of its own kernel is tests/test_ref_pin.py, these tests pin the oracle by independent means.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import raisr_oracle as O
from oclcomputervision_b200 import synth

F32 = np.float32


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_gaussian_constants_match_reference_formula():
    # raisr.py:48-60,80-82: fspecial 9x9 sigma 2 -> (81,) float32; separable to float32 rounding
    g81 = O.reference_gaussian81().reshape(9, 9)
    g1 = O.gauss1d()
    assert g1.dtype == np.float32 and g1.shape == (9,)
    assert abs(g81[4, 4] - 0.0416828) < 1e-7 and abs(g81[0, 0] - 0.00076345) < 1e-8
    assert np.abs(np.outer(g1.astype(np.float64), g1.astype(np.float64)) - g81).max() < 5e-9
    # constants baked into raisr_oracle.c / raisr_prep.cuh
    baked = [float.fromhex(h) for h in ("0x1.a22092p-3", "0x1.70fefap-3", "0x1.fb36c8p-4", "0x1.0f7df8p-4", "0x1.c4b2eep-6")]
    assert [float(x) for x in g1[4:]] == baked


@pytest.mark.parametrize("shape,s", [((33, 47), 2), ((20, 31), 3), ((16, 16), 4), ((5, 3), 2), ((1, 9), 2), ((2, 2), 3)])
def test_numpy_and_c_restatements_agree(shape, s):
    src = synth.synthetic_frame(shape[0], shape[1], 3, sigma=1.5)
    F = synth.random_filters(s, seed=5)
    a = O.raisr_ref(src, F, s)
    b = O.raisr_ref_c(src, F, s)
    for k in ("Uext", "U", "L1", "coherence", "out_f32"):
        assert np.array_equal(a[k], b[k]), k
    # atan2 comes from numpy vs glibc: allow an ulp, buckets may differ only on a bin edge
    assert np.abs(a["angle"] - b["angle"]).max() < 1e-6
    diff = a["hash"] != b["hash"]
    if diff.any():
        assert (O.edge_distance(b)[diff] < 1e-5).all()
    else:
        assert np.array_equal(a["out_u8"], b["out_u8"])


def test_stage1_matches_independent_float64_align_corners_interpolation():
    # the check basic/interpolation.py:121-133 prints as PSNR, made an assertion
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (37, 29), dtype=np.uint8)
    for s in (2, 3):
        U = O.raisr_ref_c(src, None, s, want=("U",))["U"]
        sh, sw = src.shape
        xs = np.linspace(0, sw - 1, sw * s)
        ys = np.linspace(0, sh - 1, sh * s)
        p = src.astype(np.float64) / 255.0
        x0 = np.minimum(np.floor(xs).astype(int), sw - 2); u = xs - x0
        y0 = np.minimum(np.floor(ys).astype(int), sh - 2); v = ys - y0
        ref = ((1 - u)[None] * (1 - v)[:, None] * p[np.ix_(y0, x0)] + u[None] * (1 - v)[:, None] * p[np.ix_(y0, x0 + 1)]
               + (1 - u)[None] * v[:, None] * p[np.ix_(y0 + 1, x0)] + u[None] * v[:, None] * p[np.ix_(y0 + 1, x0 + 1)])
        assert np.abs(U - ref).max() < 5e-6   # fp32 coordinate map (raisr.cl:209) vs float64 linspace
        assert U[0, 0] == F32(src[0, 0]) / F32(255) and U[-1, -1] == F32(src[-1, -1]) / F32(255)


def test_extended_domain_is_edge_replication_to_rounding():
    src = synth.synthetic_frame(20, 24, 5, sigma=1.5)
    r = O.raisr_ref_c(src, None, 2, want=("U", "Uext"))
    pad = np.pad(r["U"], 5, mode="edge")
    assert np.abs(pad - r["Uext"]).max() < 1e-6   # raisr.cl:171 CLAMP_TO_EDGE on extrapolated coords


def _edge_image(kind, n=32):
    img = np.zeros((n, n), np.uint8)
    if kind == "flat":
        img[:] = 128
    elif kind == "vertical":
        img[:, n // 2:] = 200; img[:, : n // 2] = 50
    elif kind == "horizontal":
        img[n // 2:, :] = 200; img[: n // 2, :] = 50
    elif kind == "diag":
        yy, xx = np.mgrid[0:n, 0:n]
        img[:] = np.where(xx + yy >= n, 200, 50)
    return img


def test_closed_form_buckets():
    F = synth.random_filters(2)
    # all-zero image: exactly zero gradient -> theta = atan2(0,0) = 0, L1 = 0, coherence 0 -> bucket 0
    r = O.raisr_ref_c(np.zeros((32, 32), np.uint8), F, 2)
    assert (r["hash"] // 4 == 0).all() and (r["L1"] == 0).all() and (r["out_u8"] == 0).all()
    # flat grey: the bilinear blend of equal texels is only constant to an ulp (true of raisr.cl:60 as
    # well), so angle/coherence are rounding noise, but strength stays in bin 0 and the output is
    # grey x (sum of the taps of whatever bucket was hit)
    r = O.raisr_ref_c(_edge_image("flat"), F, 2)
    assert (((r["hash"] // 4) // 3) % 3 == 0).all() and r["L1"].max() < 1e-10
    expect = (128.0 / 255.0) * F.reshape(-1, 121).astype(np.float64).sum(-1)[r["hash"]]
    assert np.abs(r["out_f32"] - np.clip(expect, 0, 1)).max() < 1e-5
    # vertical edge: gy ~ 0 -> mb, md ~ 0 -> theta = atan2(+-eps, L1) = 0 or (after the +pi fold of
    # raisr.cl:285-286) pi: angle bin 0 or 23; coherence 1 and L1 large -> top bins
    r = O.raisr_ref_c(_edge_image("vertical"), F, 2)
    on_edge = r["L1"] > 1e-3
    assert on_edge.any()
    b = r["hash"][on_edge] // 4
    assert np.isin(b // 9, (0, 23)).all() and (b % 3 == 2).all() and ((b // 3) % 3 == 2).all()
    # horizontal edge: gx ~ 0 -> L1 = md, so theta = atan2(+-eps, 0) = pi/2 after the fold -> bin 12
    r = O.raisr_ref_c(_edge_image("horizontal"), F, 2)
    assert (r["hash"][r["L1"] > 1e-3] // 36 == 12).all()
    # 45-degree edge: gx = gy -> ma = mb = md -> theta = atan2(m, m) = pi/4, which is exactly the
    # edge between angle bins 5 and 6 (pi/4 * 24/pi = 6.0), so rounding decides between the two
    r = O.raisr_ref_c(_edge_image("diag"), F, 2)
    on_edge = r["L1"] > 1e-3
    assert abs(float(np.median(r["angle"][on_edge])) - np.pi / 4) < 1e-3
    assert np.isin(r["hash"][on_edge] // 36, (5, 6)).mean() > 0.95


def test_hash_layout_matches_filter_table_indexing():
    # raisr.cl:316-317 / raisr.py:78: filters[(((a*NS+s)*NC+c)*s^2+ptype)*121 + i*11+j]
    src = synth.synthetic_frame(24, 24, 11, sigma=2.0)
    F = synth.random_filters(2)
    r = O.raisr_ref_c(src, F, 2)
    y, x = 17, 30
    h = r["hash"][y, x]
    assert h % 4 == (y % 2) * 2 + x % 2
    patch = r["Uext"][y:y + 11, x:x + 11].astype(np.float64)
    taps = F.reshape(-1, 121)[h].reshape(11, 11).astype(np.float64)
    assert abs(float((patch * taps).sum()) - float(r["out_f32"][y, x])) < 1e-5 or r["out_f32"][y, x] in (0.0, 1.0)


def test_golden_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "lenna_x2.npz"))
    src = g["src"]
    assert sha(src) == str(g["sha_src"]) and src.shape == (512, 512)
    F = synth.random_filters(2)
    r = O.raisr_ref_c(src, F, 2)
    assert sha(r["U"]) == str(g["sha_U"])
    assert sha(r["hash"]) == str(g["sha_hash"])
    assert sha(r["out_u8"]) == str(g["sha_out_u8"])
    assert np.array_equal(np.bincount(r["hash"].ravel(), minlength=864), g["hash_hist"])
    assert (g["hash_hist"] > 0).sum() == 655
    assert sha(O.bilinear_u8_c(src, 2)) == str(g["bilinear_u8_sha"])
    c = np.load(os.path.join(golden_dir, "small_cases.npz"))
    for name in ("a_x2", "b_x3", "c_x2_ragged", "d_x4"):
        s = int(c[name + "_scale"])
        Fs = synth.random_filters(s, seed=int(c[name + "_fseed"]))
        rc = O.raisr_ref_c(c[name + "_src"], Fs, s)
        rn = O.raisr_ref(c[name + "_src"], Fs, s)
        for k in ("hash", "out_u8", "U", "L1", "coherence", "out_f32"):
            assert np.array_equal(rc[k], c[name + "_" + k]), (name, k)
        assert np.array_equal(rn["U"], c[name + "_U"]) and np.array_equal(rn["L1"], c[name + "_L1"])
        d = rn["hash"] != c[name + "_hash"]
        assert d.mean() < 1e-3


def test_synthetic_recipe_reaches_every_bucket():
    src = synth.synthetic_frame(540, 960, 1000)
    h = O.raisr_ref_c(src, None, 2, want=("hash",))["hash"]
    assert len(np.unique(h)) == 864
    strength = ((h // 4) // 3) % 3
    frac = np.bincount(strength.ravel(), minlength=3) / strength.size
    assert (frac > 0.15).all()


def test_oracle_quirks_and_taps_switches():
    """quirks="as_written" / taps="fp16" exist in both restatements and agree with each other."""
    from oclcomputervision_b200.synth import random_filters, synthetic_frame
    src = synthetic_frame(40, 56, seed=3)
    flt = random_filters(2, seed=1)
    base = O.raisr_ref(src, flt, 2)
    for quirks, taps in (("as_written", "fp32"), ("intended", "fp16"), ("as_written", "fp16")):
        a = O.raisr_ref(src, flt, 2, quirks=quirks, taps=taps)
        b = O.raisr_ref_c(src, flt, 2, quirks=quirks, taps=taps)
        same = a["hash"] == b["hash"]
        assert same.mean() > 0.999 and (same | (O.edge_distance(a, quirks=quirks) < 1e-5)).all()
        assert np.abs(a["out_f32"] - b["out_f32"])[same].max() <= 1e-6
        if quirks == "as_written":
            # strength never reaches the hash: bucket = (angle*3 + 0)*3 + coherence bin
            bucket = a["hash"] // 4
            assert ((bucket // 3) % 3 == 0).all()
            assert (a["hash"] != base["hash"]).mean() > 0.2
        else:
            assert np.array_equal(a["hash"], base["hash"])
            d = np.abs(a["out_f32"] - base["out_f32"]).max()
            assert 0 < d < 5e-3          # fp16 rounding of the taps: visible, small
    # fp16 taps == running with a table that was rounded beforehand
    q = flt.astype(np.float16).astype(np.float32)
    assert np.array_equal(O.raisr_ref(src, q, 2)["out_f32"], O.raisr_ref(src, flt, 2, taps="fp16")["out_f32"])


def test_oracle_bicubic_cheap_upscaler():
    """cubic_sample (raisr.cl:63-106) as stage 1: the two restatements agree bit for bit; a constant image stays
    constant (the Catmull-Rom weights sum to 1) and the result is clamped to [0, 1]."""
    from oclcomputervision_b200.synth import random_filters, synthetic_frame
    src = synthetic_frame(33, 47, seed=5)
    flt = random_filters(3, seed=2)
    a = O.raisr_ref(src, flt, 3, upscaler="bicubic")
    b = O.raisr_ref_c(src, flt, 3, upscaler="bicubic")
    assert np.array_equal(a["Uext"], b["Uext"]) and np.array_equal(a["L1"], b["L1"])
    same = a["hash"] == b["hash"]
    assert (same | (O.edge_distance(a) < 1e-5)).all() and np.abs(a["out_f32"] - b["out_f32"])[same].max() <= 1e-6
    assert a["Uext"].min() >= 0.0 and a["Uext"].max() <= 1.0
    flat = O.raisr_ref_c(np.full((9, 11), 200, np.uint8), None, 2, upscaler="bicubic", want=("U",))["U"]
    assert np.abs(flat - np.float32(200 / 255)).max() < 1e-6
    edge = np.zeros((8, 8), np.uint8); edge[:, 4:] = 255
    u_lin = O.raisr_ref_c(edge, None, 2, want=("U",))["U"]
    u_cub = O.raisr_ref_c(edge, None, 2, upscaler="bicubic", want=("U",))["U"]
    assert np.abs(u_lin - u_cub).max() > 0.02        # the cubic kernel sharpens the step


@pytest.mark.parametrize("quirks", ["intended", "as_written"])
def test_oracle_vs_literal_float64_restatement_of_the_kernel_loops(quirks):
    """Independent pin of stages 3-7: the kernel text taken literally -- CONV3x3 as a flipped convolution
    (raisr.cl:43-46,235-253), the NON-separable 9x9 loop with gaussian[j][i] (raisr.cl:258-276), the eigen / hash
    formulas (raisr.cl:278-317) and the 121-tap loop (raisr.cl:322-330) -- evaluated in float64 with scipy, against
    the separable fp32 oracle.  Buckets may only differ where the float64 value sits within 1e-4 of a bin edge."""
    from scipy.signal import convolve2d, correlate2d
    s = 2
    src = synth.synthetic_frame(48, 64, seed=9)
    flt = synth.random_filters(s, seed=3)
    res = O.raisr_ref(src, flt, s, quirks=quirks)
    U = res["Uext"].astype(np.float64)                       # stage 1 is pinned separately (float64 test above)
    sob_x = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float64)   # raisr.py:38-47
    sob_y = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], np.float64)
    gx = convolve2d(U, sob_x, mode="valid")                  # true convolution = kernel flipped on both axes
    gy = convolve2d(U, sob_y, mode="valid")
    m, n = 4.0, 4.0
    yy, xx = np.ogrid[-m:m + 1, -n:n + 1]
    G = np.exp(-(xx * xx + yy * yy) / 8.0)
    G /= G.sum()                                             # raisr.py:48-60 with shape 9, sigma 2
    Gt = G.T                                                 # the kernel indexes gaussian[j][i]
    win = lambda a: correlate2d(a, Gt, mode="valid")
    mb = win(gx * gy)
    ma = mb if quirks == "as_written" else win(gx * gx)
    md = win(gy * gy)
    T, D = ma + md, ma * md - mb * mb
    rad = np.maximum(T * T / 4 - D, 0)
    L1 = T / 2 + np.sqrt(rad)
    L2 = np.maximum(T / 2 - np.sqrt(rad), 0)
    theta = np.arctan2(mb, L1 - md)
    theta = np.where(theta < 0, theta + np.pi, theta)
    s1, s2 = np.sqrt(L1), np.sqrt(L2)
    coh = np.where(s1 + s2 != 0, (s1 - s2) / np.where(s1 + s2 != 0, s1 + s2, 1), 0)
    a = np.clip((theta / np.pi * 24).astype(int), 0, 23)
    sq, cq = O.DEFAULT_STRENGTH_Q, O.DEFAULT_COHERENCE_Q
    si = np.where(L1 < sq[0], 0, np.where(L1 < sq[1], 1, 2))
    cval = L1 if quirks == "as_written" else coh
    ci = np.where(cval < cq[0], 0, np.where(cval < cq[1], 1, 2))
    if quirks == "as_written":
        si = np.zeros_like(si)
    dh, dw = a.shape
    yy, xx = np.mgrid[0:dh, 0:dw]
    h64 = ((a * 3 + si) * 3 + ci) * 4 + (yy % 2) * 2 + (xx % 2)
    assert h64.shape == res["hash"].shape
    bad = h64 != res["hash"]
    fa = theta / np.pi * 24
    near = np.minimum(np.abs(fa - np.rint(fa)), np.min([np.abs(L1 - q) / q for q in (sq if quirks == "intended" else cq)], axis=0))
    if quirks == "intended":
        near = np.minimum(near, np.min([np.abs(coh - q) for q in cq], axis=0))
    # two populations may legitimately differ between fp32 and float64: pixels on a bin edge, and numerically flat
    # pixels (tensor below 1e-9, i.e. gradients under 0.01 LSB: the angle of rounding noise) / near-isotropic ones
    shaky = (near < 1e-4) | (L1 < 1e-9) | (coh < 0.02)
    assert (bad & ~shaky).sum() == 0, int((bad & ~shaky).sum())
    assert (bad & (L1 >= 1e-9)).mean() < 0.005
    # stage 7 literally: i outer, j inner over the 11x11 patch with the filter of the oracle's own bucket
    taps = flt.reshape(-1, 121)[res["hash"]]
    acc = np.zeros((dh, dw))
    for i in range(11):
        for j in range(11):
            acc += U[i:i + dh, j:j + dw] * taps[:, :, i * 11 + j]
    assert np.abs(np.clip(acc, 0, 1) - res["out_f32"]).max() < 5e-6
