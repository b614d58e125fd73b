"""The CUDA path against OUTPUTS OF THE REFERENCE'S OWN KERNEL.

tests/golden/ref_cl.npz (see tests/test_ref_pin.py) holds what /root/reference/super_resolution/raisr.cl writes when
it is executed on the CPU.  Here the product (through the C-ABI) is set to the same semantics -- `quirks="as_written"`,
fp32 taps -- and compared with those images directly:

  shipped kernel (bilinear only)   bilinear_only(): identical, bit for bit
  corrected kernel text (the three slips of raisr.cl:271,310,316 fixed in the reference's text), `half`=binary32
                                   the product's DEFAULT semantics: same rule
  full kernel text, `half`=binary32  upsample(): within 1 LSB, except pixels whose as-written hash rounding decides
                                   (tests/test_ref_pin.py: undecidable() -- the oracle is used as that classifier only,
                                   the pixels compared are the product's and the reference's); counted and bounded
"""
import os

import numpy as np
import pytest

from oracle import raisr_oracle as O
from oclcomputervision_b200 import ClRaisr, synth
from tests.test_ref_pin import undecidable, luma_tensor_result, interp_index, CUBIC, lenna_reference

pytestmark = pytest.mark.gpu

GRAY = ["noise_x2", "smooth_x2", "lenna_x2", "smooth_x3", "smooth_x4", "flat_x2", "step_x2"]
BGRA = ["bgra_smooth_x2", "bgra_noise_x2"]


@pytest.fixture(scope="module")
def ref(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_cl.npz"))


def make(gray_mode, s, flt, quirks="as_written"):
    r = ClRaisr(gray_mode, quirks=quirks, taps="fp32")
    setattr(r, "filters_x%d" % s, flt)
    return r


@pytest.mark.parametrize("name", GRAY)
def test_shipped_kernel_output_bit_exact(ref, name):
    src, s = ref[name + "_src"], int(ref[name + "_scale"])
    r = make(1, s, synth.random_filters(s, seed=int(ref[name + "_fseed"])))
    dst = np.zeros((src.shape[0] * s, src.shape[1] * s), np.uint8)
    r.bilinear_only(src, dst, s)
    assert np.array_equal(dst, ref[name + "_shipped_f32"])
    r.close()


@pytest.mark.parametrize("name", GRAY)
def test_full_kernel_text_gray(ref, name):
    src, s = ref[name + "_src"], int(ref[name + "_scale"])
    r = make(1, s, synth.random_filters(s, seed=int(ref[name + "_fseed"])))
    dst = np.zeros((src.shape[0] * s, src.shape[1] * s), np.uint8)
    r.upsample(src, dst, s)
    d = np.abs(dst.astype(np.int32) - ref[name + "_full_f32"].astype(np.int32))
    loose = undecidable(O.raisr_ref(src, None, s, quirks="as_written"), s)
    print("%s: %d of %d pixels differ, %d by more than 1 LSB, %d on a rounding-decided hash" %
          (name, int((d > 0).sum()), d.size, int((d > 1).sum()), int(loose.sum())))
    assert d[~loose].max() <= 1
    assert (d[~loose] > 0).mean() < 2e-3
    if name != "step_x2":                                         # (step_x2 is made of horizontal edges on purpose)
        assert (d > 1).sum() <= max(3, 1e-3 * d.size)
    r.close()


@pytest.mark.parametrize("name", GRAY + BGRA)
def test_corrected_kernel_text_is_the_default_semantics(ref, name):
    """The product's default (`quirks="intended"`) against the reference's text with its three slips corrected
    (oracle/build_ref.py: INTENDED_FIXES), `half` = binary32."""
    src, s = ref[name + "_src"], int(ref[name + "_scale"])
    gray = src.ndim == 2
    r = make(1 if gray else 0, s, synth.random_filters(s, seed=int(ref[name + "_fseed"])), quirks="intended")
    dst = np.zeros((src.shape[0] * s, src.shape[1] * s) + src.shape[2:], np.uint8)
    r.upsample(src, dst, s)
    d = np.abs(dst.astype(np.int32) - ref[name + "_intended_f32"].astype(np.int32))
    d = d if gray else d.max(-1)
    internals = O.raisr_ref(src, None, s, quirks="intended") if gray else luma_tensor_result(src, s, quirks="intended")
    loose = undecidable(internals, s, quirks="intended")
    print("%s (intended): %d of %d pixels differ, %d by more than 1 LSB, %d on a rounding-decided hash" %
          (name, int((d > 0).sum()), d.size, int((d > 1).sum()), int(loose.sum())))
    assert d[~loose].max() <= 1
    assert (d[~loose] > 0).mean() < (2e-3 if gray else 5e-3)
    if name != "step_x2":
        assert (d > 1).sum() <= max(3, 2e-3 * d.size)
    r.close()


@pytest.mark.parametrize("name", CUBIC)
def test_cubic_sample_as_stage_one(ref, name):
    """ClRaisr(upscaler="bicubic") against the reference's text with its own cubic_sample switched in for linear_sample."""
    src, s = ref[name + "_src"], int(ref[name + "_scale"])
    gray = src.ndim == 2
    r = ClRaisr(1 if gray else 0, taps="fp32", upscaler="bicubic")
    setattr(r, "filters_x%d" % s, synth.random_filters(s, seed=int(ref[name + "_fseed"])))
    dst = np.zeros((src.shape[0] * s, src.shape[1] * s) + src.shape[2:], np.uint8)
    r.upsample(src, dst, s)
    d = np.abs(dst.astype(np.int32) - ref[name + "_cubic_intended_f32"].astype(np.int32))
    d = d if gray else d.max(-1)
    internals = (O.raisr_ref(src, None, s, upscaler="bicubic") if gray
                 else luma_tensor_result(src, s, quirks="intended", upscaler="bicubic"))
    loose = undecidable(internals, s, quirks="intended")
    print("%s (cubic stage 1): %d of %d pixels differ, %d by more than 1 LSB" % (name, int((d > 0).sum()), d.size, int((d > 1).sum())))
    assert d[~loose].max() <= 1 and (d[~loose] > 0).mean() < 5e-3
    assert (d > 1).sum() <= max(3, 2e-3 * d.size)
    r.close()


@pytest.mark.parametrize("name", BGRA)
def test_full_kernel_text_colour(ref, name):
    src, s = ref[name + "_src"], int(ref[name + "_scale"])
    r = make(0, s, synth.random_filters(s, seed=int(ref[name + "_fseed"])))
    dst = np.zeros((src.shape[0] * s, src.shape[1] * s, 4), np.uint8)
    r.upsample(src, dst, s)
    d = np.abs(dst.astype(np.int32) - ref[name + "_full_f32"].astype(np.int32)).max(-1)
    loose = undecidable(luma_tensor_result(src, s), s)
    print("%s: %d of %d pixels differ, %d by more than 1 LSB" % (name, int((d > 0).sum()), d.size, int((d > 1).sum())))
    assert d[~loose].max() <= 1
    assert (d[~loose] > 0).mean() < 5e-3
    assert (d > 1).sum() <= max(3, 2e-3 * d.size)
    r.close()


def test_interpolation_kernels_bit_exact(golden_dir):
    """clUtility.bilinear / bilinear_lds / bicubic / bicubic_lds against what the reference's interpolation.cl writes
    (`half` = binary32): identical, BGRA and gray, integer and fractional ratios, reductions."""
    from oclcomputervision_b200.interpolation import clUtility
    z, cases = interp_index(golden_dir)
    u = clUtility()
    for key, name, method, hw in cases:
        src = z[name + "_src"]
        dst = np.zeros(hw if src.ndim == 2 else hw + (4,), np.uint8)
        getattr(u, method)(src, dst)
        assert np.array_equal(dst, z[key + "_f32"]), key
    print("%d interpolation cases identical to the reference kernels' output" % len(cases))
    u.close()


def test_hist_kernels_bit_exact(golden_dir):
    """clHistEq.histGrid / histeqGlobal / histeqLocalBlock against what the reference's hist.cl writes."""
    from oclcomputervision_b200.histeq import clHistEq
    z = np.load(os.path.join(golden_dir, "ref_cl_hist.npz"))
    img = z["img"]
    cl = clHistEq.getInstance()
    hist, _ = cl.histGrid(img)
    assert np.array_equal(hist, z["hist"])
    out, _ = cl.histeqGlobal(img, z["mapping"])
    assert np.array_equal(out, z["global"])
    for k in range(int(z["n_local"])):
        crop, bs = tuple(z["local%d_crop" % k]), tuple(int(v) for v in z["local%d_block" % k])
        sub = np.ascontiguousarray(img[:crop[0], :crop[1]])
        out, _ = cl.histeqLocalBlock(sub, z["local%d_maps" % k], bs)
        assert np.array_equal(out, z["local%d_out" % k]), (k, bs)


def test_config1_lenna_full_frame(golden_dir):
    """BASELINE.json configs[0] (512 x 512 lenna luma -> 1024 x 1024) against the reference kernel's output for the whole frame."""
    src, flt, _, ref_img = lenna_reference(golden_dir)
    dst = np.zeros((1024, 1024), np.uint8)
    r = make(1, 2, flt)
    r.bilinear_only(src, dst, 2)
    assert np.array_equal(dst, ref_img["shipped"])
    for kind, quirks in (("full", "as_written"), ("intended", "intended")):
        rr = make(1, 2, flt, quirks=quirks)
        rr.upsample(src, dst, 2)
        d = np.abs(dst.astype(np.int32) - ref_img[kind].astype(np.int32))
        loose = undecidable(O.raisr_ref(src, None, 2, quirks=quirks), 2, quirks=quirks)
        print("lenna 512x512 (%s): %d of %d pixels differ, %d by more than 1 LSB" % (kind, int((d > 0).sum()), d.size, int((d > 1).sum())))
        assert d[~loose].max() <= 1 and (d[~loose] > 0).mean() < 5e-5
        assert (d > 0).sum() < 2e-4 * d.size
        rr.close()
    r.close()
