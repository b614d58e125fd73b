"""Generates tests/golden/ref_cl.npz, ref_cl_interp.npz and ref_cl_hist.npz: OUTPUTS OF THE REFERENCE'S OWN KERNELS
(raisr.cl, interpolation.cl, hist.cl), run here on the CPU.

    python oracle/make_golden_ref_cl.py          # build container only: needs /root/reference

oracle/build_ref.py compiles the kernel source where it lies against oracle/ref_shim/cl_shim.hpp; this script runs it
on small inputs the way ClRaisr.upsample does (oracle/raisr_cl_ref.py) and stores, per case, the source, the scale, the
seed of the filter table (oclcomputervision_b200.synth.random_filters) and four destination images:

    shipped_f16 / shipped_f32   the kernel as shipped (early return: bilinear only), `half` = binary16 / binary32
    full_f16 / full_f32         the early return compiled out: the whole RAISR text runs
    intended_f16 / intended_f32 as full, with the three slips of raisr.cl:271,310,316 corrected in the text (one token each,
                                oracle/build_ref.py: INTENDED_FIXES) -- the semantics the product defaults to
    cubic_intended_f32          (five cases) as intended_f32 with stage 1 calling the file's cubic_sample (CUBIC_SWITCH)

ref_cl_interp.npz holds, for BGRA and gray sources and several destination sizes (integer and fractional ratios; a
reduction for the two kernels that have no 20 x 20 local tile), what bilinear_simple / bilinear_lds / bicubic_simple /
bicubic_lds write, again with `half` as binary16 and as binary32.

These are the vectors that pin the oracle (tests/test_ref_pin.py) and, on the GPU box, the CUDA path directly
(tests/test_ref_pin_gpu.py); /root/reference does not travel, the fixture does.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import raisr_cl_ref as R  # noqa: E402
from oclcomputervision_b200 import synth  # noqa: E402

CUBIC_CASES = ("noise_x2", "smooth_x2", "lenna_x2", "smooth_x3", "bgra_noise_x2")   # + stage 1 = the file's own cubic_sample
VARIANTS = [(k, p) for k in ("shipped", "full", "intended") for p in ("f16", "f32")]


def sources():
    rng = np.random.default_rng(20260101)
    lenna = np.load(os.path.join(ROOT, "tests", "golden", "lenna_x2.npz"))["src"]     # luma of /root/reference/images/lenna.png
    step = np.full((24, 32), 30, np.uint8)
    step[:, 16:] = 230
    step[12:, :] = 255 - step[12:, :]
    smooth3 = [synth.synthetic_frame(48, 64, seed=s) for s in (5, 6, 7)]
    return {
        "noise_x2": (rng.integers(0, 256, (40, 48), dtype=np.uint8), 2),
        "smooth_x2": (synth.synthetic_frame(64, 80, seed=3), 2),
        "lenna_x2": (np.ascontiguousarray(lenna[192:320, 192:320]), 2),
        "smooth_x3": (synth.synthetic_frame(32, 48, seed=4), 3),
        "smooth_x4": (synth.synthetic_frame(24, 28, seed=11), 4),
        "flat_x2": (np.full((16, 24), 77, np.uint8), 2),
        "step_x2": (step, 2),
        "bgra_smooth_x2": (np.stack(smooth3 + [rng.integers(0, 256, (48, 64), dtype=np.uint8)], -1), 2),
        "bgra_noise_x2": (rng.integers(0, 256, (32, 48, 4), dtype=np.uint8), 2),
    }


def interp_cases():
    rng = np.random.default_rng(20260102)
    lenna = np.load(os.path.join(ROOT, "tests", "golden", "lenna_x2.npz"))["src"]
    smooth = [synth.synthetic_frame(24, 40, seed=s, sigma=2.0) for s in (21, 22, 23)]
    srcs = {
        "bgra_noise": rng.integers(0, 256, (24, 40, 4), dtype=np.uint8),
        "bgra_smooth": np.stack(smooth + [np.full((24, 40), 255, np.uint8)], -1),
        "gray_lenna": np.ascontiguousarray(lenna[200:248, 216:280]),
    }
    for name, src in srcs.items():
        h, w = src.shape[:2]
        for dh, dw in ((2 * h, 2 * w), (64, 96), (3 * h + 8, 4 * w)):      # x2; fractional ratios; anisotropic
            if dh >= h and dh % 16 == 0 and dw % 16 == 0:
                for method in ("bilinear", "bilinear_lds", "bicubic", "bicubic_lds"):
                    yield name, src, (dh, dw), method
        for method in ("bilinear", "bicubic"):                              # odd sizes and a reduction: no work-group constraint
            yield name, src, (h + 7, 2 * w - 3), method
            yield name, src, (h - 5, w - 9), method


def main_interp():
    out = {}
    index = []
    for name, src, hw, method in interp_cases():
        key = "%s_%s_%dx%d" % (name, method, hw[0], hw[1])
        out[name + "_src"] = src
        for prec in ("f16", "f32"):
            out[key + "_" + prec] = R.interp(src, hw, method, prec=prec)
        index.append(key)
    out["index"] = np.array(index)
    path = os.path.join(ROOT, "tests", "golden", "ref_cl_interp.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(index), "cases")


def main_hist():
    """histeq/hist.cl: tile histograms, the global LUT pass, the block-bilinear LUT blend (power-of-two and other block sizes;
    every image stays inside nblocks * block + block / 2, beyond which the reference indexes past its tables)."""
    from oracle import histeq_oracle as HO
    rng = np.random.default_rng(20260103)
    yy, xx = np.mgrid[0:256, 0:768].astype(np.float64)
    light = 0.2 + 0.6 * (0.5 + 0.5 * np.sin(xx / 120.0) * np.cos(yy / 70.0))
    img = np.clip(light * 255 + rng.normal(0, 18, light.shape), 0, 255).astype(np.uint8)
    img[:40, :300] = 9                                                   # a flat region (one bin takes a whole tile row)
    out = {"img": img, "hist": R.hist_grid(img)}
    mapping = HO.transfer_func(out["hist"].sum(axis=(0, 1)), 1, 0.05, 2).astype(np.uint8)
    out["mapping"] = mapping
    out["global"] = R.histeq_global(img, mapping)
    cases = []
    for k, (crop, bs) in enumerate((((256, 768), (64, 256)), ((256, 768), (256, 256)), ((128, 512), (32, 256)), ((112, 480), (48, 96)),
                                    ((240, 720), (80, 144)))):
        sub = np.ascontiguousarray(img[:crop[0], :crop[1]])
        maps = HO.block_mappings(sub, 0.5, 0.05, 3, bs)
        out["local%d_crop" % k] = np.array(crop)
        out["local%d_block" % k] = np.array(bs)
        out["local%d_maps" % k] = maps
        out["local%d_out" % k] = R.histeq_local_block(sub, maps, bs)
        cases.append(k)
    out["n_local"] = np.int32(len(cases))
    path = os.path.join(ROOT, "tests", "golden", "ref_cl_hist.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main_lenna():
    """BASELINE.json configs[0] in full: the 512 x 512 luma of the reference's images/lenna.png -> 1024 x 1024.  The three
    binary32 outputs of the reference kernel are stored as (sha256, sparse difference against the oracle's output): the test
    rebuilds the reference image from the oracle's and checks the hash, so the megabyte images need not be committed."""
    import hashlib
    from oracle import raisr_oracle as O
    src = np.load(os.path.join(ROOT, "tests", "golden", "lenna_x2.npz"))["src"]
    flt = synth.random_filters(2, seed=102)
    out = {"fseed": np.int32(102)}
    base = {"shipped": O.bilinear_u8_c(src, 2), "full": O.raisr_ref_c(src, flt, 2, quirks="as_written")["out_u8"],
            "intended": O.raisr_ref_c(src, flt, 2)["out_u8"]}
    for kind in ("shipped", "full", "intended"):
        got = R.run(src, flt, 2, kind=kind, prec="f32")
        idx = np.flatnonzero(got != base[kind])
        out[kind + "_sha256"] = np.array(hashlib.sha256(got.tobytes()).hexdigest())
        out[kind + "_diff_idx"] = idx.astype(np.int64)
        out[kind + "_diff_val"] = got.ravel()[idx]
        print("lenna 512x512", kind, "differs from the oracle at", idx.size, "of", got.size, "pixels")
    path = os.path.join(ROOT, "tests", "golden", "ref_cl_lenna.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    out = {}
    for name, (src, s) in sources().items():
        fseed = 100 + s
        flt = synth.random_filters(s, seed=fseed)
        out[name + "_src"] = src
        out[name + "_scale"] = np.int32(s)
        out[name + "_fseed"] = np.int32(fseed)
        for kind, prec in VARIANTS:
            out["%s_%s_%s" % (name, kind, prec)] = R.run(src, flt, s, kind=kind, prec=prec)
        if name in CUBIC_CASES:
            out[name + "_cubic_intended_f32"] = R.run(src, flt, s, kind="cubic_intended", prec="f32")
        print(name, src.shape, "x%d" % s)
    path = os.path.join(ROOT, "tests", "golden", "ref_cl.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main_interp()
    main_hist()
    main_lenna()
