"""Generates the committed fixtures under tests/golden/ from the C oracle (oracle/raisr_oracle.c).

Run in the build container (needs /root/reference/images/lenna.png, cv2):
    python oracle/make_golden.py
These fixtures are the ORACLE's outputs: they pin it -- and through it the CUDA path -- against regressions.
The fixtures made from the reference's own kernels (run on the CPU) are tests/golden/ref_cl*.npz, see
oracle/make_golden_ref_cl.py.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import raisr_oracle as O  # noqa: E402
from oclcomputervision_b200 import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    bgr = cv2.imread("/root/reference/images/lenna.png")
    luma = cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb)[:, :, 0].copy()  # same conversion as raisr.py:158-160
    F2 = synth.random_filters(2)
    r = O.raisr_ref_c(luma, F2, 2)
    np.savez_compressed(
        os.path.join(out, "lenna_x2.npz"), src=luma,
        sha_src=sha(luma), sha_U=sha(r["U"]), sha_hash=sha(r["hash"]), sha_out_u8=sha(r["out_u8"]),
        hash_hist=np.bincount(r["hash"].ravel(), minlength=864).astype(np.int32),
        crop_hash=r["hash"][448:512, 448:512], crop_out_u8=r["out_u8"][448:512, 448:512],
        crop_out_f32=r["out_f32"][448:512, 448:512],
        bilinear_u8_sha=sha(O.bilinear_u8_c(luma, 2)))
    cases = {}
    for name, (h, w, s, seed) in {"a_x2": (40, 56, 2, 7), "b_x3": (24, 32, 3, 8), "c_x2_ragged": (37, 53, 2, 9),
                                  "d_x4": (16, 24, 4, 10)}.items():
        src = synth.synthetic_frame(h, w, seed, sigma=2.0)
        src[:, : w // 3] = (src[:, : w // 3].astype(np.int32) * 3 // 4 + 20).astype(np.uint8)
        F = synth.random_filters(s, seed=100 + s)
        rr = O.raisr_ref_c(src, F, s)
        cases[name + "_src"] = src
        cases[name + "_scale"] = np.int32(s)
        cases[name + "_fseed"] = np.int32(100 + s)
        for k in ("hash", "angle", "L1", "coherence", "out_f32", "out_u8", "U"):
            cases[name + "_" + k] = rr[k]
    np.savez_compressed(os.path.join(out, "small_cases.npz"), **cases)
    print("wrote", os.listdir(out))


if __name__ == "__main__":
    main()
