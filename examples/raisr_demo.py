#!/usr/bin/env python
"""The reference's RAISR demo (/root/reference/super_resolution/raisr.py:137-186) on the B200 path
(SURVEY.md 8(f) row N3): upscale an LR image 2x, 20 iterations, report mean H2D + kernel + D2H ms and
PSNR of bilinear vs RAISR against the HR image.

    python examples/raisr_demo.py HR.png [LR.png] [--filters filter.p]

Differences from the reference script: gray mode only (the luma path, raisr.py:157-161,176-178; chroma is
upscaled bilinearly by OpenCV exactly as the reference does); without --filters a random-init table is
used (the pretrained filter.p needs a download); PSNR is computed locally (skimage is not needed).
"""
import argparse
import pickle
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oclcomputervision_b200 import ClRaisr, synth          # noqa: E402
from oclcomputervision_b200.interpolation import psnr      # noqa: E402


def main():
    import cv2
    ap = argparse.ArgumentParser()
    ap.add_argument("hr")
    ap.add_argument("lr", nargs="?")
    ap.add_argument("--filters", default=None, help="pickle of the (24,3,3,4,121) table (filter.p)")
    ap.add_argument("--loops", type=int, default=20)
    args = ap.parse_args()
    refHR = cv2.imread(args.hr)
    hHR, wHR = refHR.shape[0] // 2 * 2, refHR.shape[1] // 2 * 2
    refHR = refHR[:hHR, :wHR]
    bgr = cv2.imread(args.lr) if args.lr else cv2.resize(refHR, (wHR // 2, hHR // 2))   # raisr.py:146-147
    refUp = cv2.resize(bgr, (wHR, hHR), interpolation=cv2.INTER_LINEAR)                 # raisr.py:152
    if args.filters:
        with open(args.filters, "rb") as fp:
            F = pickle.load(fp).astype(np.float32)                                      # raisr.py:77-78
    else:
        F = synth.random_filters(2)
    raisr = ClRaisr(1, filters=F)
    ycrcb = cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb)                                       # raisr.py:158-161
    ycrcb_dst = cv2.resize(ycrcb, (wHR, hHR), interpolation=cv2.INTER_LINEAR)
    src = ycrcb[:, :, 0].copy()
    dst = np.zeros((hHR, wHR), dtype=src.dtype)
    total = [0.0, 0.0, 0.0]
    for _ in range(args.loops):                                                          # raisr.py:167-174
        el = raisr.upsample(src, dst, 2)
        total = [a + b for a, b in zip(total, el)]
    ycrcb_dst[:, :, 0] = dst                                                             # raisr.py:176-178
    out = cv2.cvtColor(ycrcb_dst, cv2.COLOR_YCrCb2BGR)
    cv2.imwrite("raisr-out.png", out)
    cv2.imwrite("raisr-ref-upsample.png", refUp)
    print("elapsed: {:.3f} + {:.3f} + {:.3f} ms".format(*[t / args.loops for t in total]))   # raisr.py:182
    print("PSNR: cubic {:.3f} raisr {:.3f}".format(psnr(refUp, refHR), psnr(out, refHR)))     # raisr.py:184-186


if __name__ == "__main__":
    main()
