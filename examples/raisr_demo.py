#!/usr/bin/env python
"""The reference's RAISR demo (/root/reference/super_resolution/raisr.py:137-186) on the B200 path
(SURVEY.md 8(f) row N3): upscale an LR image 2x, 20 iterations, report mean H2D + kernel + D2H ms and
PSNR of bilinear vs RAISR against the HR image.

    python examples/raisr_demo.py HR.png [LR.png] [--gray 0|1] [--filters filter.p]

Same flow as the reference script in both of its modes:
  --gray 0 (default, what the reference runs: `imgGray = 0`, raisr.py:139): BGR -> BGRA, `ClRaisr(0)`,
           the kernel does the RGB<->YUV conversions itself (raisr.py:163-164, raisr.cl:211-214,333-336);
  --gray 1: luma through `ClRaisr(1)`, chroma upscaled bilinearly by OpenCV (raisr.py:157-161,176-178).
Differences: without an LR image the HR image is halved by cv2.resize (the commented-out line raisr.py:146);
the table comes from --filters (the upstream float64 `filter.p` pickle, raisr.py:77-78 -- a pickle runs code
when loaded, only pass files you trust) or is random-init (the pretrained weights need a download);
PSNR is computed locally (skimage is not installed here).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oclcomputervision_b200 import ClRaisr, synth          # noqa: E402
from oclcomputervision_b200.interpolation import psnr      # noqa: E402


def run_demo(hr_path, lr_path=None, img_gray=0, filter_path=None, loopcount=20, out_dir=".", device=0, write=True):
    """Returns dict(elapsed=[h2d, kernel, d2h] mean ms, psnr_cubic, psnr_raisr, out=BGR(A) image)."""
    import cv2
    raisr = ClRaisr(img_gray, filter_path=filter_path, device=device) if filter_path else \
        ClRaisr(img_gray, filters=synth.random_filters(2), device=device)           # raisr.py:141
    refHR = cv2.imread(hr_path)                                                     # raisr.py:143
    if refHR is None:
        raise FileNotFoundError(hr_path)
    if lr_path:
        bgr = cv2.imread(lr_path)                                                   # raisr.py:147
        if bgr is None:
            raise FileNotFoundError(lr_path)
        hHR, wHR = 2 * bgr.shape[0], 2 * bgr.shape[1]
        refHR = refHR[:hHR, :wHR]
    else:
        hHR, wHR = refHR.shape[0] // 2 * 2, refHR.shape[1] // 2 * 2
        refHR = refHR[:hHR, :wHR]
        bgr = cv2.resize(refHR, (wHR // 2, hHR // 2))                               # raisr.py:146
    refUp = cv2.resize(bgr, (wHR, hHR), interpolation=cv2.INTER_LINEAR)             # raisr.py:152
    if img_gray == 1:                                                               # raisr.py:157-161
        ycrcb = cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb)
        ycrcb_dst = cv2.resize(ycrcb, (wHR, hHR), interpolation=cv2.INTER_LINEAR)
        src = ycrcb[:, :, 0].copy()
        dst = np.zeros((hHR, wHR), dtype=src.dtype)
    else:                                                                           # raisr.py:163-164
        src = cv2.cvtColor(bgr, cv2.COLOR_BGR2BGRA)
        dst = np.zeros((hHR, wHR, 4), dtype=np.uint8)
    elapsed_list, count = None, 0
    while count < loopcount:                                                        # raisr.py:167-174
        elapsed = raisr.upsample(src, dst, 2)
        if elapsed_list is None:
            elapsed_list = elapsed
        else:
            for i in range(len(elapsed)):
                elapsed_list[i] += elapsed[i]
        count += 1
    if img_gray == 1:                                                               # raisr.py:176-178
        ycrcb_dst[:, :, 0] = dst
        dst = cv2.cvtColor(ycrcb_dst, cv2.COLOR_YCrCb2BGR)
    if write:
        cv2.imwrite(os.path.join(out_dir, "raisr-out.png"), dst)                    # raisr.py:180-181
        cv2.imwrite(os.path.join(out_dir, "raisr-ref-upsample.png"), refUp)
    res = dict(elapsed=[t / count for t in elapsed_list], psnr_cubic=psnr(refUp, refHR),
               psnr_raisr=psnr(dst[:, :, 0:3], refHR), out=dst)                     # raisr.py:184-185
    raisr.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("hr")
    ap.add_argument("lr", nargs="?")
    ap.add_argument("--gray", type=int, default=0, choices=[0, 1], help="imgGray of raisr.py:139 (default 0 = BGRA colour mode)")
    ap.add_argument("--filters", default=None, help="pickle of the (24,3,3,4,121) table (filter.p)")
    ap.add_argument("--loops", type=int, default=20)
    args = ap.parse_args()
    res = run_demo(args.hr, args.lr, args.gray, args.filters, args.loops)
    print("elapsed: {:.3f} + {:.3f} + {:.3f} ms".format(*res["elapsed"]))                       # raisr.py:182
    print("PSNR: cubic {:.3f} raisr {:.3f}".format(res["psnr_cubic"], res["psnr_raisr"]))       # raisr.py:186


if __name__ == "__main__":
    main()
