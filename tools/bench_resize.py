"""Device-resident timing of the stand-alone resizers (bilinear_lds, bicubic, bilinear) against
the HBM roofline.    python tools/bench_resize.py [--frames 16] [--reps 20]
One JSON line per (mode, channels): algorithmic bytes (source + destination, once) / median kernel time."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oclcomputervision_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import torch
    lib = _cabi.load()
    h = ctypes.c_void_p()
    _cabi.check(lib.raisr_create(ctypes.byref(h), 0, 24, 3, 3, 11))
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6451.0))
    except Exception:
        peak = 6451.0
    sw, sh, s = 1920, 1080, 2
    dw, dh = sw * s, sh * s
    rng = np.random.default_rng(0)
    ms = (ctypes.c_float * 3)()
    for ch in (1, 4):
        src = torch.from_numpy(rng.integers(0, 256, (a.frames, sh, sw * ch), dtype=np.uint8)).cuda()
        dst = torch.empty((a.frames, dh, dw * ch), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        for mode, name in ((0, "bilinear_lds"), (1, "bicubic"), (2, "bilinear")):
            times = []
            for i in range(a.reps + 3):
                _cabi.check(lib.raisr_resize_u8(h, ctypes.c_void_p(src.data_ptr()), sw, sh, sw * ch, ch, ctypes.c_void_p(dst.data_ptr()), dw, dh, dw * ch,
                                                mode, a.frames, _cabi.RAISR_DEVICE, ms))
                if i >= 3:
                    times.append(ms[1])
            med = float(np.median(times))
            nbytes = a.frames * (sw * sh + dw * dh) * ch
            print(json.dumps({"kernel": "resize_kernel<%d>" % ch, "mode": name, "frames": a.frames, "src": "%dx%d" % (sw, sh), "dst": "%dx%d" % (dw, dh),
                              "ms": round(med, 4), "out_gpix_s": round(a.frames * dw * dh / med / 1e6, 1), "algorithmic_bytes": nbytes,
                              "achieved_gbps": round(nbytes / med / 1e6, 1), "peak_gbps": peak, "frac": round(nbytes / med / 1e6 / peak, 3)}))
    lib.raisr_destroy(h)


if __name__ == "__main__":
    main()
