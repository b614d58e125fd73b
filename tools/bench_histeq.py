"""Device-resident timing of the three histeq kernels against the HBM roofline (N4).
    python tools/bench_histeq.py [--size 16384] [--reps 20]
Prints one JSON line per kernel: algorithmic bytes / median kernel time (CUDA events inside the C-ABI call).
The 16K x 16K image (268 MB) is larger than L2, so every launch streams from HBM."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oclcomputervision_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    lib = _cabi.load()
    h = ctypes.c_void_p()
    _cabi.check(lib.raisr_create(ctypes.byref(h), 0, 24, 3, 3, 11))
    n = a.size
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6451.0)) if isinstance(peaks, dict) else 6451.0

    def dev(nbytes):
        p = ctypes.c_void_p()
        _cabi.check(lib.raisr_dev_alloc(h, ctypes.byref(p), nbytes))
        return p

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (n, n), dtype=np.uint8)
    ddst = dev(n * n)
    import torch        # device memory for the inputs only
    t = torch.from_numpy(img).cuda()
    torch.cuda.synchronize()
    src_ptr = ctypes.c_void_p(t.data_ptr())
    ny, nx = n // 256, n // 256
    dhist = dev(n // 32 * nx * 256 * 4)
    maps = torch.from_numpy((rng.random((ny, nx, 256)) * 255).astype(np.float32)).cuda()
    lut = torch.from_numpy(rng.permutation(256).astype(np.uint8)).cuda()
    ms = (ctypes.c_float * 3)()
    runs = {
        "hist_tiles_kernel": (lambda: lib.ocv_hist_grid_u8(h, src_ptr, n, n, n, dhist, _cabi.RAISR_DEVICE, ms), n * n + n // 32 * nx * 1024),
        "lut_apply_kernel": (lambda: lib.ocv_histeq_global_u8(h, src_ptr, n, n, n, ddst, n, ctypes.c_void_p(lut.data_ptr()), _cabi.RAISR_DEVICE, ms), 2 * n * n),
        "lut_blend_kernel": (lambda: lib.ocv_histeq_local_block_u8(h, src_ptr, n, n, n, ddst, n, ctypes.c_void_p(maps.data_ptr()), nx, ny, 256, 256, _cabi.RAISR_DEVICE, ms), 2 * n * n),
    }
    for name, (fn, nbytes) in runs.items():
        times = []
        for i in range(a.reps + 3):
            _cabi.check(fn())
            if i >= 3:
                times.append(ms[1])
        med = float(np.median(times))
        print(json.dumps({"kernel": name, "image": [n, n], "ms": round(med, 4), "algorithmic_bytes": nbytes,
                          "achieved_gbps": round(nbytes / med / 1e6, 1), "peak_gbps": peak,
                          "frac": round(nbytes / med / 1e6 / peak, 3)}))
    for p in (ddst, dhist):
        lib.raisr_dev_free(h, p)
    lib.raisr_destroy(h)


if __name__ == "__main__":
    main()
