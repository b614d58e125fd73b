"""Colour (BGRA) RAISR throughput, SURVEY.md 8(f) row N1 -- what the reference's own __main__ runs.
    python tools/bench_color.py [--frames 8] [--steps 10] [--src 1920x1080] [--scale 2] [--cpu-frames 1]
One JSON line: device-resident and end-to-end (pinned host buffers, copies inside) output Mpix/s of
`raisr_upsample_bgra_u8`, the per-kernel split is in the launch list (profiles/).  One output pixel = 4 samples."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oclcomputervision_b200 import _cabi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--src", default="1920x1080")
    ap.add_argument("--scale", type=int, default=2)
    ap.add_argument("--cpu-frames", type=int, default=1)
    a = ap.parse_args()
    sw, sh = (int(v) for v in a.src.split("x"))
    s = a.scale
    dw, dh = sw * s, sh * s
    lib = _cabi.load()
    h = ctypes.c_void_p()
    _cabi.check(lib.raisr_create(ctypes.byref(h), 0, 24, 3, 3, 11))
    flt = synth.random_filters(s)
    _cabi.check(lib.raisr_set_filters(h, s, flt.ctypes.data, flt.size))
    # BGRA frames: three differently seeded luma fields + opaque alpha
    pool = min(a.frames, 4)
    frames = np.empty((pool, sh, sw, 4), np.uint8)
    for k in range(pool):
        for c in range(3):
            frames[k, :, :, c] = synth.synthetic_frame(sh, sw, seed=2000 + 7 * k + c)
        frames[k, :, :, 3] = 255
    src_frame, dst_frame = sw * sh * 4, dw * dh * 4
    hsrc, hdst = ctypes.c_void_p(), ctypes.c_void_p()
    _cabi.check(lib.raisr_host_alloc(ctypes.byref(hsrc), src_frame * a.frames))
    _cabi.check(lib.raisr_host_alloc(ctypes.byref(hdst), dst_frame * a.frames))
    src_np = np.ctypeslib.as_array(ctypes.cast(hsrc, ctypes.POINTER(ctypes.c_uint8)), (a.frames, sh, sw, 4))
    dst_np = np.ctypeslib.as_array(ctypes.cast(hdst, ctypes.POINTER(ctypes.c_uint8)), (a.frames, dh, dw, 4))
    for k in range(a.frames):
        src_np[k] = frames[k % pool]
    dsrc, ddst = ctypes.c_void_p(), ctypes.c_void_p()
    _cabi.check(lib.raisr_dev_alloc(h, ctypes.byref(dsrc), src_frame * a.frames))
    _cabi.check(lib.raisr_dev_alloc(h, ctypes.byref(ddst), dst_frame * a.frames))
    ms = (ctypes.c_float * 3)()
    # e2e first: it also leaves the inputs on the device for the resident run
    e2e = []
    for i in range(a.steps + a.warmup):
        t0 = time.perf_counter()
        _cabi.check(lib.raisr_upsample_bgra_u8(h, hsrc, sw, sh, sw * 4, hdst, dw, dh, dw * 4, s, a.frames, _cabi.RAISR_HOST, ms))
        t1 = time.perf_counter()
        if i >= a.warmup:
            e2e.append(((t1 - t0) * 1e3, ms[0], ms[1], ms[2]))
    import torch
    t = torch.from_numpy(src_np.copy()).cuda()
    out = torch.empty((a.frames, dh, dw, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    n0 = lib.raisr_launch_count(h)
    dev = []
    for i in range(a.steps + a.warmup):
        _cabi.check(lib.raisr_upsample_bgra_u8(h, ctypes.c_void_p(t.data_ptr()), sw, sh, sw * 4, ctypes.c_void_p(out.data_ptr()), dw, dh, dw * 4, s,
                                               a.frames, _cabi.RAISR_DEVICE, ms))
        if i >= a.warmup:
            dev.append(ms[1])
    launches = (lib.raisr_launch_count(h) - n0) // (a.steps + a.warmup)
    same = bool(np.array_equal(out.cpu().numpy(), dst_np))
    mpix = a.frames * dw * dh / 1e6
    line = {"metric": "RAISR %dx BGRA output Mpix/s" % s, "unit": "Mpix/s", "frames": a.frames, "src": a.src, "dst": "%dx%d" % (dw, dh),
            "value": round(mpix / (np.median(dev) / 1e3), 1), "ms_per_step": round(float(np.median(dev)), 3),
            "e2e": {"value": round(mpix / (np.median([e[0] for e in e2e]) / 1e3), 1), "h2d_kernel_d2h_ms": [round(float(np.median([e[k] for e in e2e])), 3) for k in (1, 2, 3)],
                    "h2d_bytes_per_step": src_frame * a.frames, "d2h_bytes_per_step": dst_frame * a.frames},
            "gpu_launches_per_step": int(launches), "host_path_matches_device_path": same,
            "flop_per_px": 13 * 4 + 28 + 10 + 105 + 40 + 242 * 4 + 28 + 8,
            "note": "one pixel = 4 samples (B,G,R,A); the hashed filter runs on all four planes (raisr.cl:322-330)"}
    line["achieved_tflops"] = round(line["flop_per_px"] * line["value"] * 1e6 / 1e12, 2)
    if a.cpu_frames > 0:
        from oracle import raisr_oracle as O
        t0 = time.perf_counter()
        for k in range(a.cpu_frames):
            ref = O.raisr_ref_bgra_c(frames[k % pool], flt, s)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(a.cpu_frames * dw * dh / 1e6 / dt, 2), "unit": "Mpix/s", "cores": O.c_max_threads(), "kind": "port",
                                "sample": "%d frame(s), C oracle raisr_oracle_run_bgra" % a.cpu_frames}
        got = dst_np[0].astype(int)
        line["max_abs_diff_vs_oracle_u8"] = int(np.abs(got - ref["out_u8"].astype(int)).max())
    print(json.dumps(line))


if __name__ == "__main__":
    main()
