#!/usr/bin/env python
"""Runs the five BASELINE.json configs on one B200 (device-resident kernel time, CUDA events, median of
the timed repetitions after 3 warm-ups) next to the C oracle on the host cores, and checks parity of one
frame per config against the oracle.  Writes one JSON object per config to stdout / --out.

    python tools/bench_configs.py [--configs 1,2,3,4,5] [--reps 5] [--out profiles/r1_configs.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(name="lenna 512x512 -> 1024x1024 (tests/golden fixture), x2", sw=512, sh=512, s=2, frames=1),
    2: dict(name="1080p -> 4K, batch 64, x2", sw=1920, sh=1080, s=2, frames=64),
    3: dict(name="4K -> 8K, batch 512 (64 resident per pass on one GPU), x2", sw=3840, sh=2160, s=2, frames=64, nominal_frames=512),
    4: dict(name="16384x16384 -> 49152x49152 single image, x3 (one GPU, no banding)", sw=16384, sh=16384, s=3, frames=1),
    5: dict(name="720p -> 1440p, batch 2048 (512 resident per pass), x2", sw=1280, sh=720, s=2, frames=512, nominal_frames=2048),
}


def make_frames(cfg):
    from oclcomputervision_b200 import synth
    if cfg["sw"] == 512 and cfg["frames"] == 1:
        g = np.load(os.path.join(ROOT, "tests", "golden", "lenna_x2.npz"))
        return g["src"][None].copy()
    if cfg["sw"] >= 8192:   # tile a 2048x2048 synthetic frame
        t = synth.synthetic_frame(2048, 2048, 1000)
        reps = cfg["sw"] // 2048
        return np.ascontiguousarray(np.tile(t, (reps, reps)))[None]
    pool = synth.synthetic_batch(8, cfg["sh"], cfg["sw"], pool=8, seed=1000)
    reps = (cfg["frames"] + 7) // 8
    return np.ascontiguousarray(np.tile(pool, (reps, 1, 1))[:cfg["frames"]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from oclcomputervision_b200 import ClRaisr, synth
    from oracle import raisr_oracle as O
    threads = len(os.sched_getaffinity(0))
    out = open(args.out, "w") if args.out else None
    for cid in [int(c) for c in args.configs.split(",")]:
        cfg = CONFIGS[cid]
        s, sw, sh, n = cfg["s"], cfg["sw"], cfg["sh"], cfg["frames"]
        dw, dh = sw * s, sh * s
        F = synth.random_filters(s)
        r = ClRaisr(1)
        setattr(r, "filters_x%d" % s, F)
        tap_format, b24_bound = r.effective_filters(s)[1:]
        host = make_frames(cfg)
        src = torch.from_numpy(host).cuda()
        dst = torch.empty((n, dh, dw), dtype=torch.uint8, device="cuda")
        stream = torch.cuda.current_stream()
        r.set_stream(stream.cuda_stream)
        run = lambda timed: r.upsample_device(src.data_ptr(), sw, sh, sw, dst.data_ptr(), dw, s, n, np.uint8, timed=timed)
        for _ in range(3):
            run(False)
        torch.cuda.synchronize()
        times, kern = [], []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(True)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            kern.append(r.last_kernel_ms())
        ms = float(np.median(times))
        prep_ms, filt_ms = [float(np.median([k[i] for k in kern])) for i in (0, 1)]
        mpix = n * dw * dh / 1e6
        gpu_mpix_s = mpix / (ms * 1e-3)
        # one frame through the reference-shaped call: [h2d, kernel, d2h] ms (raisr.py:135,182)
        one_dst = np.zeros((dh, dw), np.uint8)
        r.set_stream(0)
        one_ms = r.upsample(host[0], one_dst, s)
        # CPU oracle on a bounded subset
        res = dict(config=cid, name=cfg["name"], src="%dx%d" % (sw, sh), dst="%dx%d" % (dw, dh), scale=s, frames_timed=n,
                   nominal_frames=cfg.get("nominal_frames", n), gpu_ms=round(ms, 3), gpu_mpix_s=round(gpu_mpix_s, 1),
                   prep_ms=round(prep_ms, 3), filter_ms=round(filt_ms, 3),
                   ffma_roofline_frac=round(412.0 * mpix * 1e6 / (ms * 1e-3) / 74.45e12, 4),
                   one_frame_h2d_kernel_d2h_ms=[round(x, 3) for x in one_ms], taps=tap_format, b24_bound=b24_bound)
        if not args.no_check:
            t0 = time.perf_counter()
            small = dw * dh <= 64e6      # the dense float planes of the classification fit comfortably
            ref = O.raisr_ref_c(host[0], F, s, nthreads=threads, want=("hash", "out_u8", "angle", "L1", "coherence") if small else ("hash", "out_u8"))
            one = time.perf_counter() - t0
            k = int(max(0, min(7, args.cpu_seconds / max(one, 1e-3) - 1)))
            t0 = time.perf_counter()
            for i in range(k):
                O.raisr_ref_c(host[min(i + 1, n - 1)], F, s, nthreads=threads, want=("out_u8",))
            tot = one + (time.perf_counter() - t0)
            cpu_mpix_s = (k + 1) * dw * dh / 1e6 / tot
            h = r.debug_hash(host[0], s)[0] if dw * dh <= 64e6 else None
            got = dst[0].cpu().numpy()
            d = np.abs(got.astype(np.int16) - ref["out_u8"].astype(np.int16))
            if h is not None:
                same = h == ref["hash"]
                excused = (~same) & (O.edge_distance(ref) < 1e-5)      # oracle value within 1e-5 of a bin edge (north_star)
                res.update(hash_mismatch=int((~same).sum()), hash_mismatch_excused=int(excused.sum()),
                           hash_mismatch_unexcused=int(((~same) & ~excused).sum()),
                           out_u8_max_diff_where_hash_equal=int(d[same].max()))
            else:   # image too large for the dense int32/float debug planes: bound the pixels that may differ
                res.update(pixels_over_1lsb=int((d > 1).sum()), pixels=int(d.size),
                           note="the per-pixel classification needs 40 GB of float planes at this size; the same x3 kernels are "
                                "classified in full at 4128x4200 by tests/test_configs_gpu.py (6 excused, 0 unexcused of 17.3 M), "
                                "i.e. ~3.5e-7 of the pixels sit on a bin edge: ~840 expected here")
            res.update(cpu_oracle_mpix_s=round(cpu_mpix_s, 2), cpu_threads=threads, cpu_frames=k + 1,
                       speedup_vs_cpu_oracle=round(gpu_mpix_s / cpu_mpix_s, 1))
        line = json.dumps(res)
        print(line, flush=True)
        if out:
            out.write(line + "\n"); out.flush()
        r.close()
        del src, dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
