"""Soak test of the barrier-free tile pipeline: many batches of random-ish frames through the pipelined filter kernel
and through the per-tile-barrier kernel (`filter_pipe=0`); every output byte must agree.  python tools/soak_pipe.py [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oclcomputervision_b200 import ClRaisr, synth  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    bad = 0
    for s, shape, n in ((2, (1080, 1920), 6), (3, (600, 800), 4), (2, (333, 517), 9), (4, (270, 480), 3)):
        flt = synth.random_filters(s, seed=s)
        a = ClRaisr(1, device=0)
        b = ClRaisr(1, device=0)
        setattr(a, "filters_x%d" % s, flt)
        setattr(b, "filters_x%d" % s, flt)
        b.set_option("filter_pipe", 0)
        rng = np.random.default_rng(s)
        for it in range(iters):
            frames = np.stack([synth.synthetic_frame(shape[0], shape[1], seed=int(rng.integers(1 << 30))) for _ in range(2)])
            frames = np.ascontiguousarray(np.tile(frames, ((n + 1) // 2, 1, 1))[:n])
            frames[n // 2] = rng.integers(0, 256, shape, dtype=np.uint8)        # one pure-noise frame: every bucket
            da = np.empty((n, shape[0] * s, shape[1] * s), np.uint8)
            db = np.empty_like(da)
            a.upsample_batch(frames, da, s)
            b.upsample_batch(frames, db, s)
            if not np.array_equal(da, db):
                bad += 1
                print("MISMATCH scale", s, "iteration", it, int((da != db).sum()), "bytes")
        a.close(); b.close()
        print("scale", s, shape, "x", n, "frames:", iters, "iterations compared")
    print("soak:", "FAILED" if bad else "ok")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
