"""Per-call latency of ClRaisr.upsample on small frames (the reference's own loop, raisr.py:166-182, calls it once per image).
Prints wall time per call next to the three event-timed legs, pageable and pinned arrays."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oclcomputervision_b200 import ClRaisr, synth  # noqa: E402


def main():
    out = []
    for gray, shape in ((1, (512, 512)), (1, (256, 256)), (1, (1080, 1920)), (0, (512, 512))):
        r = ClRaisr(gray)
        r.filters_x2 = synth.random_filters(2)
        src = synth.synthetic_frame(*shape, seed=5)
        if not gray:
            src = np.ascontiguousarray(np.stack([src] * 3 + [np.full(shape, 255, np.uint8)], -1))
        dst = np.zeros((shape[0] * 2, shape[1] * 2) + src.shape[2:], np.uint8)
        for pinned in (False, True):
            if pinned:
                ClRaisr.pin(src); ClRaisr.pin(dst)
            for _ in range(20):
                ms = r.upsample(src, dst, 2)
            n = 200
            t0 = time.perf_counter()
            acc = np.zeros(3)
            for _ in range(n):
                acc += np.asarray(r.upsample(src, dst, 2))
            wall = (time.perf_counter() - t0) / n * 1e3
            out.append(dict(gray=gray, src="%dx%d" % (shape[1], shape[0]), pinned=pinned, wall_ms_per_call=round(wall, 4),
                            h2d_kernel_d2h_ms=[round(float(v) / n, 4) for v in acc], mpix_s=round(dst.shape[0] * dst.shape[1] / wall / 1e3, 1)))
            print(json.dumps(out[-1]), flush=True)
        ClRaisr.unpin(src); ClRaisr.unpin(dst)
        r.close()


if __name__ == "__main__":
    main()
