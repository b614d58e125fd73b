"""Single-call latency of the reference-style `upsample(src, dst, 2)` on one 1080p frame, with pageable and with
`ClRaisr.pin`-ned caller arrays (python tools/pin_bench.py)."""
import os
import sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oclcomputervision_b200 import ClRaisr, synth
r = ClRaisr(1, filters=synth.random_filters(2))
src = synth.synthetic_frame(1080, 1920, 1); dst = np.empty((2160, 3840), np.uint8)
def run(tag, s, d):
    for _ in range(3): r.upsample(s, d, 2)
    t0 = time.perf_counter()
    for _ in range(20): ms = r.upsample(s, d, 2)
    dt = (time.perf_counter() - t0) / 20 * 1e3
    print(tag, "wall ms/call %.3f" % dt, "h2d/kernel/d2h", [round(x, 3) for x in ms])
run("pageable", src, dst)
ClRaisr.pin(src); ClRaisr.pin(dst)
run("pinned  ", src, dst)
