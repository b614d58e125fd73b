"""Soak: the two-types-per-CTA filter kernel (filter_duo = 1) against the one-type kernel over many random batches
(sizes, frame counts, chunk budgets): any difference would be a race in the tile pipeline or a geometry slip."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oclcomputervision_b200 import ClRaisr, synth

def main(iters=int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    rng = np.random.default_rng(123)
    flt = synth.random_filters(2, seed=3)
    a, b = ClRaisr(1, filters=flt), ClRaisr(1, filters=flt)
    b.set_option("filter_duo", 0)
    bad = 0
    t0 = time.time()
    for it in range(iters):
        sh, sw, n = int(rng.integers(1, 400)), int(rng.integers(1, 700)), int(rng.integers(1, 9))
        frames = rng.integers(0, 256, (n, sh, sw), dtype=np.uint8)
        if it % 3 == 0:
            frames[:] = synth.synthetic_frame(max(sh, 8), max(sw, 8), seed=it)[:sh, :sw]
        budget = int(rng.choice([1 << 20, 8 << 20, 208 << 20]))
        outs = []
        for r in (a, b):
            r.set_option("chunk_budget_bytes", budget)
            dst = np.empty((n, 2 * sh, 2 * sw), np.uint8)
            r.upsample_batch(frames, dst, 2)
            outs.append(dst)
        if not np.array_equal(outs[0], outs[1]):
            bad += 1
            print("MISMATCH", it, sh, sw, n, budget, int((outs[0] != outs[1]).sum()))
    print("soak_duo: %d iterations, %d mismatches, %.1f s" % (iters, bad, time.time() - t0))
    a.close(); b.close()
    return bad

if __name__ == "__main__":
    sys.exit(1 if main() else 0)
