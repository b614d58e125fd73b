"""Golden vectors for the histeq path, produced by the REFERENCE's own CPU code.

Runs only in the development container (needs /root/reference):
    python oracle/make_golden_histeq.py
Imports /root/reference/histeq/eq_global.py and eq_local_block.py unmodified.  Their module-level
`import pyopencl` / `import matplotlib.pyplot` are satisfied by empty stubs (neither package is in this
image and only the `use_gpu=True` branch touches OpenCL); the `use_gpu=False` branches are plain numpy.
Writes tests/golden/histeq_ref.npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/histeq"


def _import_reference():
    cl = types.ModuleType("pyopencl")
    cl.get_platforms = lambda: []          # clHistEq.__init__ loops over platforms: none -> no kernels built
    sys.modules.setdefault("pyopencl", cl)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.path.insert(0, REF)
    import eq_global
    import eq_local_block
    return eq_global, eq_local_block


def test_image(h, w, seed):
    """Smooth illumination field x texture: local contrast differs block to block."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    light = 0.15 + 0.7 * (0.5 + 0.5 * np.sin(xx / w * 5.1 + 0.3) * np.cos(yy / h * 3.7 - 0.8))
    tex = 0.5 + 0.25 * np.sin(xx * 0.21) * np.sin(yy * 0.17) + 0.08 * rng.standard_normal((h, w))
    return np.clip(light * tex * 255.0 * 1.4, 0, 255).astype(np.uint8)


def main():
    eq_global, eq_local_block = _import_reference()
    out = {}
    rng = np.random.default_rng(7)
    # transfer-function known answers over the parameter ranges the reference's callers use
    hists, params, curves = [], [], []
    for k in range(24):
        kind = k % 4
        if kind == 0:
            hist = rng.integers(0, 5000, 256)
        elif kind == 1:
            hist = np.zeros(256, np.int64); hist[rng.integers(0, 256, 12)] = rng.integers(1, 9000, 12)
        elif kind == 2:
            hist = (4000 * np.exp(-0.5 * ((np.arange(256) - rng.integers(30, 220)) / rng.uniform(4, 40)) ** 2)).astype(np.int64) + 1
        else:
            hist = rng.integers(0, 3, 256) * rng.integers(0, 70000, 256)
            hist[rng.integers(0, 256)] += 1
        alpha, punch, clip = [(1, 0.05, 2), (0.5, 0.05, 3), (0.8, 0.01, 4), (0.3, 0.1, 1.5)][k // 6]
        hists.append(hist.astype(np.uint32)); params.append((alpha, punch, clip))
        curves.append(eq_global.calc_transfer_func(hist.astype(np.uint32), alpha, punch, clip))
    out["tf_hist"] = np.stack(hists); out["tf_params"] = np.array(params, np.float64); out["tf_curve"] = np.stack(curves)

    h, w, seed = 512, 512, 11
    gray = test_image(h, w, seed)
    out["img_shape_seed"] = np.array([h, w, seed])
    out["img_crc"] = np.array([int(gray.astype(np.uint64).sum()), int((gray.astype(np.uint64) * (np.arange(gray.size).reshape(gray.shape) % 251)).sum())])
    out["global_default"] = eq_global.histeq_global(gray.copy(), use_gpu=False).astype(np.uint8)
    ga = eq_global.histeq_global(gray.copy(), alpha=0.6, punch=0.02, clip=3, use_gpu=False).astype(np.uint8)
    out["global_a06_rows"] = ga[::16]       # a pure LUT pass: every 16th row pins it
    out["local_default"] = eq_local_block.histeq_local_block(gray.copy(), use_gpu=False)
    out["local_128x256"] = eq_local_block.histeq_local_block(gray.copy(), alpha=0.7, punch=0.03, clip=2.5,
                                                             blockshape=(128, 256), use_gpu=False)
    path = os.path.join(ROOT, "tests", "golden", "histeq_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
