#!/usr/bin/env python
"""Builds oracle/_ref/: the reference's OWN kernel sources, executable on the CPU.  TEST INFRASTRUCTURE.

    python oracle/build_ref.py            # needs /root/reference (this container); the GPU box uses committed fixtures

The reference's device code is OpenCL C -- /root/reference/super_resolution/raisr.cl (the hot path) and
/root/reference/basic/interpolation.cl (the stand-alone resizers), /root/reference/histeq/hist.cl
(histogram equalisation, N4); there is no OpenCL platform here
(SURVEY.md F5).  This recipe reads those files WHERE THEY LIE, applies the spelling changes C++ needs -- each a fixed
regular expression, listed below and checked to have fired -- and compiles the result against
oracle/ref_shim/cl_shim.hpp (OpenCL C types, built-ins, images, one host thread per work-item of a work-group with a
real barrier) into shared objects:

    oracle/_ref/libraisr_ref_shipped_f16.so   raisr.cl exactly as shipped (`#if 1` early return, raisr.cl:219-230)
    oracle/_ref/libraisr_ref_full_f16.so      the same text with that one `#if 1` switched off: the "dead" RAISR code runs
    oracle/_ref/libraisr_ref_{shipped,full}_f32.so   the same two with `half` kept in binary32 (the shim's
                                              CL_SHIM_HALF_IS_FLOAT): the kernel text in the arithmetic the oracle restates
    oracle/_ref/libraisr_ref_intended_{f16,f32}.so   the full text with the three slips of raisr.cl:271,310,316 corrected (INTENDED_FIXES)
    oracle/_ref/libraisr_ref_cubic_intended_f32.so   the same with stage 1 switched to the file's own cubic_sample (CUBIC_SWITCH)
    oracle/_ref/libinterp_ref_{f16,f32}.so    interpolation.cl: bilinear_simple, bilinear_lds, bicubic_simple, bicubic_lds
    oracle/_ref/libhist_ref.so                histeq/hist.cl: hist, histeq_global, histeq_local_block

Nothing of the reference is copied into the repository: the rewritten text is piped to the compiler, and oracle/_ref/
(git-ignored) only ever holds the shared objects.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CL = "/root/reference/super_resolution/raisr.cl"
REF_INTERP_CL = "/root/reference/basic/interpolation.cl"
REF_HIST_CL = "/root/reference/histeq/hist.cl"
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim", "cl_shim.hpp")

# (pattern, replacement, minimum number of hits) -- spelling only, no change of meaning
COMMON = [
    (r"\((half[234]|float[24]|int2)\)\s*\(", r"\1(", 10),            # vector literals: (half4)(a, b, c, d) -> half4(a, b, c, d)
    (r"(?<![\w.])(\d+\.\d+)h\b", r"half(\1f)", 10),                   # half literals: 0.5h -> half(0.5f)
    (r"^\s*#pragma .*$", r"", 5),                                     # OPENCL EXTENSION / unroll pragmas
]
HIST = [
    (r"\((uint4|int2)\)\s*\(", r"\1(", 7),                            # vector literals
    (r"^\s*#pragma .*$", r"", 4),
    (r"__local\s+(\w+)\s*\*", r"\1 *", 1),                           # pointer INTO local memory (hist.cl:83), not a local variable
]
RAISR_ONLY = [
    (r"\.s210\b", r".s210()", 3),                                     # swizzle used by CONV3x3
    (r"__local\s+half4\s+block\[", r"half4 block[", 2),               # __local pointer parameters of the two samplers
    (r"^#if 1\s*$", r"#if RAISR_EARLY_RETURN", 1),                    # the early return of raisr.cl:219, now a build switch
]

# The "intended" build only: the three slips of the dead code corrected IN THE REFERENCE'S TEXT, one token
# each, so that the semantics the product defaults to are also computed by the reference's own kernel -- everything around the three
# tokens (window geometry, Sobel flip, gaussian[j][i], the eigen formulas, NaN behaviour of sqrt(L2), the 121-tap loop) is the reference's
INTENDED_FIXES = [
    (r"ma \+= gx \* gy \* gaussian\[j\]\[i\];", r"ma += gx * gx * gaussian[j][i];", 1),                       # raisr.cl:271
    (r"if \(L1 < coherence_quantizers\[i\]\)", r"if (coherence < coherence_quantizers[i])", 1),              # raisr.cl:310
    (r"\(\(\(angle_idx \* NUM_STRENGTH\) \* NUM_COHERENCE", r"(((angle_idx * NUM_STRENGTH + strength_idx) * NUM_COHERENCE", 1),  # raisr.cl:316
]

# The "cubic" build only: stage 1 calls the file's own (never called) cubic_sample instead of linear_sample (raisr.cl:63-106,210)
CUBIC_SWITCH = [(r"= linear_sample\(preload_block", r"= cubic_sample(preload_block", 1)]

PREAMBLE = r'''
#include <thread>
#include <vector>
thread_local cl_item cl_self;
pthread_barrier_t* cl_group_barrier = nullptr;

// one host thread per work-item of a work-group; the groups run one after the other
template <class K>
static void run_groups(int dw, int dh, int lw, int lh, K kernel)
{
    const int n = lw * lh;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, n);
    cl_group_barrier = &bar;
    std::vector<std::thread> pool;
    for (int t = 0; t < n; ++t)
        pool.emplace_back([&, t]() {
            for (int gy = 0; gy < dh / lh; ++gy)
                for (int gx = 0; gx < dw / lw; ++gx) {
                    cl_self = cl_item{{gx * lw + t % lw, gy * lh + t / lw}, {t % lw, t / lw}, {gx, gy}, {lw, lh}, {dw, dh}, {dw / lw, dh / lh}};
                    kernel();
                    pthread_barrier_wait(&bar);      // the next work-group re-uses the __local (static) arrays
                }
        });
    for (auto& th : pool) th.join();
    pthread_barrier_destroy(&bar);
}
'''

RAISR_DRIVER = PREAMBLE + r'''
extern "C" int raisr_cl_run(const uint8_t* src, int sw, int sh, int src_pitch, int channels, uint8_t* dst, int dw, int dh,
                            int dst_pitch, const float* grad_x, const float* grad_y, const float* csc_to, const float* csc_from,
                            const float* gaussian, const float* sq, const float* cq, int scale, const float* filters)
{
    // raisr.py:117-131: global size = dst size, local size (16, 16); the reference silently needs multiples of 16
    if (dw % WORK_GROUP_WIDTH || dh % WORK_GROUP_HEIGHT || (channels != 1 && channels != 4)) return -1;
    image2d s{const_cast<uint8_t*>(src), sw, sh, src_pitch, channels == 1 ? CLK_R : CLK_BGRA};
    image2d d{dst, dw, dh, dst_pitch, channels == 1 ? CLK_R : CLK_BGRA};
    run_groups(dw, dh, WORK_GROUP_WIDTH, WORK_GROUP_HEIGHT,
               [&]() { raisr(&s, &d, grad_x, grad_y, csc_to, csc_from, gaussian, sq, cq, scale, filters); });
    return 0;
}
'''

INTERP_DRIVER = PREAMBLE + r'''
// interpolation.py:37-117: BGRA UNORM_INT8 images (CL_R accepted too); global size = dst size; local size None for the two
// "simple" kernels (one work-item is a group here), (16, 16) for the two LDS kernels, which therefore need multiples of 16
extern "C" int interp_cl_run(int kernel, const uint8_t* src, int sw, int sh, int src_pitch, int channels, uint8_t* dst, int dw,
                             int dh, int dst_pitch)
{
    if (channels != 1 && channels != 4) return -1;
    image2d s{const_cast<uint8_t*>(src), sw, sh, src_pitch, channels == 1 ? CLK_R : CLK_BGRA};
    image2d d{dst, dw, dh, dst_pitch, channels == 1 ? CLK_R : CLK_BGRA};
    const bool lds = kernel == 1 || kernel == 3;
    if (lds && (dw % 16 || dh % 16)) return -1;
    const int l = lds ? 16 : 1;
    switch (kernel) {
        case 0: run_groups(dw, dh, l, l, [&]() { bilinear_simple(&s, &d); }); break;
        case 1: run_groups(dw, dh, l, l, [&]() { bilinear_lds(&s, &d); }); break;
        case 2: run_groups(dw, dh, l, l, [&]() { bicubic_simple(&s, &d); }); break;
        case 3: run_groups(dw, dh, l, l, [&]() { bicubic_lds(&s, &d); }); break;
        default: return -1;
    }
    return 0;
}
'''

HIST_DRIVER = PREAMBLE + r'''
// eq_opencl.py:37-89: CL_R / CL_UNSIGNED_INT8 images; `hist` runs (w / 256, h) work-items in groups of (1, 32), the two
// equalisation kernels (w, h) in groups of (16, 16)
extern "C" int hist_cl_run(const uint8_t* img, int w, int h, int pitch, uint32_t* hist_out)
{
    if (w % HIST_BINS || h % HIST_THREAD_NUM) return -1;
    image2d s{const_cast<uint8_t*>(img), w, h, pitch, CLK_R};
    run_groups(w / HIST_BINS, h, 1, HIST_THREAD_NUM, [&]() { hist(&s, hist_out); });
    return 0;
}
extern "C" int histeq_global_cl_run(const uint8_t* img, int w, int h, int pitch, uint8_t* dst, int dst_pitch, const uint8_t* mapping)
{
    if (w % 16 || h % 16) return -1;
    image2d s{const_cast<uint8_t*>(img), w, h, pitch, CLK_R}, d{dst, w, h, dst_pitch, CLK_R};
    run_groups(w, h, 16, 16, [&]() { histeq_global(&s, &d, mapping); });
    return 0;
}
extern "C" int histeq_local_block_cl_run(const uint8_t* img, int w, int h, int pitch, uint8_t* dst, int dst_pitch, const float* grid,
                                         int bw, int bh, int nx, int ny)
{
    if (w % 16 || h % 16) return -1;
    image2d s{const_cast<uint8_t*>(img), w, h, pitch, CLK_R}, d{dst, w, h, dst_pitch, CLK_R};
    run_groups(w, h, 16, 16, [&]() { histeq_local_block(&s, &d, grid, bw, bh, nx, ny); });
    return 0;
}
'''

# name -> (source, rewrites, driver, [(library suffix, extra compiler flags)])
UNITS = {
    "raisr": (REF_CL, COMMON + RAISR_ONLY, RAISR_DRIVER,
              [("libraisr_ref_%s_%s.so" % (k, p), ["-DRAISR_EARLY_RETURN=%d" % (k == "shipped")] + (["-DCL_SHIM_HALF_IS_FLOAT"] if p == "f32" else []))
               for k in ("shipped", "full") for p in ("f16", "f32")]),
    "interp": (REF_INTERP_CL, COMMON, INTERP_DRIVER,
               [("libinterp_ref_%s.so" % p, ["-DCL_SHIM_HALF_IS_FLOAT"] if p == "f32" else []) for p in ("f16", "f32")]),
    "raisr_intended": (REF_CL, COMMON + RAISR_ONLY + INTENDED_FIXES, RAISR_DRIVER,
                       [("libraisr_ref_intended_%s.so" % p, ["-DRAISR_EARLY_RETURN=0"] + (["-DCL_SHIM_HALF_IS_FLOAT"] if p == "f32" else []))
                        for p in ("f16", "f32")]),
    "raisr_cubic": (REF_CL, COMMON + RAISR_ONLY + INTENDED_FIXES + CUBIC_SWITCH, RAISR_DRIVER,
                    [("libraisr_ref_cubic_intended_f32.so", ["-DRAISR_EARLY_RETURN=0", "-DCL_SHIM_HALF_IS_FLOAT"])]),
    # eq_opencl.py:26: -DHIST_BINS=256 -DHIST_THREAD_NUM=32 -DHIST_N=8 (no half arithmetic in this file)
    "hist": (REF_HIST_CL, HIST, HIST_DRIVER, [("libhist_ref.so", ["-DHIST_BINS=256", "-DHIST_THREAD_NUM=32", "-DHIST_N=8"])]),
}


def lib_paths():
    return [os.path.join(OUT, lib) for unit in UNITS.values() for lib, _ in unit[3]]


def build(force=False):
    if not os.path.exists(REF_CL):
        raise FileNotFoundError(REF_CL + " (the reference tree only exists in the build container)")
    os.makedirs(OUT, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for name, (source, rewrites, driver, libs) in UNITS.items():
        paths = [os.path.join(OUT, lib) for lib, _ in libs]
        newest = max(os.path.getmtime(p) for p in (source, __file__, SHIM))
        if not force and all(os.path.exists(l) and os.path.getmtime(l) >= newest for l in paths):
            continue
        text = open(source).read()
        for pat, rep, min_hits in rewrites:
            text, hits = re.subn(pat, rep, text, flags=re.M)
            if hits < min_hits:
                raise RuntimeError("%s: rewrite %r fired %d times (expected >= %d): the reference source changed" % (name, pat, hits, min_hits))
        # the translation unit goes to the compiler through a pipe: no reference-derived source file is ever written
        unit = ("// generated by oracle/build_ref.py from %s\n" % source) + '#include "cl_shim.hpp"\n' + text + driver
        for (lib, flags), path in zip(libs, paths):
            subprocess.run([cxx, "-x", "c++", "-", "-I", os.path.dirname(SHIM), "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread",
                            "-ffp-contract=off", "-fno-fast-math", "-w"] + flags + ["-o", path], input=unit.encode(), check=True)
    return lib_paths()


if __name__ == "__main__":
    print("\n".join(build(force="--force" in sys.argv)))
