// raisr_resize.cuh -- stand-alone interpolation kernels: the four entry points of
// /root/reference/basic/interpolation.py:37-107 (clUtility.bilinear / bilinear_lds / bicubic /
// bicubic_lds) for interleaved 8-bit images with 1 or 4 channels, any output size.
//
//   mode 0  bilinear_lds     interpolation.cl:17-71   align-corners map c = (x/(wout-1))*(win-1)
//   mode 1  bicubic(_lds)    interpolation.cl:79-211  same map, Catmull-Rom (a=-0.5) 4x4, clamp to [0,1]
//   mode 2  bilinear_simple  interpolation.cl:3-15    CLK_NORMALIZED_COORDS_TRUE + CLK_FILTER_LINEAR:
//                            texel position = (x/(wout-1))*win - 0.5 (OpenCL 1.2 spec 8.2, fp32 weights;
//                            a hardware sampler would use 8-bit fixed-point weights)
// All reads go through CLAMP_TO_EDGE; the store is write_imagef to UNORM_INT8 (saturate, round to
// nearest even).  fp32 with explicit _rn intrinsics, bit-identical to oracle/raisr_oracle.c.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "packed_f32.cuh"

namespace raisr {

struct ResizeParams {
    const uint8_t* src; size_t src_pitch, src_frame_stride;
    uint8_t* dst; size_t dst_pitch, dst_frame_stride;
    int sw, sh, dw, dh, channels, mode;
};

__device__ __forceinline__ float unorm8(uint8_t v) { return __fdiv_rn((float)v, 255.0f); }
__device__ __forceinline__ uint8_t to_unorm8(float v)
{
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    return (uint8_t)__float2uint_rn(__fmul_rn(v, 255.0f));
}
__device__ __forceinline__ void cubic_weights(float u, float w[4])
{
    // dot((1,u,u2,u3), cubic_matrix[k]) left to right, interpolation.cl:73-78,104-109
    const float u2 = __fmul_rn(u, u), u3 = __fmul_rn(u2, u);
    w[0] = __fadd_rn(__fadd_rn(__fmul_rn(u, -0.5f), u2), __fmul_rn(u3, -0.5f));
    w[1] = __fadd_rn(__fadd_rn(1.0f, __fmul_rn(u2, -2.5f)), __fmul_rn(u3, 1.5f));
    w[2] = __fadd_rn(__fadd_rn(__fmul_rn(u, 0.5f), __fmul_rn(u2, 2.0f)), __fmul_rn(u3, -1.5f));
    w[3] = __fadd_rn(__fmul_rn(u2, -0.5f), __fmul_rn(u3, 0.5f));
}

// One CTA = 256 output columns x kResizeRows output rows.  Everything that depends only on the column (source
// x indices, x weights) is computed once per thread, everything that depends only on the row once per CTA (a small
// shared table).  The source window the tile needs is decoded ONCE into shared memory as floats (exact quotients
// v/255, the read_imagef UNORM8 decode) -- this is the reference's "_lds" idea (interpolation.cl:17-71) -- so the
// per-pixel loop is 4 (bilinear) or 16 (bicubic) shared loads per channel and no global gathers.  Strong
// down-scaling makes the window too large for shared memory; such tiles read global memory through a 256-entry
// decode table instead.  (The first version recomputed two coordinate divisions and up to 64 decode divisions for
// every output pixel and ran at 1-5 % of the HBM roofline, tools/bench_resize.py.)
constexpr int kResizeRows = 16;
constexpr int kResizeWin = 10240;             // floats of decoded source window per CTA (40 KB)

template <int CH>
__global__ void __launch_bounds__(256) resize_kernel(const ResizeParams p)
{
    __shared__ __align__(16) float win[kResizeWin];
    __shared__ float lut[256];
    __shared__ int rowoff[kResizeRows][4];      // clamped source row indices (2 used by the bilinear modes)
    __shared__ float roww[kResizeRows][4];      // bilinear: (1-v, v); bicubic: the four y weights
    __shared__ int wbox[2];                     // first / last source column of the tile
    const int tid = threadIdx.x;
    const int x = min(blockIdx.x * 256 + tid, p.dw - 1), y0 = blockIdx.y * kResizeRows;
    const bool active = blockIdx.x * 256 + tid < p.dw;
    const int rows = min(kResizeRows, p.dh - y0);
    const int ntap = p.mode == 1 ? 4 : 2, first_tap = p.mode == 1 ? -1 : 0;
    lut[tid] = __fdiv_rn((float)tid, 255.0f);
    if (tid < kResizeRows) {
        const int y = min(y0 + tid, p.dh - 1);
        const float fy = p.mode == 2 ? __fsub_rn(__fmul_rn(__fdiv_rn((float)y, (float)(p.dh - 1)), (float)p.sh), 0.5f)
                                     : __fmul_rn(__fdiv_rn((float)y, (float)(p.dh - 1)), (float)(p.sh - 1));
        const float fly = floorf(fy);
        const int yi = (int)fly;
        const float v = __fsub_rn(fy, fly);
        if (p.mode != 1) {
            rowoff[tid][0] = min(max(yi, 0), p.sh - 1); rowoff[tid][1] = min(max(yi + 1, 0), p.sh - 1);
            roww[tid][0] = __fsub_rn(1.0f, v); roww[tid][1] = v;
        } else {
            float yw[4];
            cubic_weights(v, yw);
#pragma unroll
            for (int i = 0; i < 4; ++i) { rowoff[tid][i] = min(max(yi - 1 + i, 0), p.sh - 1); roww[tid][i] = yw[i]; }
        }
    }
    const uint8_t* src = p.src + (size_t)blockIdx.z * p.src_frame_stride;
    uint8_t* dcol = p.dst + (size_t)blockIdx.z * p.dst_frame_stride + (size_t)x * CH;
    const float fx = p.mode == 2 ? __fsub_rn(__fmul_rn(__fdiv_rn((float)x, (float)(p.dw - 1)), (float)p.sw), 0.5f)
                                 : __fmul_rn(__fdiv_rn((float)x, (float)(p.dw - 1)), (float)(p.sw - 1));
    const float flx = floorf(fx);
    const int xi = (int)flx;
    const float u = __fsub_rn(fx, flx);
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = min(max(xi + first_tap + j, 0), p.sw - 1);
    // clamped indices are monotone in x and y: the first / last thread and row bound the window
    if (tid == 0) wbox[0] = xs[0];
    if (tid == 255) wbox[1] = xs[ntap - 1];
    __syncthreads();
    const int wx0 = wbox[0], ww = wbox[1] - wx0 + 1;
    const int wy0 = rowoff[0][0], wh = rowoff[rows - 1][ntap - 1] - wy0 + 1;
    const bool staged = ww * wh * CH <= kResizeWin;          // CTA-uniform
    const bool src4 = CH == 4 && ((reinterpret_cast<uintptr_t>(src) | p.src_pitch) & 3) == 0;
    __syncthreads();                                         // everybody has read the absolute row table
    if (staged) {
        // window-relative float offsets: a tap address is then row offset + column offset
        if (tid < kResizeRows) {
#pragma unroll
            for (int i = 0; i < 4; ++i) rowoff[tid][i] = (rowoff[tid][i] - wy0) * ww * CH;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) xs[j] = (xs[j] - wx0) * CH;
        for (int idx = tid; idx < ww * wh; idx += 256) {
            const int r = idx / ww, c = idx - r * ww;
            const uint8_t* sp = src + (size_t)(wy0 + r) * p.src_pitch + (size_t)(wx0 + c) * CH;
            if (CH == 4 && src4) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(sp));
                *reinterpret_cast<float4*>(&win[4 * idx]) = make_float4(lut[w & 0xffu], lut[(w >> 8) & 0xffu], lut[(w >> 16) & 0xffu], lut[w >> 24]);
            } else {
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) win[CH * idx + ch] = lut[__ldg(sp + ch)];
            }
        }
        __syncthreads();
    }
    if (!active) return;
    // one source pixel -> CH decoded channels, from the staged window or (fallback) from global memory
    auto fetch_win = [&](int ry, int xx, float (&px)[CH]) {
        const float* w = &win[ry + xx];
        if (CH == 4) { const float4 f = *reinterpret_cast<const float4*>(w); px[0] = f.x; px[1 % CH] = f.y; px[2 % CH] = f.z; px[3 % CH] = f.w; }
        else px[0] = w[0];
    };
    auto fetch_gmem = [&](int ry, int xx, float (&px)[CH]) {
        const uint8_t* row = src + (size_t)ry * p.src_pitch;
        if (CH == 4 && src4) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(row) + xx);
            px[0] = lut[w & 0xffu]; px[1 % CH] = lut[(w >> 8) & 0xffu]; px[2 % CH] = lut[(w >> 16) & 0xffu]; px[3 % CH] = lut[w >> 24];
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) px[c] = lut[__ldg(row + xx * CH + c)];
        }
    };
    auto store = [&](uint8_t* d, const float (&out)[CH]) {
        if (CH == 4) *reinterpret_cast<uchar4*>(d) = make_uchar4(to_unorm8(out[0]), to_unorm8(out[1 % CH]), to_unorm8(out[2 % CH]), to_unorm8(out[3 % CH]));
        else d[0] = to_unorm8(out[0]);
    };
    auto run = [&](auto fetch) {
        if (p.mode != 1) {
            const float omu = __fsub_rn(1.0f, u);
            for (int r = 0; r < rows; ++r) {
                const int ya = rowoff[r][0], yb = rowoff[r][1];
                const float omv = roww[r][0], v = roww[r][1];
                const float w00 = __fmul_rn(omu, omv), w01 = __fmul_rn(u, omv), w10 = __fmul_rn(omu, v), w11 = __fmul_rn(u, v);
                float p00[CH], p01[CH], p10[CH], p11[CH], out[CH];
                fetch(ya, xs[0], p00); fetch(ya, xs[1], p01); fetch(yb, xs[0], p10); fetch(yb, xs[1], p11);
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    float acc = __fmul_rn(w00, p00[c]);
                    acc = __fadd_rn(acc, __fmul_rn(w01, p01[c]));
                    acc = __fadd_rn(acc, __fmul_rn(w10, p10[c]));
                    acc = __fadd_rn(acc, __fmul_rn(w11, p11[c]));
                    out[c] = acc;
                }
                store(dcol + (size_t)(y0 + r) * p.dst_pitch, out);
            }
        } else {
            float xw[4];
            cubic_weights(u, xw);
            for (int r = 0; r < rows; ++r) {
                float out[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) out[c] = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ry = rowoff[r][i];
                    const float yw = roww[r][i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float px[CH];
                        fetch(ry, xs[j], px);
#pragma unroll
                        for (int c = 0; c < CH; ++c) out[c] = __fadd_rn(out[c], __fmul_rn(__fmul_rn(px[c], xw[j]), yw));
                    }
                }
                store(dcol + (size_t)(y0 + r) * p.dst_pitch, out);
            }
        }
    };
    if (staged) run(fetch_win);
    else run(fetch_gmem);
}

// ---- fast path of the two bilinear modes for 8-bit gray images (also what the SHIPPED raisr kernel computes,
// raisr.cl:219-230): a byte stream of 1 + 1/s^2 bytes per output pixel that the generic kernel above ran at 9 % of
// the HBM roofline (one byte store and ~25 instructions per pixel).  Here one thread produces FOUR adjacent output
// pixels of a row -- one 32-bit store, 128 contiguous bytes per warp -- the column terms (two window offsets, u, 1-u)
// of its four columns stay in registers for all rows of the tile, the row terms come from a small shared table, the
// source window is fetched with aligned 32-bit loads (four texels each) and decoded once into shared memory, and the
// arithmetic runs on column pairs with packed fp32 (products only: the additions stay scalar so that ptxas cannot
// contract them into FMAs -- every pixel keeps the oracle's two roundings per term, bit for bit).
constexpr int kFastCols = 512, kFastRows = 32, kFastThreads = 128;
constexpr int kFastWin = 10240;               // floats (40 KB) of decoded window

struct FastRow { int ya, yb; float omv, v; };  // window-relative byte offsets of the two source rows, y weights

// upper bound of the decoded window of one tile, in floats (dynamic shared memory of the launch)
inline long long resize_fast_win_floats(const ResizeParams& p)
{
    const long long ww = ((long long)kFastCols * p.sw / p.dw + 12 + 3) / 4 * 4, wh = (long long)kFastRows * p.sh / p.dh + 5;
    return ww * wh;
}

// host-side eligibility: gray, bilinear mode, 4-byte aligned rows on both sides, window of a tile fits
inline bool resize_fast_ok(const ResizeParams& p)
{
    if (p.channels != 1 || p.mode == 1) return false;
    if (((reinterpret_cast<uintptr_t>(p.src) | p.src_pitch | p.src_frame_stride | reinterpret_cast<uintptr_t>(p.dst) | p.dst_pitch | p.dst_frame_stride) & 3) != 0) return false;
    if (p.dw < 4) return false;
    return resize_fast_win_floats(p) <= kFastWin;
}

__global__ void __launch_bounds__(kFastThreads) resize_bilinear_gray_kernel(const ResizeParams p)
{
    extern __shared__ __align__(16) float win[];               // resize_fast_win_floats(p) floats
    __shared__ float lut[256];
    __shared__ __align__(16) FastRow rowt[kFastRows];
    __shared__ int wbox[4];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFastCols + 4 * tid, y0 = blockIdx.y * kFastRows;
    const int rows = min(kFastRows, p.dh - y0);
    lut[tid] = __fdiv_rn((float)tid, 255.0f);
    lut[tid + 128] = __fdiv_rn((float)(tid + 128), 255.0f);
    auto map = [&](int i, int nd, int ns) {     // interpolation.cl:58-69 (mode 0) / the normalised sampler (mode 2)
        return p.mode == 2 ? __fsub_rn(__fmul_rn(__fdiv_rn((float)i, (float)(nd - 1)), (float)ns), 0.5f)
                           : __fmul_rn(__fdiv_rn((float)i, (float)(nd - 1)), (float)(ns - 1));
    };
    int ya_abs = 0, yb_abs = 0;
    float vrow = 0.0f;
    if (tid < kFastRows) {
        const float fy = map(min(y0 + tid, p.dh - 1), p.dh, p.sh);
        const float fl = floorf(fy);
        ya_abs = min(max((int)fl, 0), p.sh - 1); yb_abs = min(max((int)fl + 1, 0), p.sh - 1);
        vrow = __fsub_rn(fy, fl);
        if (tid == 0) wbox[2] = ya_abs;
        if (tid == rows - 1) wbox[3] = yb_abs;
    }
    int xa[4], xb[4];
    float u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float fx = map(min(x0 + k, p.dw - 1), p.dw, p.sw);
        const float fl = floorf(fx);
        xa[k] = min(max((int)fl, 0), p.sw - 1); xb[k] = min(max((int)fl + 1, 0), p.sw - 1);
        u[k] = __fsub_rn(fx, fl);
    }
    if (tid == 0) wbox[0] = xa[0] & ~3;                       // the window starts on a 4-byte boundary of the source row
    if (tid == kFastThreads - 1) wbox[1] = xb[3];             // clamped indices are monotone in x
    __syncthreads();
    const int wx0 = wbox[0], ww4 = (wbox[1] - wx0) / 4 + 1;   // words (4 texels) per window row
    const int wy0 = wbox[2], wh = wbox[3] - wy0 + 1;
    const int wpitch = 4 * ww4;                               // floats per window row
    if (tid < kFastRows) rowt[tid] = FastRow{(ya_abs - wy0) * wpitch * 4, (yb_abs - wy0) * wpitch * 4, __fsub_rn(1.0f, vrow), vrow};
    const uint8_t* src = p.src + (size_t)blockIdx.z * p.src_frame_stride + (size_t)wy0 * p.src_pitch + wx0;
    for (int r = 0; r < wh; r += 4) {                          // four rows per pass: their loads are in flight together
        for (int c = tid; c < ww4; c += kFastThreads) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = (r + k < wh) ? __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(r + k) * p.src_pitch) + c) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (r + k < wh)
                    *reinterpret_cast<float4*>(&win[(r + k) * wpitch + 4 * c]) =
                        make_float4(lut[w[k] & 0xffu], lut[(w[k] >> 8) & 0xffu], lut[(w[k] >> 16) & 0xffu], lut[w[k] >> 24]);
        }
    }
    __syncthreads();
    if (x0 >= p.dw) return;
    int oa[4], ob[4];                                          // window-relative byte offsets of the two taps of each column
#pragma unroll
    for (int k = 0; k < 4; ++k) { oa[k] = 4 * (xa[k] - wx0); ob[k] = 4 * (xb[k] - wx0); }
    const p2 U01 = pk(u[0], u[1]), U23 = pk(u[2], u[3]);
    const p2 OMU01 = pk(__fsub_rn(1.0f, u[0]), __fsub_rn(1.0f, u[1])), OMU23 = pk(__fsub_rn(1.0f, u[2]), __fsub_rn(1.0f, u[3]));
    const bool full = x0 + 3 < p.dw;
    uint8_t* drow = p.dst + (size_t)blockIdx.z * p.dst_frame_stride + (size_t)y0 * p.dst_pitch + x0;
    const char* wbase = reinterpret_cast<const char*>(win);
    auto W = [&](int off) { return *reinterpret_cast<const float*>(wbase + off); };
    // Texels of the two source rows of the current output row, for the thread's four columns (.lo/.hi = column pairs).
    // Consecutive output rows mostly use the same source rows, or the lower one becomes the upper one: the row table
    // is the same for the whole CTA, so these are uniform branches and an up-scale by s loads 2/s texels per pixel.
    p2 ta01 = 0, tb01 = 0, ta23 = 0, tb23 = 0;      // upper row: taps a (x0) and b (x1)
    p2 ba01 = 0, bb01 = 0, ba23 = 0, bb23 = 0;      // lower row
    auto load_row = [&](int yoff, p2& a01, p2& b01, p2& a23, p2& b23) {
        a01 = pk(W(yoff + oa[0]), W(yoff + oa[1])); b01 = pk(W(yoff + ob[0]), W(yoff + ob[1]));
        a23 = pk(W(yoff + oa[2]), W(yoff + oa[3])); b23 = pk(W(yoff + ob[2]), W(yoff + ob[3]));
    };
    auto blend = [&](p2 OMU, p2 U, p2 OMV, p2 V, p2 ta, p2 tb, p2 ba, p2 bb, unsigned& q0, unsigned& q1) {
        float l0, h0, l1, h1;
        upk(mul2(mul2(OMU, OMV), ta), l0, h0);
        upk(mul2(mul2(U, OMV), tb), l1, h1);
        float accl = __fadd_rn(l0, l1), acch = __fadd_rn(h0, h1);
        upk(mul2(mul2(OMU, V), ba), l0, h0);
        accl = __fadd_rn(accl, l0); acch = __fadd_rn(acch, h0);
        upk(mul2(mul2(U, V), bb), l0, h0);
        accl = __fadd_rn(accl, l0); acch = __fadd_rn(acch, h0);
        // write_imagef to UNORM_INT8: saturate, x255, round to nearest even.  A bilinear blend of [0,1] texels is never
        // negative and at most an ulp above 1, which rounds to 255 like the clamped value: no explicit clamp needed.
        q0 = __float2uint_rn(__fmul_rn(accl, 255.0f));
        q1 = __float2uint_rn(__fmul_rn(acch, 255.0f));
    };
    int cur_ya = -1, cur_yb = -1;
    for (int r = 0; r < rows; ++r) {
        const FastRow rt = rowt[r];
        if (rt.ya != cur_ya || rt.yb != cur_yb) {
            if (rt.ya == cur_yb) { ta01 = ba01; tb01 = bb01; ta23 = ba23; tb23 = bb23; }
            else if (rt.ya != cur_ya) load_row(rt.ya, ta01, tb01, ta23, tb23);
            if (rt.yb == rt.ya) { ba01 = ta01; bb01 = tb01; ba23 = ta23; bb23 = tb23; }
            else load_row(rt.yb, ba01, bb01, ba23, bb23);
            cur_ya = rt.ya; cur_yb = rt.yb;
        }
        const p2 OMV = bc(rt.omv), V = bc(rt.v);
        unsigned q0, q1, q2, q3;
        blend(OMU01, U01, OMV, V, ta01, tb01, ba01, bb01, q0, q1);
        blend(OMU23, U23, OMV, V, ta23, tb23, ba23, bb23, q2, q3);
        uint8_t* d = drow + (size_t)r * p.dst_pitch;
        if (full) {
            *reinterpret_cast<uint32_t*>(d) = __byte_perm(__byte_perm(q0, q1, 0x0040), __byte_perm(q2, q3, 0x0040), 0x5410);
        } else {
            d[0] = (uint8_t)q0;
            if (x0 + 1 < p.dw) d[1] = (uint8_t)q1;
            if (x0 + 2 < p.dw) d[2] = (uint8_t)q2;
        }
    }
}

// ---- the same recipe for interleaved 4-channel (BGRA) images, the format the reference's clUtility works on
// (interpolation.py:43,61,79,97): one thread = one output column (one 32-bit store per row, 128 contiguous bytes per
// warp), its two window offsets and x weights in registers for all rows of the tile, texels decoded once into a
// float4 window, vertical texel reuse, the products of channel pairs (B,G) and (R,A) as packed FMUL2.
constexpr int kFast4Cols = 128, kFast4Rows = 32, kFast4Threads = 128;

inline long long resize_fast4_win_floats(const ResizeParams& p)
{
    const long long ww = (long long)kFast4Cols * p.sw / p.dw + 4, wh = (long long)kFast4Rows * p.sh / p.dh + 5;
    return 4 * ww * wh;
}

inline bool resize_fast4_ok(const ResizeParams& p)
{
    if (p.channels != 4 || p.mode == 1) return false;
    if (((reinterpret_cast<uintptr_t>(p.src) | p.src_pitch | p.src_frame_stride | reinterpret_cast<uintptr_t>(p.dst) | p.dst_pitch | p.dst_frame_stride) & 3) != 0) return false;
    return resize_fast4_win_floats(p) <= kFastWin;
}

__global__ void __launch_bounds__(kFast4Threads) resize_bilinear_bgra_kernel(const ResizeParams p)
{
    extern __shared__ __align__(16) float win[];               // resize_fast4_win_floats(p) floats: [row][column] float4
    __shared__ float lut[256];
    __shared__ __align__(16) FastRow rowt[kFast4Rows];
    __shared__ int wbox[4];
    const int tid = threadIdx.x;
    const int x = blockIdx.x * kFast4Cols + tid, y0 = blockIdx.y * kFast4Rows;
    const int rows = min(kFast4Rows, p.dh - y0);
    lut[tid] = __fdiv_rn((float)tid, 255.0f);
    lut[tid + 128] = __fdiv_rn((float)(tid + 128), 255.0f);
    auto map = [&](int i, int nd, int ns) {
        return p.mode == 2 ? __fsub_rn(__fmul_rn(__fdiv_rn((float)i, (float)(nd - 1)), (float)ns), 0.5f)
                           : __fmul_rn(__fdiv_rn((float)i, (float)(nd - 1)), (float)(ns - 1));
    };
    int ya_abs = 0, yb_abs = 0;
    float vrow = 0.0f;
    if (tid < kFast4Rows) {
        const float fy = map(min(y0 + tid, p.dh - 1), p.dh, p.sh);
        const float fl = floorf(fy);
        ya_abs = min(max((int)fl, 0), p.sh - 1); yb_abs = min(max((int)fl + 1, 0), p.sh - 1);
        vrow = __fsub_rn(fy, fl);
        if (tid == 0) wbox[2] = ya_abs;
        if (tid == rows - 1) wbox[3] = yb_abs;
    }
    const float fx = map(min(x, p.dw - 1), p.dw, p.sw);
    const float flx = floorf(fx);
    const int xa = min(max((int)flx, 0), p.sw - 1), xb = min(max((int)flx + 1, 0), p.sw - 1);
    const float u = __fsub_rn(fx, flx), omu = __fsub_rn(1.0f, u);
    if (tid == 0) wbox[0] = xa;
    if (tid == kFast4Threads - 1) wbox[1] = xb;
    __syncthreads();
    const int wx0 = wbox[0], ww = wbox[1] - wx0 + 1;
    const int wy0 = wbox[2], wh = wbox[3] - wy0 + 1;
    if (tid < kFast4Rows) rowt[tid] = FastRow{(ya_abs - wy0) * ww * 16, (yb_abs - wy0) * ww * 16, __fsub_rn(1.0f, vrow), vrow};
    const uint8_t* src = p.src + (size_t)blockIdx.z * p.src_frame_stride + (size_t)wy0 * p.src_pitch + (size_t)wx0 * 4;
    for (int r = 0; r < wh; r += 4) {                          // four rows per pass: their loads are in flight together
        for (int c = tid; c < ww; c += kFast4Threads) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = (r + k < wh) ? __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(r + k) * p.src_pitch) + c) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (r + k < wh)
                    *reinterpret_cast<float4*>(&win[4 * ((r + k) * ww + c)]) =
                        make_float4(lut[w[k] & 0xffu], lut[(w[k] >> 8) & 0xffu], lut[(w[k] >> 16) & 0xffu], lut[w[k] >> 24]);
        }
    }
    __syncthreads();
    if (x >= p.dw) return;
    const int oa = 16 * (xa - wx0), ob = 16 * (xb - wx0);
    const p2 OMU = bc(omu), U = bc(u);
    uint8_t* drow = p.dst + (size_t)blockIdx.z * p.dst_frame_stride + (size_t)y0 * p.dst_pitch + (size_t)x * 4;
    const char* wbase = reinterpret_cast<const char*>(win);
    auto W4 = [&](int off, p2& bg, p2& ra) {
        const float4 f = *reinterpret_cast<const float4*>(wbase + off);
        bg = pk(f.x, f.y); ra = pk(f.z, f.w);
    };
    p2 ta_bg = 0, ta_ra = 0, tb_bg = 0, tb_ra = 0, ba_bg = 0, ba_ra = 0, bb_bg = 0, bb_ra = 0;   // upper / lower row, taps a / b
    int cur_ya = -1, cur_yb = -1;
    for (int r = 0; r < rows; ++r) {
        const FastRow rt = rowt[r];
        if (rt.ya != cur_ya || rt.yb != cur_yb) {              // the row table is CTA-uniform: uniform branches
            if (rt.ya == cur_yb) { ta_bg = ba_bg; ta_ra = ba_ra; tb_bg = bb_bg; tb_ra = bb_ra; }
            else if (rt.ya != cur_ya) { W4(rt.ya + oa, ta_bg, ta_ra); W4(rt.ya + ob, tb_bg, tb_ra); }
            if (rt.yb == rt.ya) { ba_bg = ta_bg; ba_ra = ta_ra; bb_bg = tb_bg; bb_ra = tb_ra; }
            else { W4(rt.yb + oa, ba_bg, ba_ra); W4(rt.yb + ob, bb_bg, bb_ra); }
            cur_ya = rt.ya; cur_yb = rt.yb;
        }
        const p2 OMV = bc(rt.omv), V = bc(rt.v);
        const p2 w00 = mul2(OMU, OMV), w01 = mul2(U, OMV), w10 = mul2(OMU, V), w11 = mul2(U, V);   // both halves equal
        float q[4];
        auto blend = [&](p2 ta, p2 tb, p2 ba, p2 bb, float& o0, float& o1) {
            float l0, h0, l1, h1;
            upk(mul2(w00, ta), l0, h0);
            upk(mul2(w01, tb), l1, h1);
            float a0 = __fadd_rn(l0, l1), a1 = __fadd_rn(h0, h1);
            upk(mul2(w10, ba), l0, h0);
            a0 = __fadd_rn(a0, l0); a1 = __fadd_rn(a1, h0);
            upk(mul2(w11, bb), l0, h0);
            o0 = __fadd_rn(a0, l0); o1 = __fadd_rn(a1, h0);
        };
        blend(ta_bg, tb_bg, ba_bg, bb_bg, q[0], q[1]);
        blend(ta_ra, tb_ra, ba_ra, bb_ra, q[2], q[3]);
        // write_imagef to UNORM_INT8 (a bilinear blend of [0,1] texels needs no explicit clamp, see the gray kernel)
        const unsigned b0 = __float2uint_rn(__fmul_rn(q[0], 255.0f)), b1 = __float2uint_rn(__fmul_rn(q[1], 255.0f));
        const unsigned b2 = __float2uint_rn(__fmul_rn(q[2], 255.0f)), b3 = __float2uint_rn(__fmul_rn(q[3], 255.0f));
        *reinterpret_cast<uint32_t*>(drow + (size_t)r * p.dst_pitch) = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
    }
}

// ---- bicubic (Catmull-Rom, interpolation.cl:79-211) for BGRA images: one thread = one output column.  The oracle's
// expression is acc += (pix[i][j] * xw[j]) * yw[i], i (row) outer, j inner: the 16 products pix * xw of a column
// depend on the source row only, so they are kept in registers (four row slots, slot = source row & 3) and re-used by
// every output row that needs the row; a new source row costs 4 shared loads and 16 multiplies per column, an output
// row 64 multiplies (as 32 packed FMUL2 over the channel pairs (B,G), (R,A)) and 64 additions in the oracle's order.
// The row schedule (which source rows, their weights, which slot holds which tap) is the same for the whole CTA, so
// the four orders in which the slots can be read are four copies of the blend behind a uniform switch.
struct CubicRow { int y[4]; float w[4]; int base; int pad[3]; };   // window-relative byte offsets of the 4 rows, y weights, first (unclamped) source row

inline long long resize_cubic4_win_floats(const ResizeParams& p)
{
    const long long ww = (long long)kFast4Cols * p.sw / p.dw + 7, wh = (long long)kFast4Rows * p.sh / p.dh + 8;
    return 4 * ww * wh;
}

inline bool resize_cubic4_ok(const ResizeParams& p)
{
    if (p.channels != 4 || p.mode != 1) return false;
    if (((reinterpret_cast<uintptr_t>(p.src) | p.src_pitch | p.src_frame_stride | reinterpret_cast<uintptr_t>(p.dst) | p.dst_pitch | p.dst_frame_stride) & 3) != 0) return false;
    return resize_cubic4_win_floats(p) <= kFastWin;
}

__global__ void __launch_bounds__(kFast4Threads) resize_bicubic_bgra_kernel(const ResizeParams p)
{
    extern __shared__ __align__(16) float win[];
    __shared__ float lut[256];
    __shared__ __align__(16) CubicRow rowt[kFast4Rows];
    __shared__ int wbox[4];
    const int tid = threadIdx.x;
    const int x = blockIdx.x * kFast4Cols + tid, y0 = blockIdx.y * kFast4Rows;
    const int rows = min(kFast4Rows, p.dh - y0);
    lut[tid] = __fdiv_rn((float)tid, 255.0f);
    lut[tid + 128] = __fdiv_rn((float)(tid + 128), 255.0f);
    int yabs[4] = {0, 0, 0, 0}, ybase = 0;
    float yw[4] = {0, 0, 0, 0};
    if (tid < kFast4Rows) {
        const float fy = __fmul_rn(__fdiv_rn((float)min(y0 + tid, p.dh - 1), (float)(p.dh - 1)), (float)(p.sh - 1));
        const float fl = floorf(fy);
        ybase = (int)fl - 1;
        cubic_weights(__fsub_rn(fy, fl), yw);
#pragma unroll
        for (int i = 0; i < 4; ++i) yabs[i] = min(max(ybase + i, 0), p.sh - 1);
        if (tid == 0) wbox[2] = yabs[0];
        if (tid == rows - 1) wbox[3] = yabs[3];
    }
    const float fx = __fmul_rn(__fdiv_rn((float)min(x, p.dw - 1), (float)(p.dw - 1)), (float)(p.sw - 1));
    const float flx = floorf(fx);
    float xw[4];
    cubic_weights(__fsub_rn(fx, flx), xw);
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = min(max((int)flx - 1 + j, 0), p.sw - 1);
    if (tid == 0) wbox[0] = xs[0];
    if (tid == kFast4Threads - 1) wbox[1] = xs[3];
    __syncthreads();
    const int wx0 = wbox[0], ww = wbox[1] - wx0 + 1;
    const int wy0 = wbox[2], wh = wbox[3] - wy0 + 1;
    if (tid < kFast4Rows) {
        CubicRow cr;
#pragma unroll
        for (int i = 0; i < 4; ++i) { cr.y[i] = (yabs[i] - wy0) * ww * 16; cr.w[i] = yw[i]; }
        cr.base = ybase;
        rowt[tid] = cr;
    }
    const uint8_t* src = p.src + (size_t)blockIdx.z * p.src_frame_stride + (size_t)wy0 * p.src_pitch + (size_t)wx0 * 4;
    for (int r = 0; r < wh; r += 4) {
        for (int c = tid; c < ww; c += kFast4Threads) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = (r + k < wh) ? __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(r + k) * p.src_pitch) + c) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (r + k < wh)
                    *reinterpret_cast<float4*>(&win[4 * ((r + k) * ww + c)]) =
                        make_float4(lut[w[k] & 0xffu], lut[(w[k] >> 8) & 0xffu], lut[(w[k] >> 16) & 0xffu], lut[w[k] >> 24]);
        }
    }
    __syncthreads();
    if (x >= p.dw) return;
    int xo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xo[j] = 16 * (xs[j] - wx0);
    uint8_t* drow = p.dst + (size_t)blockIdx.z * p.dst_frame_stride + (size_t)y0 * p.dst_pitch + (size_t)x * 4;
    const char* wbase = reinterpret_cast<const char*>(win);
    // h[slot][j] = (pix * xw[j]) of source row `slot` (mod 4), as channel pairs (B,G) and (R,A)
    p2 hbg[4][4], hra[4][4];
    auto load_row = [&](int slot, int yoff) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4)
            if (s4 == slot) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 f = *reinterpret_cast<const float4*>(wbase + yoff + xo[j]);
                    hbg[s4][j] = mul2(pk(f.x, f.y), bc(xw[j]));
                    hra[s4][j] = mul2(pk(f.z, f.w), bc(xw[j]));
                }
            }
    };
    int have_lo = 1 << 30, have_hi = -(1 << 30);    // unclamped source rows currently in the slots: [have_lo, have_hi]
    for (int r = 0; r < rows; ++r) {
        const CubicRow cr = rowt[r];
        // bring rows cr.base .. cr.base + 3 into the slots (uniform control flow: the row table is per CTA)
        for (int i = 0; i < 4; ++i) {
            const int yr = cr.base + i;
            if (yr < have_lo || yr > have_hi) load_row(yr & 3, cr.y[i]);
        }
        have_lo = cr.base; have_hi = cr.base + 3;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};     // B, G, R, A
        auto blend = [&](int s0) {                   // s0 = slot of tap row 0; rows follow in slots s0+1, s0+2, s0+3 (mod 4)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const p2 YW = bc(cr.w[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float b, g, rr, a;
                    upk(mul2(hbg[(s0 + i) & 3][j], YW), b, g);
                    upk(mul2(hra[(s0 + i) & 3][j], YW), rr, a);
                    acc[0] = __fadd_rn(acc[0], b); acc[1] = __fadd_rn(acc[1], g);
                    acc[2] = __fadd_rn(acc[2], rr); acc[3] = __fadd_rn(acc[3], a);
                }
            }
        };
        switch (cr.base & 3) {
        case 0: blend(0); break;
        case 1: blend(1); break;
        case 2: blend(2); break;
        default: blend(3); break;
        }
        unsigned q[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) q[c] = __float2uint_rn(__fmul_rn(fminf(fmaxf(acc[c], 0.0f), 1.0f), 255.0f));
        *reinterpret_cast<uint32_t*>(drow + (size_t)r * p.dst_pitch) = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
    }
}

}  // namespace raisr
