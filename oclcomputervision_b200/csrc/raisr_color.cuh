// raisr_color.cuh -- colour (BGRA) RAISR, SURVEY.md 8(f) row N1: what the reference's own __main__ runs
// (/root/reference/super_resolution/raisr.py:101-104,139,163-164 with imgGray = 0).
//
//   color_upscale_kernel   bilinear upscale of B,G,R,A on the extended domain (raisr.cl:48-61 on half4)
//                          + CSC to YUV per sample (raisr.cl:211-214, matrix raisr.py:20-25 applied to
//                          (R,G,B,A): read_imagef of a CL_BGRA image returns RGBA order)
//                          -> four column-major uext planes Y,U,V,A (same layout as the gray uext)
//   prep_kernel<FROM_U>    structure tensor + hash from the Y plane (raisr_prep.cuh)
//   filter_octet_kernel    run once per plane with the shared hash, raw (unclamped) float output: the
//                          reference accumulates a half4 with one luma-derived filter (raisr.cl:322-330)
//   color_pack_kernel      CSC back (raisr.cl:333-336, raisr.py:26-31) + saturating UNORM8 store, BGRA
// fp32 with explicit _rn intrinsics, dot() left to right: bit-identical to raisr_oracle_run_bgra up to
// the hash; the dot may differ in summation order like the gray path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "raisr_prep.cuh"

namespace raisr {

struct ColorUpParams {
    const uint8_t* src; size_t src_pitch;   // (sh, sw, 4) u8 BGRA
    int sw, sh, dw, dh;
    float* plane[4];                         // Y,U,V,A uext planes
    size_t pitch;                            // floats per column
    int cubic;                               // 1 = the reference's cubic_sample (raisr.cl:63-106, a half4 routine) instead of linear_sample
};

__constant__ float kCscToYuv[16] = {0.299f, 0.587f, 0.114f, 0.0f, -0.14713f, -0.28886f, 0.436f, 0.0f,
                                    0.615f, -0.51499f, -0.10001f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};
__constant__ float kCscFromYuv[16] = {1.0f, 0.0f, 1.13983f, 0.0f, 1.0f, -0.39465f, -0.58060f, 0.0f,
                                      1.0f, 2.03211f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};

__device__ __forceinline__ float dot4_rn(const float* m, float a, float b, float c, float d)
{
    float acc = __fmul_rn(m[0], a);
    acc = __fadd_rn(acc, __fmul_rn(m[1], b));
    acc = __fadd_rn(acc, __fmul_rn(m[2], c));
    acc = __fadd_rn(acc, __fmul_rn(m[3], d));
    return acc;
}

// One thread = one extended column x kColorUpRows consecutive extended rows: 32 contiguous bytes (a full sector) per
// plane of the column-major uext.
constexpr int kColorUpRows = 8;
__global__ void __launch_bounds__(256) color_upscale_kernel(const ColorUpParams p)
{
    // read_imagef's UNORM8 decode v / 255 (correctly rounded) from a 256-entry table: the bilinear branch decodes 16
    // bytes per sample, and an IEEE division each made this kernel 2.5x slower than its memory traffic allows
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    __syncthreads();
    const int ec = blockIdx.x * blockDim.x + threadIdx.x;       // extended column
    const int er0 = blockIdx.y * kColorUpRows;                   // first extended row of the thread's run
    const int ew = p.dw + 2 * kMargin, eh = p.dh + 2 * kMargin;
    if (ec >= ew || er0 >= eh) return;
    const float fx = __fmul_rn(__fdiv_rn((float)(ec - kMargin), (float)(p.dw - 1)), (float)(p.sw - 1));
    const float flx = floorf(fx);
    const int xi = (int)flx;
    const float u = __fsub_rn(fx, flx), omu = __fsub_rn(1.0f, u);
    const int x0 = min(max(xi, 0), p.sw - 1), x1 = min(max(xi + 1, 0), p.sw - 1);
    float out[4][kColorUpRows];
    if (p.cubic) {
        // cubic_sample (raisr.cl:63-106): 4x4 taps around floor(coord), w_k = dot((1,u,u2,u3), cubic_matrix[k]) left to
        // right, acc += (pix * xweight[j]) * yweight[i] with i outer / j inner, every channel clamped to [0,1]
        auto weights = [](float t, float (&w)[4]) {
            const float t2 = __fmul_rn(t, t), t3 = __fmul_rn(t2, t);
            auto dot = [&](float m0, float m1, float m2, float m3) {
                float acc = __fadd_rn(__fmul_rn(1.0f, m0), __fmul_rn(t, m1));
                acc = __fadd_rn(acc, __fmul_rn(t2, m2));
                return __fadd_rn(acc, __fmul_rn(t3, m3));
            };
            w[0] = dot(0.0f, -0.5f, 1.0f, -0.5f); w[1] = dot(1.0f, 0.0f, -2.5f, 1.5f);
            w[2] = dot(0.0f, 0.5f, 2.0f, -1.5f);  w[3] = dot(0.0f, 0.0f, -0.5f, 0.5f);
        };
        float xw[4];
        weights(u, xw);
        int xs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xs[j] = min(max(xi - 1 + j, 0), p.sw - 1);
#pragma unroll 1
        for (int k = 0; k < kColorUpRows; ++k) {
            const int er = er0 + k;
            const float fy = __fmul_rn(__fdiv_rn((float)(er - kMargin), (float)(p.dh - 1)), (float)(p.sh - 1));
            const float fly = floorf(fy);
            const int yi = (int)fly;
            float yw[4];
            weights(__fsub_rn(fy, fly), yw);
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};      // B, G, R, A
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uchar4* row = reinterpret_cast<const uchar4*>(p.src + (size_t)min(max(yi - 1 + i, 0), p.sh - 1) * p.src_pitch);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uchar4 q = __ldg(row + xs[j]);
                    const unsigned char ch[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[c] = __fadd_rn(acc[c], __fmul_rn(__fmul_rn(lut[ch[c]], xw[j]), yw[i]));
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = fminf(fmaxf(acc[c], 0.0f), 1.0f);
#pragma unroll
            for (int c = 0; c < 4; ++c) out[c][k] = dot4_rn(kCscToYuv + 4 * c, acc[2], acc[1], acc[0], acc[3]);
        }
    } else
#pragma unroll
    for (int k = 0; k < kColorUpRows; ++k) {
        const int er = er0 + k;
        const float fy = __fmul_rn(__fdiv_rn((float)(er - kMargin), (float)(p.dh - 1)), (float)(p.sh - 1));
        const float fly = floorf(fy);
        const int yi = (int)fly;
        const float v = __fsub_rn(fy, fly), omv = __fsub_rn(1.0f, v);
        const uchar4* r0 = reinterpret_cast<const uchar4*>(p.src + (size_t)min(max(yi, 0), p.sh - 1) * p.src_pitch);
        const uchar4* r1 = reinterpret_cast<const uchar4*>(p.src + (size_t)min(max(yi + 1, 0), p.sh - 1) * p.src_pitch);
        const uchar4 q00 = __ldg(r0 + x0), q01 = __ldg(r0 + x1), q10 = __ldg(r1 + x0), q11 = __ldg(r1 + x1);
        const float w00 = __fmul_rn(omu, omv), w01 = __fmul_rn(u, omv), w10 = __fmul_rn(omu, v), w11 = __fmul_rn(u, v);
        auto bil = [&](unsigned char a, unsigned char b, unsigned char c, unsigned char d) {
            float acc = __fmul_rn(w00, lut[a]);
            acc = __fadd_rn(acc, __fmul_rn(w01, lut[b]));
            acc = __fadd_rn(acc, __fmul_rn(w10, lut[c]));
            acc = __fadd_rn(acc, __fmul_rn(w11, lut[d]));
            return acc;
        };
        const float B = bil(q00.x, q01.x, q10.x, q11.x), G = bil(q00.y, q01.y, q10.y, q11.y);
        const float R = bil(q00.z, q01.z, q10.z, q11.z), A = bil(q00.w, q01.w, q10.w, q11.w);
#pragma unroll
        for (int c = 0; c < 4; ++c) out[c][k] = dot4_rn(kCscToYuv + 4 * c, R, G, B, A);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < kColorUpRows / 4; ++q)      // the column pitch covers the rounded-up row count
            *reinterpret_cast<float4*>(p.plane[c] + (size_t)ec * p.pitch + er0 + 4 * q) =
                make_float4(out[c][4 * q], out[c][4 * q + 1], out[c][4 * q + 2], out[c][4 * q + 3]);
}

struct ColorPackParams {
    const float* plane[4];   // filtered Y,U,V,A, dense dh x dw, row pitch `pitch` floats
    size_t pitch;
    uint8_t* dst; size_t dst_pitch;          // (dh, dw, 4) u8 BGRA
    float* dst_f32; size_t dst_f32_pitch;    // optional (dh, dw, 4) float, pitch in floats
    int dw, dh;
};

__global__ void __launch_bounds__(256) color_pack_kernel(const ColorPackParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.dw) return;
    const size_t o = (size_t)y * p.pitch + x;
    const float Y = p.plane[0][o], U = p.plane[1][o], V = p.plane[2][o], A = p.plane[3][o];
    float rgba[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rgba[k] = dot4_rn(kCscFromYuv + 4 * k, Y, U, V, A);
    const float bgra[4] = {rgba[2], rgba[1], rgba[0], rgba[3]};
    float cl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cl[k] = fminf(fmaxf(bgra[k], 0.0f), 1.0f);
    if (p.dst)
        *reinterpret_cast<uchar4*>(p.dst + (size_t)y * p.dst_pitch + 4 * x) =
            make_uchar4((unsigned char)__float2uint_rn(__fmul_rn(cl[0], 255.0f)), (unsigned char)__float2uint_rn(__fmul_rn(cl[1], 255.0f)),
                        (unsigned char)__float2uint_rn(__fmul_rn(cl[2], 255.0f)), (unsigned char)__float2uint_rn(__fmul_rn(cl[3], 255.0f)));
    if (p.dst_f32) *reinterpret_cast<float4*>(p.dst_f32 + (size_t)y * p.dst_f32_pitch + 4 * x) = make_float4(cl[0], cl[1], cl[2], cl[3]);
}

}  // namespace raisr
