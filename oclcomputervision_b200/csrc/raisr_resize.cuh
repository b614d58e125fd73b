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
// nearest even).  fp32 with explicit _rn intrinsics, bit-identical to oracle/raisr_oracle.c.  These
// kernels are pure streaming work (HBM-bound, no reuse worth staging): one thread per output pixel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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

template <int CH>
__global__ void __launch_bounds__(256) resize_kernel(const ResizeParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.dw) return;
    const uint8_t* src = p.src + (size_t)blockIdx.z * p.src_frame_stride;
    uint8_t* dst = p.dst + (size_t)blockIdx.z * p.dst_frame_stride + (size_t)y * p.dst_pitch + (size_t)x * CH;
    float fx, fy;
    if (p.mode == 2) {
        fx = __fsub_rn(__fmul_rn(__fdiv_rn((float)x, (float)(p.dw - 1)), (float)p.sw), 0.5f);
        fy = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, (float)(p.dh - 1)), (float)p.sh), 0.5f);
    } else {
        fx = __fmul_rn(__fdiv_rn((float)x, (float)(p.dw - 1)), (float)(p.sw - 1));
        fy = __fmul_rn(__fdiv_rn((float)y, (float)(p.dh - 1)), (float)(p.sh - 1));
    }
    const float flx = floorf(fx), fly = floorf(fy);
    const int xi = (int)flx, yi = (int)fly;
    const float u = __fsub_rn(fx, flx), v = __fsub_rn(fy, fly);
    float out[CH];
    if (p.mode != 1) {
        const int x0 = min(max(xi, 0), p.sw - 1), x1 = min(max(xi + 1, 0), p.sw - 1);
        const uint8_t* r0 = src + (size_t)min(max(yi, 0), p.sh - 1) * p.src_pitch;
        const uint8_t* r1 = src + (size_t)min(max(yi + 1, 0), p.sh - 1) * p.src_pitch;
        const float omu = __fsub_rn(1.0f, u), omv = __fsub_rn(1.0f, v);
        const float w00 = __fmul_rn(omu, omv), w01 = __fmul_rn(u, omv), w10 = __fmul_rn(omu, v), w11 = __fmul_rn(u, v);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float acc = __fmul_rn(w00, unorm8(__ldg(r0 + x0 * CH + c)));
            acc = __fadd_rn(acc, __fmul_rn(w01, unorm8(__ldg(r0 + x1 * CH + c))));
            acc = __fadd_rn(acc, __fmul_rn(w10, unorm8(__ldg(r1 + x0 * CH + c))));
            acc = __fadd_rn(acc, __fmul_rn(w11, unorm8(__ldg(r1 + x1 * CH + c))));
            out[c] = acc;
        }
    } else {
        float xw[4], yw[4];
        cubic_weights(u, xw);
        cubic_weights(v, yw);
#pragma unroll
        for (int c = 0; c < CH; ++c) out[c] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint8_t* r = src + (size_t)min(max(yi - 1 + i, 0), p.sh - 1) * p.src_pitch;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = min(max(xi - 1 + j, 0), p.sw - 1);
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    out[c] = __fadd_rn(out[c], __fmul_rn(__fmul_rn(unorm8(__ldg(r + xx * CH + c)), xw[j]), yw[i]));
            }
        }
    }
    if (CH == 4) {
        *reinterpret_cast<uchar4*>(dst) = make_uchar4(to_unorm8(out[0]), to_unorm8(out[1 % CH]), to_unorm8(out[2 % CH]), to_unorm8(out[3 % CH]));
    } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) dst[c] = to_unorm8(out[c]);
    }
}

}  // namespace raisr
