// histeq.cuh -- histogram-equalisation kernels, SURVEY.md 8(f) row N4: the three OpenCL kernels of
// /root/reference/histeq/hist.cl behind clHistEq (/root/reference/histeq/eq_opencl.py:8-89).
//
//   hist_tiles_kernel     hist.cl:41-90    256-bin histogram of every 256x32 tile -> uint32 (h/32, w/256, 256)
//   lut_apply_kernel      hist.cl:92-102   out = mapping[in]                      (histeq_global)
//   lut_blend_kernel      hist.cl:104-147  bilinear blend of the four neighbouring block LUTs (fp32),
//                                          clamp to [0,255], truncate             (histeq_local_block)
// All three are byte streams with no arithmetic to speak of: HBM-bound.  The reference builds its tile
// histogram from 32 private ushort rows in LDS (hist.cl:57-85, tuned for AMD gfx902); here one CTA owns a
// tile, reads it with 16-byte loads and counts into per-warp shared-memory histograms (no inter-warp
// contention), then folds them -- integer work, bit-exact by construction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace raisr {

constexpr int kHistBins = 256;     // HIST_BINS,       eq_opencl.py:13
constexpr int kHistTileH = 32;     // HIST_THREAD_NUM, eq_opencl.py:14 (rows per tile); tile width = HIST_BINS

struct HistParams {
    const uint8_t* img; size_t pitch;
    uint32_t* hist;                 // (tiles_y, tiles_x, 256)
    int tiles_x, tiles_y;
};

__global__ void __launch_bounds__(256) hist_tiles_kernel(const HistParams p)
{
    __shared__ uint32_t sh[8][kHistBins];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * kHistBins; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int tx = blockIdx.x, ty = blockIdx.y;
    // 256 x 32 bytes = 512 chunks of 16 bytes; thread t takes chunks t and t+256 (rows t/16 and 16+t/16)
    const uint8_t* base = p.img + (size_t)ty * kHistTileH * p.pitch + (size_t)tx * kHistBins;
    const bool aligned = ((reinterpret_cast<uintptr_t>(base) | p.pitch) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int chunk = tid + 256 * k, row = chunk >> 4, col = (chunk & 15) * 16;
        const uint8_t* src = base + (size_t)row * p.pitch + col;
        uint32_t w[4];
        if (aligned) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[i] = (uint32_t)__ldg(src + 4 * i) | ((uint32_t)__ldg(src + 4 * i + 1) << 8) |
                       ((uint32_t)__ldg(src + 4 * i + 2) << 16) | ((uint32_t)__ldg(src + 4 * i + 3) << 24);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) atomicAdd(&sh[warp][(w[i] >> (8 * b)) & 0xffu], 1u);
    }
    __syncthreads();
    uint32_t acc = 0;
#pragma unroll
    for (int wgt = 0; wgt < 8; ++wgt) acc += sh[wgt][tid];
    p.hist[((size_t)ty * p.tiles_x + tx) * kHistBins + tid] = acc;
}

struct LutParams {
    const uint8_t* src; size_t src_pitch;
    uint8_t* dst; size_t dst_pitch;
    int w, h;
    const uint8_t* mapping;         // 256 bytes (histeq_global)
    const float* grid;              // (ny, nx, 256) floats (histeq_local_block)
    int block_w, block_h, nx, ny;
};

__global__ void __launch_bounds__(256) lut_apply_kernel(const LutParams p)
{
    __shared__ uint8_t lut[kHistBins];
    lut[threadIdx.x] = p.mapping[threadIdx.x];
    __syncthreads();
    const int y = blockIdx.y;
    const uint8_t* s = p.src + (size_t)y * p.src_pitch;
    uint8_t* d = p.dst + (size_t)y * p.dst_pitch;
    const bool aligned = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
    for (int x0 = (blockIdx.x * 256 + threadIdx.x) * 16; x0 < p.w; x0 += gridDim.x * 256 * 16) {
        if (aligned && x0 + 16 <= p.w) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(s + x0));
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[i] = (uint32_t)lut[w[i] & 0xff] | ((uint32_t)lut[(w[i] >> 8) & 0xff] << 8) |
                       ((uint32_t)lut[(w[i] >> 16) & 0xff] << 16) | ((uint32_t)lut[w[i] >> 24] << 24);
            *reinterpret_cast<uint4*>(d + x0) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int x = x0; x < min(x0 + 16, p.w); ++x) d[x] = lut[__ldg(s + x)];
        }
    }
}

__global__ void __launch_bounds__(256) lut_blend_kernel(const LutParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.w) return;
    // hist.cl:115-137: C integer division truncates toward zero, so pixels left of / above the first
    // block centre use block 0 with weight clamped to 0
    const int b00idx = (x - p.block_w / 2) / p.block_w, b00idy = (y - p.block_h / 2) / p.block_h;
    const int b00x = b00idx * p.block_w + p.block_w / 2, b00y = b00idy * p.block_h + p.block_h / 2;
    const int b01idx = min(b00idx + 1, p.nx - 1), b10idy = min(b00idy + 1, p.ny - 1);
    float s = __fdiv_rn((float)(x - b00x), (float)p.block_w), t = __fdiv_rn((float)(y - b00y), (float)p.block_h);
    s = fminf(fmaxf(s, 0.0f), 1.0f);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const int v = __ldg(p.src + (size_t)y * p.src_pitch + x);
    const float f00 = __ldg(p.grid + ((size_t)b00idy * p.nx + b00idx) * kHistBins + v);
    const float f01 = __ldg(p.grid + ((size_t)b00idy * p.nx + b01idx) * kHistBins + v);
    const float f10 = __ldg(p.grid + ((size_t)b10idy * p.nx + b00idx) * kHistBins + v);
    const float f11 = __ldg(p.grid + ((size_t)b10idy * p.nx + b01idx) * kHistBins + v);
    const float oms = __fsub_rn(1.0f, s), omt = __fsub_rn(1.0f, t);
    float acc = __fmul_rn(__fmul_rn(oms, omt), f00);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(s, omt), f01));
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(oms, t), f10));
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(s, t), f11));
    acc = fminf(fmaxf(acc, 0.0f), 255.0f);
    p.dst[(size_t)y * p.dst_pitch + x] = (uint8_t)acc;     // float -> uchar conversion truncates (hist.cl:144)
}

}  // namespace raisr
