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

#include "packed_f32.cuh"

namespace raisr {

constexpr int kHistBins = 256;     // HIST_BINS,       eq_opencl.py:13
constexpr int kHistTileH = 32;     // HIST_THREAD_NUM, eq_opencl.py:14 (rows per tile); tile width = HIST_BINS

struct HistParams {
    const uint8_t* img; size_t pitch;
    uint32_t* hist;                 // (tiles_y, tiles_x, 256)
    int tiles_x, tiles_y;
};

// One CTA per tile.  The counting rate is set by the shared-memory atomic unit, and with one 256-entry
// histogram per warp most of its time goes to bank collisions of unrelated bins (bank = bin % 32: 3.7
// wavefronts per warp-atomic on random bytes).  The CTA therefore keeps 16 interleaved copies,
// counter (bin, lane % 16) at word bin*16 + lane%16, so the bank of an update is 16*(bin & 1) + lane%16:
// only lanes l and l+16 can collide (~1.5 wavefronts per warp-atomic), and the 16-copy fold per tile costs
// less than it saves.  Equal bytes are merged before they reach the atomic unit: a 16-byte chunk of one
// value costs one atomic instead of sixteen, which keeps flat regions (sky, saturation) as fast as noise.
constexpr int kHistCopies = 16;

__global__ void __launch_bounds__(256) hist_tiles_kernel(const HistParams p)
{
    __shared__ __align__(16) uint32_t sh[kHistBins * kHistCopies];   // 16 KB
    const int tid = threadIdx.x;
    {
        uint4* z = reinterpret_cast<uint4*>(sh);
#pragma unroll
        for (int i = 0; i < kHistBins * kHistCopies / 4 / 256; ++i) z[tid + 256 * i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int tx = blockIdx.x, ty = blockIdx.y;
    // 256 x 32 bytes = 512 chunks of 16 bytes; thread t takes chunks t and t+256 (rows t/16 and 16+t/16)
    const uint8_t* base = p.img + (size_t)ty * kHistTileH * p.pitch + (size_t)tx * kHistBins;
    const bool aligned = ((reinterpret_cast<uintptr_t>(base) | p.pitch) & 15) == 0;
    uint32_t w[2][4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int chunk = tid + 256 * k, row = chunk >> 4, col = (chunk & 15) * 16;
        const uint8_t* src = base + (size_t)row * p.pitch + col;
        if (aligned) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
            w[k][0] = v.x; w[k][1] = v.y; w[k][2] = v.z; w[k][3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[k][i] = (uint32_t)__ldg(src + 4 * i) | ((uint32_t)__ldg(src + 4 * i + 1) << 8) |
                          ((uint32_t)__ldg(src + 4 * i + 2) << 16) | ((uint32_t)__ldg(src + 4 * i + 3) << 24);
        }
    }
    uint32_t* mine = sh + (tid & (kHistCopies - 1));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t w0 = w[k][0];
        if (w0 == __byte_perm(w0, 0, 0) && w[k][1] == w0 && w[k][2] == w0 && w[k][3] == w0) {
            atomicAdd(mine + (w0 & 0xffu) * kHistCopies, 16u);
            continue;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t v = w[k][i];
            if (v == __byte_perm(v, 0, 0)) {
                atomicAdd(mine + (v & 0xffu) * kHistCopies, 4u);
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) atomicAdd(mine + ((v >> (8 * b)) & 0xffu) * kHistCopies, 1u);
            }
        }
    }
    __syncthreads();
    // fold: thread = bin; its 16 copies are 64 contiguous bytes, read as four 16-byte pieces starting at a
    // lane-dependent piece so that the eight lanes of a quarter warp touch different bank groups
    const uint4* row = reinterpret_cast<const uint4*>(sh + tid * kHistCopies);
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 v = row[(i + (tid >> 1)) & 3];
        acc += v.x + v.y + v.z + v.w;
    }
    p.hist[((size_t)ty * p.tiles_x + tx) * kHistBins + tid] = acc;
}

struct LutParams {
    const uint8_t* src; size_t src_pitch;
    uint8_t* dst; size_t dst_pitch;
    int w, h;
    const uint8_t* mapping;         // 256 bytes (histeq_global)
    const float* grid;              // (ny, nx, 256) floats (histeq_local_block)
    int block_w, block_h, nx, ny;
    int cells_x, cells_y, strips;   // lut_blend_kernel work decomposition
};

// out = mapping[in].  A byte gather from a 256-entry shared table is bank-conflict bound (32 random lanes
// over 32 banks: ~4 wavefronts per lookup, measured 4.4 TB/s); the table is therefore replicated once per
// lane (word [v][lane], 32 KB) so that every lookup is a single conflict-free wavefront, and the CTAs are
// persistent so the replication is paid once.
__global__ void __launch_bounds__(256) lut_apply_kernel(const LutParams p)
{
    extern __shared__ uint32_t lutw[];                       // [256][32]
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < kHistBins * 32; i += 256) lutw[i] = p.mapping[i >> 5];
    __syncthreads();
    const uint32_t* mylut = lutw + lane;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.src) | reinterpret_cast<uintptr_t>(p.dst) | p.src_pitch | p.dst_pitch) & 15) == 0;
    const int cpr = (p.w + 15) >> 4;                          // 16-byte chunks per row
    const long long nchunks = (long long)cpr * p.h;
    auto map16 = [&](uint4 v) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t a = mylut[(w[i] & 0xffu) << 5], b = mylut[((w[i] >> 8) & 0xffu) << 5];
            const uint32_t c = mylut[((w[i] >> 16) & 0xffu) << 5], e = mylut[(w[i] >> 24) << 5];
            w[i] = __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, e, 0x0040), 0x5410);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    // chunk j = (row y, chunk c); j advances by `step`, tracked as (y, c) so the loop needs no division
    const unsigned step = gridDim.x * 256u, step_y = step / (unsigned)cpr, step_c = step % (unsigned)cpr;
    unsigned first = blockIdx.x * 256u + tid;
    if ((long long)first >= nchunks) return;
    int y = (int)(first / (unsigned)cpr), c = (int)(first % (unsigned)cpr);
    auto advance = [&]() { y += step_y; c += step_c; if (c >= cpr) { c -= cpr; ++y; } };
    if (aligned && (p.w & 15) == 0) {          // whole rows of 16-byte chunks: four loads in flight per thread
        while (y < p.h) {
            uint4 v[4];
            size_t dof[4];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ok[k] = y < p.h;
                if (ok[k]) {
                    dof[k] = (size_t)y * p.dst_pitch + 16 * c;
                    v[k] = __ldg(reinterpret_cast<const uint4*>(p.src + (size_t)y * p.src_pitch + 16 * c));
                }
                advance();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (ok[k]) *reinterpret_cast<uint4*>(p.dst + dof[k]) = map16(v[k]);
        }
        return;
    }
    for (; y < p.h; advance()) {
        const int x0 = 16 * c;
        const uint8_t* s = p.src + (size_t)y * p.src_pitch;
        uint8_t* d = p.dst + (size_t)y * p.dst_pitch;
        if (aligned && x0 + 16 <= p.w) {
            *reinterpret_cast<uint4*>(d + x0) = map16(__ldg(reinterpret_cast<const uint4*>(s + x0)));
        } else {
            for (int x = x0; x < min(x0 + 16, p.w); ++x) d[x] = (uint8_t)mylut[(uint32_t)__ldg(s + x) << 5];
        }
    }
}

// hist.cl:104-147.  Between four block centres ("cell") every pixel blends the same four mappings, so a CTA
// takes a 64-row strip of one cell, interleaves those four tables as float4 per grey level and replicates
// them over the eight 16-byte bank groups ([v][lane % 8], 32 KB): one conflict-free LDS.128 per pixel
// instead of four global gathers.  Cells: index k covers [k ? k*bw + bw/2 : 0, k == n-1 ? w : (k+1)*bw + bw/2)
// -- exactly the pixels for which the reference's truncating division yields b00idx == k.
constexpr int kBlendStripRows = 64;

__device__ __forceinline__ float blend_px(const float4 f, float w00, float w01, float w10, float w11)
{
    float acc = __fmul_rn(w00, f.x);
    acc = __fadd_rn(acc, __fmul_rn(w01, f.y));
    acc = __fadd_rn(acc, __fmul_rn(w10, f.z));
    acc = __fadd_rn(acc, __fmul_rn(w11, f.w));
    // clamp, then truncate: adding 2^23 with round-toward-zero leaves floor(acc) in the low mantissa bits
    // (acc is in [0,255]) -- a full-rate FADD instead of the quarter-rate F2I
    acc = fminf(fmaxf(acc, 0.0f), 255.0f);
    return __fadd_rz(acc, 8388608.0f);
}
__device__ __forceinline__ uint32_t blend_byte(float r) { return __float_as_uint(r) & 0xffu; }

// Same value as blend_px(f, oms*omt, s*omt, oms*t, s*t) with the eight products issued as four packed FMUL2
// (os = (1-s, s) of the pixel's column, tt = (t, 1-t) of its row); the three additions stay scalar -- ptxas would
// contract a packed multiply feeding a packed add into one FFMA2 and drop a rounding.
__device__ __forceinline__ float blend_px_packed(const float4 f, p2 os, float t, float omt)
{
    float p00, p01, p10, p11;
    upk(mul2(mul2(os, bc(omt)), pk(f.x, f.y)), p00, p01);
    upk(mul2(mul2(os, bc(t)), pk(f.z, f.w)), p10, p11);
    float acc = __fadd_rn(p00, p01);
    acc = __fadd_rn(acc, p10);
    acc = __fadd_rn(acc, p11);
    acc = fminf(fmaxf(acc, 0.0f), 255.0f);
    return __fadd_rz(acc, 8388608.0f);
}

__global__ void __launch_bounds__(256) lut_blend_kernel(const LutParams p)
{
    extern __shared__ float4 lut4[];                         // [256][8] replicated + [256] staging
    float4* stage = lut4 + kHistBins * 8;
    __shared__ float2 trow[kBlendStripRows];                 // (t, 1-t) of every row of the strip
    const int tid = threadIdx.x;
    int job = blockIdx.x;
    const int strip = job % p.strips; job /= p.strips;
    const int kx = job % p.cells_x, ky = job / p.cells_x;
    const int x_lo = kx ? kx * p.block_w + p.block_w / 2 : 0;
    const int x_hi = (kx == p.cells_x - 1) ? p.w : (kx + 1) * p.block_w + p.block_w / 2;
    const int y_lo = ky ? ky * p.block_h + p.block_h / 2 : 0;
    const int y_hi = (ky == p.cells_y - 1) ? p.h : (ky + 1) * p.block_h + p.block_h / 2;
    const int r_lo = y_lo + strip * kBlendStripRows, r_hi = min(r_lo + kBlendStripRows, y_hi);
    if (r_lo >= r_hi || x_lo >= x_hi) return;
    {
        const int kx1 = min(kx + 1, p.nx - 1), ky1 = min(ky + 1, p.ny - 1);
        const float* g = p.grid;
        stage[tid] = make_float4(__ldg(g + ((size_t)ky * p.nx + kx) * kHistBins + tid),
                                 __ldg(g + ((size_t)ky * p.nx + kx1) * kHistBins + tid),
                                 __ldg(g + ((size_t)ky1 * p.nx + kx) * kHistBins + tid),
                                 __ldg(g + ((size_t)ky1 * p.nx + kx1) * kHistBins + tid));
        if (tid < kBlendStripRows) {
            const int cy0 = ky * p.block_h + p.block_h / 2;
            const float t = fminf(fmaxf(__fdiv_rn((float)(r_lo + tid - cy0), (float)p.block_h), 0.0f), 1.0f);
            trow[tid] = make_float2(t, __fsub_rn(1.0f, t));
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) lut4[tid + 256 * r] = stage[(tid + 256 * r) >> 3];
    }
    __syncthreads();
    const float4* mylut = lut4 + (tid & 7);
    const int cx = kx * p.block_w + p.block_w / 2;
    const float fbw = (float)p.block_w;
    const int tx = tid & 31, ty = tid >> 5;
    // 8-pixel groups aligned to 8 bytes in the image; head/tail pixels of the cell are masked
    const bool vec = ((reinterpret_cast<uintptr_t>(p.src) | reinterpret_cast<uintptr_t>(p.dst) | p.src_pitch | p.dst_pitch) & 7) == 0;
    const int g_lo = x_lo & ~7;
    for (int gx = g_lo + 8 * tx; gx < x_hi; gx += 256) {
        float s[8], oms[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i] = fminf(fmaxf(__fdiv_rn((float)(gx + i - cx), fbw), 0.0f), 1.0f);
            oms[i] = __fsub_rn(1.0f, s[i]);
        }
        const bool full = vec && gx >= x_lo && gx + 8 <= x_hi;
        const uint8_t* srow = p.src + (size_t)(r_lo + ty) * p.src_pitch + gx;
        uint8_t* drow = p.dst + (size_t)(r_lo + ty) * p.dst_pitch + gx;
        if (full) {
            // four rows (8 apart) per batch: their loads are all in flight before the first blend
            for (int y0 = r_lo + ty; y0 < r_hi; y0 += 32, srow += 32 * p.src_pitch, drow += 32 * p.dst_pitch) {
                uint2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (y0 + 8 * k < r_hi) v[k] = __ldg(reinterpret_cast<const uint2*>(srow + (size_t)(8 * k) * p.src_pitch));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (y0 + 8 * k >= r_hi) break;
                    const float2 tt = trow[y0 + 8 * k - r_lo];
                    const float t = tt.x, omt = tt.y;
                    const uint32_t vin[2] = {v[k].x, v[k].y};
                    uint32_t o[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t b[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = 4 * h + i;
                            const float4 f = mylut[((vin[h] >> (8 * i)) & 0xffu) << 3];
                            b[i] = __float_as_uint(blend_px_packed(f, pk(oms[c], s[c]), t, omt));
                        }
                        // low mantissa byte of each result = the truncated value (hist.cl:144)
                        o[h] = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
                    }
                    *reinterpret_cast<uint2*>(drow + (size_t)(8 * k) * p.dst_pitch) = make_uint2(o[0], o[1]);
                }
            }
        } else {
            for (int y = r_lo + ty; y < r_hi; y += 8, srow += 8 * p.src_pitch, drow += 8 * p.dst_pitch) {
                const float2 tt = trow[y - r_lo];
                const float t = tt.x, omt = tt.y;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int x = gx + i;
                    if (x < x_lo || x >= x_hi) continue;
                    const float4 f = mylut[(uint32_t)__ldg(srow + i) << 3];
                    drow[i] = (uint8_t)blend_byte(blend_px(f, __fmul_rn(oms[i], omt), __fmul_rn(s[i], omt), __fmul_rn(oms[i], t), __fmul_rn(s[i], t)));
                }
            }
        }
    }
}

}  // namespace raisr
