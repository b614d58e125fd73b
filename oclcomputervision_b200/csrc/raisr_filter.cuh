// raisr_filter.cuh -- kernel B of the RAISR path: per-pixel 11x11 gather-dot against the hashed filter.
//
// Restates /root/reference/super_resolution/raisr.cl:316-337 (filter lookup, 121-tap dot, saturating
// store) in fp32.  The reference reads the taps from __global with per-lane-divergent addresses
// (raisr.cl:317,328); here the table slice of ONE pixel type (n_buckets x 121 fp32 = 104.5 KB for
// 24x3x3) is resident in shared memory and every CTA is persistent and bound to one pixel type
// (type = blockIdx.x % S*S), so the only HBM/L2 streams are the upscaled tile, one hash byte per
// pixel and the output.
//
// v1 mapping ("block"): one thread owns BR x BC same-type output pixels (stride S in the dense
// image), streams the rows of their joint patch region through registers and fetches the taps of
// each (pixel, filter row) with three 128-bit shared loads (filter rows padded to 12 floats).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace raisr {

constexpr int kFlen = 11;
constexpr int kTaps = kFlen * kFlen;
constexpr int kRowPad = 12;                    // floats per padded filter row
constexpr int kFStride = kFlen * kRowPad;      // 132 floats per filter (33 x 16 B: odd -> banks spread)

// tap formats of the table slice resident in shared memory (raisr_octet.cuh)
enum { kTapsF32 = 0, kTapsF16 = 1, kTapsB24 = 2 };

struct FilterParams {
    const float* uext;        // (own_rows*S + 10) rows per frame, column-major, see raisr_prep.cuh
    size_t uext_pitch;        // floats per image column
    int uext_cols;            // valid columns (dw + 10)
    size_t uext_frame_stride; // floats
    int uext_rows;            // valid rows in uext per frame
    const uint8_t* hash;      // planar by pixel type
    size_t hash_pitch, hash_plane_stride, hash_frame_stride;
    const float* table;       // [type][n_buckets][kFStride] (block layout)
    int n_buckets;
    void* dst;                // dense output, first row = first output row of this launch
    size_t dst_pitch, dst_frame_stride;  // bytes
    int ow, oh;               // own columns (= source width) / own rows of this launch (= rows / S)
    int n_frames;
    int tiles_x, tiles_y;
    int raw_f32;              // float output is stored unclamped (colour path: CSC back happens before saturation)
    // "eigen_in_filter": the filter kernel reads the structure tensor (three float planes, geometry of the hash image,
    // see PrepParams::tens) and does the eigen-solve / hash itself
    const float* tens;
    size_t tens_plane_stride;
    float sq[2], cq[2];
    int n_angle, n_strength, n_coherence, as_written;
};

template <int S, int OTW, int OTH, int BR, int BC>
struct BlockCfg {
    static constexpr int NTX = OTW / BC, NTY = OTH / BR, NT = NTX * NTY;
    static constexpr int TUH = S * (OTH - 1) + kFlen;                       // tile rows of U
    static constexpr int TUW = ((S * (OTW - 1) + kFlen + (S - 1)) + 3) / 4 * 4;  // tile cols (floats)
    static constexpr int NV = ((S - 1) + S * (BC - 1) + kFlen + 3) / 4;     // float4 per thread per row
    static constexpr int RR = S * (BR - 1) + kFlen;                         // region rows per thread
    static_assert((S * BC) % 4 == 0, "thread column origin must stay 16-byte aligned");
    static_assert(OTW % BC == 0 && OTH % BR == 0, "tile must be a whole number of thread blocks");
};

__device__ __forceinline__ void store_px(uint8_t* p, float v)
{
    v = fminf(fmaxf(v, 0.0f), 1.0f);  // fmaxf(NaN,0)=0: NaN saturates to 0 like write_imagef
    *p = (uint8_t)__float2uint_rn(v * 255.0f);
}
__device__ __forceinline__ void store_px(float* p, float v) { *p = fminf(fmaxf(v, 0.0f), 1.0f); }

template <int S, int OTW, int OTH, int BR, int BC, int PX, typename OutT>
__device__ __forceinline__ void block_tile_body(const FilterParams& p, const float* tab, const float* ut,
                                                int frame, int type, int py, int oy0, int ox0)
{
    using C = BlockCfg<S, OTW, OTH, BR, BC>;
    const int lx = threadIdx.x % C::NTX, ly = threadIdx.x / C::NTX;
    // bucket -> tap offset (floats) for the BR x BC own pixels
    int toff[BR][BC];
    const uint8_t* hp = p.hash + (size_t)frame * p.hash_frame_stride + (size_t)type * p.hash_plane_stride;
#pragma unroll
    for (int k = 0; k < BR; ++k) {
        int oy = min(oy0 + BR * ly + k, p.oh - 1);
        const uint8_t* hr = hp + (size_t)oy * p.hash_pitch + ox0 + BC * lx;
#pragma unroll
        for (int m = 0; m < BC; ++m) {
            int b = (ox0 + BC * lx + m < p.ow) ? (int)__ldg(hr + m) : 0;
            toff[k][m] = min(b, p.n_buckets - 1) * kFStride;
        }
    }
    float acc[BR][BC];
#pragma unroll
    for (int k = 0; k < BR; ++k)
#pragma unroll
        for (int m = 0; m < BC; ++m) acc[k][m] = 0.0f;

    const float4* ut4 = reinterpret_cast<const float4*>(ut);
#pragma unroll
    for (int rr = 0; rr < C::RR; ++rr) {
        float u[C::NV * 4];
        const float4* urow = ut4 + (size_t)(S * BR * ly + rr) * (C::TUW / 4) + (S * BC * lx) / 4;
#pragma unroll
        for (int n = 0; n < C::NV; ++n) {
            float4 t = urow[n];
            u[4 * n] = t.x; u[4 * n + 1] = t.y; u[4 * n + 2] = t.z; u[4 * n + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < BR; ++k) {
            const int i = rr - S * k;  // filter row of pixel-row k that meets region row rr
            if (i >= 0 && i < kFlen) {
#pragma unroll
                for (int m = 0; m < BC; ++m) {
                    const float4* t4 = reinterpret_cast<const float4*>(tab + toff[k][m] + i * kRowPad);
                    float4 t0 = t4[0], t1 = t4[1], t2 = t4[2];
                    float a = acc[k][m];
                    a = fmaf(u[PX + S * m + 0], t0.x, a);
                    a = fmaf(u[PX + S * m + 1], t0.y, a);
                    a = fmaf(u[PX + S * m + 2], t0.z, a);
                    a = fmaf(u[PX + S * m + 3], t0.w, a);
                    a = fmaf(u[PX + S * m + 4], t1.x, a);
                    a = fmaf(u[PX + S * m + 5], t1.y, a);
                    a = fmaf(u[PX + S * m + 6], t1.z, a);
                    a = fmaf(u[PX + S * m + 7], t1.w, a);
                    a = fmaf(u[PX + S * m + 8], t2.x, a);
                    a = fmaf(u[PX + S * m + 9], t2.y, a);
                    a = fmaf(u[PX + S * m + 10], t2.z, a);
                    acc[k][m] = a;
                }
            }
        }
    }
    OutT* dst = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(p.dst) + (size_t)frame * p.dst_frame_stride);
#pragma unroll
    for (int k = 0; k < BR; ++k) {
        int oy = oy0 + BR * ly + k;
        if (oy >= p.oh) continue;
        OutT* drow = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(dst) + (size_t)(S * oy + py) * p.dst_pitch);
#pragma unroll
        for (int m = 0; m < BC; ++m) {
            int ox = ox0 + BC * lx + m;
            if (ox < p.ow) store_px(drow + (S * ox + PX), acc[k][m]);
        }
    }
}

template <int S, int OTW, int OTH, int BR, int BC, typename OutT>
__global__ void __launch_bounds__(BlockCfg<S, OTW, OTH, BR, BC>::NT, 1) filter_block_kernel(const FilterParams p)
{
    using C = BlockCfg<S, OTW, OTH, BR, BC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);
    float* ut = tab + (size_t)p.n_buckets * kFStride;
    const int tid = threadIdx.x;
    const int ntypes = S * S;
    const int type = blockIdx.x % ntypes, worker = blockIdx.x / ntypes, nworkers = gridDim.x / ntypes;
    const int py = type / S, px = type % S;

    {   // resident table slice of this pixel type
        const float4* g = reinterpret_cast<const float4*>(p.table + (size_t)type * p.n_buckets * kFStride);
        float4* s = reinterpret_cast<float4*>(tab);
        for (int i = tid; i < p.n_buckets * (kFStride / 4); i += C::NT) s[i] = __ldg(g + i);
    }
    const int tiles_per_frame = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_per_frame * p.n_frames;
    for (int tile = worker; tile < ntiles; tile += nworkers) {
        const int frame = tile / tiles_per_frame;
        const int rem = tile - frame * tiles_per_frame;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int oy0 = ty * OTH, ox0 = tx * OTW;
        // U tile: extended rows S*oy0+py .., extended cols from (S*ox0+px) rounded down to 4
        const int er0 = S * oy0 + py;
        const int ec0 = (S * ox0 + px) & ~3;
        const float* ug = p.uext + (size_t)frame * p.uext_frame_stride;
        __syncthreads();  // previous tile fully consumed (also orders the table fill)
        for (int idx = tid; idx < C::TUH * C::TUW; idx += C::NT) {
            int r = idx / C::TUW, c = idx - r * C::TUW;
            int gr = min(er0 + r, p.uext_rows - 1);
            int gc = min(ec0 + c, p.uext_cols - 1);
            ut[idx] = __ldg(ug + (size_t)gc * p.uext_pitch + gr);
        }
        __syncthreads();
        if (S == 2) {
            if (px == 0) block_tile_body<S, OTW, OTH, BR, BC, 0, OutT>(p, tab, ut, frame, type, py, oy0, ox0);
            else block_tile_body<S, OTW, OTH, BR, BC, (S > 1 ? 1 : 0), OutT>(p, tab, ut, frame, type, py, oy0, ox0);
        } else if (S == 3) {
            if (px == 0) block_tile_body<S, OTW, OTH, BR, BC, 0, OutT>(p, tab, ut, frame, type, py, oy0, ox0);
            else if (px == 1) block_tile_body<S, OTW, OTH, BR, BC, (S > 1 ? 1 : 0), OutT>(p, tab, ut, frame, type, py, oy0, ox0);
            else block_tile_body<S, OTW, OTH, BR, BC, (S > 2 ? 2 : 0), OutT>(p, tab, ut, frame, type, py, oy0, ox0);
        } else {
            if (px == 0) block_tile_body<S, OTW, OTH, BR, BC, 0, OutT>(p, tab, ut, frame, type, py, oy0, ox0);
            else if (px == 1) block_tile_body<S, OTW, OTH, BR, BC, (S > 1 ? 1 : 0), OutT>(p, tab, ut, frame, type, py, oy0, ox0);
            else if (px == 2) block_tile_body<S, OTW, OTH, BR, BC, (S > 2 ? 2 : 0), OutT>(p, tab, ut, frame, type, py, oy0, ox0);
            else block_tile_body<S, OTW, OTH, BR, BC, (S > 3 ? 3 : 0), OutT>(p, tab, ut, frame, type, py, oy0, ox0);
        }
    }
}

}  // namespace raisr
