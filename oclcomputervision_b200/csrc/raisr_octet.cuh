// raisr_octet.cuh -- kernel B, "octet" mapping: eight lanes cooperate on one output pixel.
//
// Same contract as filter_block_kernel (raisr_filter.cuh): restates the filter lookup, 121-tap dot
// and saturating store of /root/reference/super_resolution/raisr.cl:316-337 in fp32 with the table
// slice of one pixel type resident in shared memory.
//
// Why eight lanes per pixel.  Every pixel needs its own 121 fp32 taps, so the kernel is bound by
// the 128 B/clk shared-memory pipe, not by FFMA.  B200 resolves a 128-bit shared load in rigid
// quarter-warp phases (measured, tools/microbench.cu: random per-lane filter rows cost 10.3
// cycles per LDS.128 instead of 4), so a lane-per-pixel gather wastes ~60% of that pipe on bank
// conflicts.  Here each filter occupies one 512-byte, 128-byte-aligned record of 32 chunks
// (16 B each) and lane p of an octet only ever reads chunks p, p+8, p+16, p+24: the eight lanes
// of a quarter-warp hit eight different bank groups whatever the hashes are -- conflict-free by
// construction.  Each lane owns 16 fixed taps (one full filter row plus a short run of rows 8-10),
// keeps the matching patch values in a register window that slides along the output row (S new
// values per run and pixel), and eight pixels' partial sums are combined with a 7-shuffle
// transposing butterfly.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "raisr_filter.cuh"

namespace raisr {

constexpr int kOctStride = 128;  // floats per filter record (512 B)

// Host-side packing of one 11x11 filter (row-major taps f[i*11+j]) into the lane-major record:
// record float (p + 8n)*4 + c  <->  slot s = 4n + c of lane p.
//   slots 0..10  : filter row p, columns 0..10            (p = 0..7)
//   slots 11..15 : lanes 0..5: row 8 + p/2, columns 5*(p%2) .. +4
//                  lanes 6,7 : the three leftover taps (8,10) (9,10) (10,10) in the slots that are
//                              freshly loaded for every pixel (the last min(S,5) slots); which
//                              ones depends on S, so the record is packed per scale.
inline void octet_single_slot(int S, int idx, int* lane, int* slot)
{
    const int newp = S < 5 ? S : 5;  // fresh slots per pixel in the 5-run: slots 16-newp .. 15
    *lane = 6 + idx / newp;
    *slot = 16 - newp + idx % newp;
}

inline void octet_pack_filter_s(const float* f, float* rec, int S)
{
    for (int i = 0; i < kOctStride; ++i) rec[i] = 0.0f;
    auto put = [&](int lane, int slot, float v) { rec[(lane + 8 * (slot / 4)) * 4 + (slot % 4)] = v; };
    for (int p = 0; p < 8; ++p)
        for (int j = 0; j < kFlen; ++j) put(p, j, f[p * kFlen + j]);
    for (int p = 0; p < 6; ++p)
        for (int t = 0; t < 5; ++t) put(p, 11 + t, f[(8 + p / 2) * kFlen + 5 * (p % 2) + t]);
    for (int idx = 0; idx < 3; ++idx) {
        int lane, slot;
        octet_single_slot(S, idx, &lane, &slot);
        put(lane, slot, f[(8 + idx) * kFlen + 10]);
    }
}

template <int S>
struct OctetCfg;
// OTW x OTH own pixels per tile, items of IW pixels (one octet each), NT threads, TUW = odd row
// pitch of the U tile in floats chosen so the per-pixel LDS.32 of a warp (4 octets on consecutive
// own rows) are bank-conflict-free for the 11-runs and nearly so for the 5-runs.
template <>
struct OctetCfg<2> { static constexpr int OTW = 64, OTH = 32, IW = 32, NT = 512, TUW = 153; };
template <>
struct OctetCfg<3> { static constexpr int OTW = 64, OTH = 16, IW = 16, NT = 512, TUW = 203; };
template <>
struct OctetCfg<4> { static constexpr int OTW = 32, OTH = 16, IW = 16, NT = 256, TUW = 149; };

template <int S>
struct OctetGeom {
    using C = OctetCfg<S>;
    static constexpr int TUH = S * (C::OTH - 1) + kFlen;     // tile rows
    static constexpr int NCOLS = S * (C::OTW - 1) + kFlen;   // tile columns actually used
    static constexpr int TUW = C::TUW;
    static constexpr int NEWF = S;                 // fresh values per pixel in the 11-run
    static constexpr int NEWP = S < 5 ? S : 5;     // fresh values per pixel in the 5-run
    static constexpr int SEGS = C::OTW / C::IW;
    static constexpr int ITEMS = C::OTH * SEGS;
    static constexpr int NOCT = C::NT / 8;
    static constexpr int TILE_FLOATS = (TUH * TUW + 3) / 4 * 4;
    static_assert(C::IW % 16 == 0 && C::OTW % C::IW == 0, "items are whole batches of 8 pixels, hashes come 16 at a time");
    static_assert(TUW >= NCOLS && (TUW & 1), "odd pitch that holds a tile row");
};

template <int S>
inline size_t octet_smem_bytes(int n_buckets)
{
    using G = OctetGeom<S>;
    return ((size_t)n_buckets * kOctStride + 2 * (size_t)G::TILE_FLOATS) * sizeof(float);
}

__device__ __forceinline__ void cp_async4(unsigned smem_addr, const float* gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Asynchronous fill of one U tile: rows er0.., columns ec0.. of the extended upscaled frame.
template <int S>
__device__ __forceinline__ void octet_issue_tile(const FilterParams& p, float* buf, int frame, int er0, int ec0)
{
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* ug = p.uext + (size_t)frame * p.uext_frame_stride;
    const int maxc = (int)p.uext_pitch - 1 - ec0;
    for (int r = warp; r < G::TUH; r += C::NT / 32) {
        const float* grow = ug + (size_t)min(er0 + r, p.uext_rows - 1) * p.uext_pitch + ec0;
        const unsigned srow = (unsigned)__cvta_generic_to_shared(buf + r * G::TUW);
#pragma unroll
        for (int c = lane; c < G::NCOLS; c += 32) cp_async4(srow + 4u * c, grow + min(c, maxc));
    }
}

template <int S, typename OutT>
__global__ void __launch_bounds__(OctetCfg<S>::NT, 1) filter_octet_kernel(const FilterParams p)
{
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);                 // 512-byte records, 128-B aligned
    float* ubuf0 = tab + (size_t)p.n_buckets * kOctStride;
    float* ubuf1 = ubuf0 + G::TILE_FLOATS;
    const int tid = threadIdx.x;
    const int ntypes = S * S;
    const int type = blockIdx.x % ntypes, worker = blockIdx.x / ntypes, nworkers = gridDim.x / ntypes;
    const int py = type / S, px = type % S;
    const int lane8 = tid & 7, octet = tid >> 3;
    if ((__cvta_generic_to_shared(tab) & 127) != 0) __trap();  // records must be 128-byte aligned

    const int tiles_per_frame = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_per_frame * p.n_frames;
    auto tile_coords = [&](int tile, int& frame, int& oy0, int& ox0) {
        frame = tile / tiles_per_frame;
        const int rem = tile - frame * tiles_per_frame;
        const int ty = rem / p.tiles_x;
        oy0 = ty * C::OTH;
        ox0 = (rem - ty * p.tiles_x) * C::OTW;
    };
    if (worker < ntiles) {   // first tile in flight while the table slice is copied
        int f, oy0, ox0;
        tile_coords(worker, f, oy0, ox0);
        octet_issue_tile<S>(p, ubuf0, f, S * oy0 + py, S * ox0 + px);
    }
    cp_async_commit();
    {
        const float4* g = reinterpret_cast<const float4*>(p.table + (size_t)type * p.n_buckets * kOctStride);
        float4* s = reinterpret_cast<float4*>(tab);
        for (int i = tid; i < p.n_buckets * (kOctStride / 4); i += C::NT) s[i] = __ldg(g + i);
    }

    // Lane geometry (floats relative to the patch origin of the current pixel in the U tile).
    const int off_full = lane8 * G::TUW;  // filter row = lane8, column 0
    int off_part[G::NEWP];                // addresses of the freshly loaded 5-run slots
    if (lane8 < 6) {
#pragma unroll
        for (int t = 0; t < G::NEWP; ++t)
            off_part[t] = (8 + lane8 / 2) * G::TUW + 5 * (lane8 % 2) + (5 - G::NEWP) + t;
    } else {
#pragma unroll
        for (int t = 0; t < G::NEWP; ++t) {
            int idx = (lane8 - 6) * G::NEWP + t;   // leftover taps (8,10) (9,10) (10,10)
            off_part[t] = (idx < 3 ? (8 + idx) : 10) * G::TUW + 10;
        }
    }
    const float4* tab_lane = reinterpret_cast<const float4*>(tab) + lane8;
    const unsigned omask = 0xffu << (tid & 24);  // the eight lanes of this octet

    int it = 0;
    for (int tile = worker; tile < ntiles; tile += nworkers, ++it) {
        int frame, oy0, ox0;
        tile_coords(tile, frame, oy0, ox0);
        float* ut = (it & 1) ? ubuf1 : ubuf0;
        if (tile + nworkers < ntiles) {   // prefetch the next tile into the other buffer
            int f2, oy2, ox2;
            tile_coords(tile + nworkers, f2, oy2, ox2);
            octet_issue_tile<S>(p, (it & 1) ? ubuf0 : ubuf1, f2, S * oy2 + py, S * ox2 + px);
        }
        cp_async_commit();
        cp_async_wait<1>();   // everything but the newest group (the prefetch) has landed
        __syncthreads();

        const uint8_t* hplane = p.hash + (size_t)frame * p.hash_frame_stride + (size_t)type * p.hash_plane_stride;
        OutT* dst = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(p.dst) + (size_t)frame * p.dst_frame_stride);

        for (int item = octet; item < G::ITEMS; item += G::NOCT) {
            const int seg = item / C::OTH, row = item - seg * C::OTH;   // a warp = 4 consecutive rows
            const int oy = oy0 + row;
            const int oxs = ox0 + seg * C::IW;              // first own column of the item
            if (oy >= p.oh || oxs >= p.ow) continue;         // octet-uniform
            const uint8_t* hrow = hplane + (size_t)oy * p.hash_pitch + oxs;
            uint4 hq[C::IW / 16];
#pragma unroll
            for (int i = 0; i < C::IW / 16; ++i) hq[i] = __ldg(reinterpret_cast<const uint4*>(hrow) + i);
            // patch origin of own pixel (row, seg*IW) in the tile
            const float* base = ut + (S * row) * G::TUW + S * (seg * C::IW);
            const float* pf = base + off_full;
            float w11[kFlen], w5[5];
            // windows primed for the virtual pixel one step to the left
#pragma unroll
            for (int j = G::NEWF; j < kFlen; ++j) w11[j] = pf[j - S];
#pragma unroll
            for (int t = 0; t < 5; ++t) w5[t] = 0.0f;
            if (lane8 < 6) {
                const float* pp = base + (8 + lane8 / 2) * G::TUW + 5 * (lane8 % 2);
#pragma unroll
                for (int t = G::NEWP; t < 5; ++t) w5[t] = pp[t - S];
            }
            OutT* drow = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(dst) + (size_t)(S * oy + py) * p.dst_pitch);
            unsigned prev_bucket = 0xffffffffu;
            float4 t0 = make_float4(0, 0, 0, 0), t1 = t0, t2 = t0, t3 = t0;

#pragma unroll
            for (int b0 = 0; b0 < C::IW; b0 += 8) {
                if (oxs + b0 >= p.ow) break;                  // octet-uniform
                const uint4 hv = hq[b0 / 16];
                const unsigned hlo = (b0 & 8) ? hv.z : hv.x, hhi = (b0 & 8) ? hv.w : hv.y;
                float acc[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int pix = b0 + b;                   // pixel within the item
                    // slide the windows by S and fetch the fresh patch values
#pragma unroll
                    for (int j = 0; j < kFlen - G::NEWF; ++j) w11[j] = w11[j + S];
#pragma unroll
                    for (int t = 0; t < G::NEWF; ++t) w11[kFlen - G::NEWF + t] = pf[S * pix + kFlen - G::NEWF + t];
#pragma unroll
                    for (int t = 0; t < 5 - G::NEWP; ++t) w5[t] = w5[t + S];
#pragma unroll
                    for (int t = 0; t < G::NEWP; ++t) w5[5 - G::NEWP + t] = base[S * pix + off_part[t]];
                    unsigned bucket = ((b < 4 ? hlo : hhi) >> (8 * (b & 3))) & 0xffu;
                    bucket = min(bucket, (unsigned)(p.n_buckets - 1));
                    if (bucket != prev_bucket) {              // octet-uniform: neighbours often share a filter
                        const float4* tp = tab_lane + bucket * (kOctStride / 4);
                        t0 = tp[0]; t1 = tp[8]; t2 = tp[16]; t3 = tp[24];
                        prev_bucket = bucket;
                    }
                    float a0 = w11[0] * t0.x, a1 = w11[1] * t0.y;
                    a0 = fmaf(w11[2], t0.z, a0); a1 = fmaf(w11[3], t0.w, a1);
                    a0 = fmaf(w11[4], t1.x, a0); a1 = fmaf(w11[5], t1.y, a1);
                    a0 = fmaf(w11[6], t1.z, a0); a1 = fmaf(w11[7], t1.w, a1);
                    a0 = fmaf(w11[8], t2.x, a0); a1 = fmaf(w11[9], t2.y, a1);
                    a0 = fmaf(w11[10], t2.z, a0); a1 = fmaf(w5[0], t2.w, a1);
                    a0 = fmaf(w5[1], t3.x, a0); a1 = fmaf(w5[2], t3.y, a1);
                    a0 = fmaf(w5[3], t3.z, a0); a1 = fmaf(w5[4], t3.w, a1);
                    acc[b] = a0 + a1;
                }
                // transposing butterfly: lane q of the octet ends with the sum of pixel b0+q
                float r4[4], r2[2];
                const bool h2 = lane8 & 4, h1 = lane8 & 2, h0 = lane8 & 1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float send = h2 ? acc[i] : acc[i + 4];
                    float keep = h2 ? acc[i + 4] : acc[i];
                    r4[i] = keep + __shfl_xor_sync(omask, send, 4);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float send = h1 ? r4[i] : r4[i + 2];
                    float keep = h1 ? r4[i + 2] : r4[i];
                    r2[i] = keep + __shfl_xor_sync(omask, send, 2);
                }
                float send = h0 ? r2[0] : r2[1];
                float keep = h0 ? r2[1] : r2[0];
                float v = keep + __shfl_xor_sync(omask, send, 1);
                const int ox = oxs + b0 + lane8;
                if (ox < p.ow) store_px(drow + (S * ox + px), v);
            }
        }
        __syncthreads();   // tile consumed: its buffer may be refilled by the next prefetch
    }
    cp_async_wait<0>();
}

}  // namespace raisr
