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
// OTW x OTH own pixels per tile, one item of IW pixels per octet, NT threads.
template <>
struct OctetCfg<2> { static constexpr int OTW = 64, OTH = 32, IW = 32, NT = 512; };
template <>
struct OctetCfg<3> { static constexpr int OTW = 64, OTH = 16, IW = 16, NT = 512; };
template <>
struct OctetCfg<4> { static constexpr int OTW = 32, OTH = 16, IW = 16, NT = 256; };

template <int S>
struct OctetGeom {
    using C = OctetCfg<S>;
    static constexpr int TUH = S * (C::OTH - 1) + kFlen;                              // tile rows
    static constexpr int TUW = ((S * (C::OTW - 1) + kFlen + (S - 1)) + 3) / 4 * 4;      // tile pitch, 16-byte rows
    static constexpr int NEWF = S;                 // fresh values per pixel in the 11-run
    static constexpr int NEWP = S < 5 ? S : 5;     // fresh values per pixel in the 5-run
    static constexpr int WF = (S == 3) ? 12 : 16;  // circular register window of the 11-run: 8*S % WF == 0
    static constexpr int WP = 8;                   // circular register window of the 5-run:  8*S % 8 == 0
    static constexpr int SEGS = C::OTW / C::IW;
    static constexpr int ITEMS = C::OTH * SEGS;
    static constexpr int NOCT = C::NT / 8;
    static constexpr int TILE_FLOATS = TUH * TUW;
    static constexpr int HASH_BYTES = C::OTH * C::OTW;   // one byte per own pixel of the tile
    static constexpr int BUF_BYTES = TILE_FLOATS * 4 + HASH_BYTES;
    static_assert(C::IW % 8 == 0 && C::OTW % C::IW == 0 && C::OTW % 16 == 0, "items are whole batches of 8 pixels");
    static_assert((8 * S) % WF == 0 && WF >= kFlen && BUF_BYTES % 16 == 0, "window period / alignment");
    static_assert(ITEMS == NOCT, "one item per octet and tile");
};

template <int S>
inline size_t octet_smem_bytes(int n_buckets)
{
    using G = OctetGeom<S>;
    size_t table = (size_t)(n_buckets > 256 ? n_buckets : 256) * kOctStride * sizeof(float);  // any hash byte stays inside
    return table + 2 * (size_t)G::BUF_BYTES;
}

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Position of the persistent worker in the (frame, tile row, tile column) sequence, advanced by
// `nworkers` tiles at a time without integer divisions.
struct TileCursor {
    int frame, ty, tx;
    __device__ __forceinline__ void init(int tile, int tiles_x, int tiles_y)
    {
        const int per_frame = tiles_x * tiles_y;
        frame = tile / per_frame;
        const int rem = tile - frame * per_frame;
        ty = rem / tiles_x;
        tx = rem - ty * tiles_x;
    }
    __device__ __forceinline__ void advance(int n, int tiles_x, int tiles_y)
    {
        tx += n;
        while (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        while (ty >= tiles_y) { ty -= tiles_y; ++frame; }
    }
};

// Asynchronous fill of one tile buffer: the U tile (rows er0.., columns ec0.. of the extended
// upscaled frame, ec0 a multiple of 4) followed by the tile's hash bytes (OTH rows of OTW bytes).
template <int S>
__device__ __forceinline__ void octet_issue_tile(const FilterParams& p, unsigned char* buf, const TileCursor& tc, int type, int py, int px)
{
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    const int er0 = S * tc.ty * C::OTH + py;
    const int ec0 = (S * tc.tx * C::OTW + px) & ~3;
    const float* ug = p.uext + (size_t)tc.frame * p.uext_frame_stride;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
    constexpr int C4 = G::TUW / 4;
    const int maxc4 = ((int)p.uext_pitch - ec0) / 4 - 1;
    for (int idx = threadIdx.x; idx < G::TUH * C4; idx += C::NT) {
        const int r = idx / C4, c4 = idx - r * C4;
        const float* g = ug + (size_t)min(er0 + r, p.uext_rows - 1) * p.uext_pitch + ec0 + 4 * min(c4, maxc4);
        cp_async16(sbase + 16u * idx, g);
    }
    const uint8_t* hp = p.hash + (size_t)tc.frame * p.hash_frame_stride + (size_t)type * p.hash_plane_stride;
    constexpr int H16 = C::OTW / 16;
    const int maxh = ((int)p.hash_pitch - tc.tx * C::OTW) / 16 - 1;
    for (int idx = threadIdx.x; idx < C::OTH * H16; idx += C::NT) {
        const int r = idx / H16, c = idx - r * H16;
        const uint8_t* g = hp + (size_t)min(tc.ty * C::OTH + r, p.oh - 1) * p.hash_pitch + tc.tx * C::OTW + 16 * min(c, maxh);
        cp_async16(sbase + G::TILE_FLOATS * 4 + 16u * idx, g);
    }
}

template <int S, typename OutT>
__global__ void __launch_bounds__(OctetCfg<S>::NT, 1) filter_octet_kernel(const FilterParams p)
{
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);                 // 512-byte records, 128-B aligned
    unsigned char* buf0 = smem_raw + (size_t)max(p.n_buckets, 256) * kOctStride * sizeof(float);
    unsigned char* buf1 = buf0 + G::BUF_BYTES;
    const int tid = threadIdx.x;
    const int ntypes = S * S;
    const int type = blockIdx.x % ntypes, worker = blockIdx.x / ntypes, nworkers = gridDim.x / ntypes;
    const int py = type / S, px = type % S;
    const int lane8 = tid & 7, octet = tid >> 3;
    if ((__cvta_generic_to_shared(tab) & 127) != 0) __trap();  // records must be 128-byte aligned

    const int ntiles = p.tiles_x * p.tiles_y * p.n_frames;
    // this octet's item: a warp = 4 consecutive own rows of one segment
    const int seg = octet / C::OTH, row = octet - seg * C::OTH;

    TileCursor cur, nxt;
    cur.init(min(worker, max(ntiles - 1, 0)), p.tiles_x, p.tiles_y);
    nxt = cur;
    if (worker < ntiles) octet_issue_tile<S>(p, buf0, cur, type, py, px);   // in flight while the table is copied
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
    const int item_off = (S * row) * G::TUW + S * (seg * C::IW) + px;   // + px: the tile starts at a multiple of 4

    int it = 0;
    for (int tile = worker; tile < ntiles; tile += nworkers, ++it) {
        unsigned char* buf = (it & 1) ? buf1 : buf0;
        nxt.advance(nworkers, p.tiles_x, p.tiles_y);
        if (tile + nworkers < ntiles && !(p.dbg_flags & 4)) octet_issue_tile<S>(p, (it & 1) ? buf0 : buf1, nxt, type, py, px);
        cp_async_commit();
        cp_async_wait<1>();   // everything but the newest group (the prefetch) has landed
        __syncthreads();

        const int oy = cur.ty * C::OTH + row;
        const int oxs = cur.tx * C::OTW + seg * C::IW;      // first own column of the item
        if (oy < p.oh && oxs < p.ow) {                      // octet-uniform
            OutT* drow = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(p.dst) + (size_t)cur.frame * p.dst_frame_stride +
                                                 (size_t)(S * oy + py) * p.dst_pitch);
            const float* base = reinterpret_cast<const float*>(buf) + item_off;   // patch origin of the item's first pixel
            const float* pf = base + off_full;
            const uint2* hrow = reinterpret_cast<const uint2*>(buf + G::TILE_FLOATS * 4 + row * C::OTW + seg * C::IW);
            // circular register windows: element j of pixel i lives in slot (S*i + j) % W
            float w11[G::WF], w5[G::WP];
#pragma unroll
            for (int j = 0; j < G::WF; ++j) w11[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < G::WP; ++j) w5[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < kFlen - G::NEWF; ++j) w11[j] = pf[j];
            if (lane8 < 6) {
                const float* pp = base + (8 + lane8 / 2) * G::TUW + 5 * (lane8 % 2);
#pragma unroll
                for (int t = 0; t < 5 - G::NEWP; ++t) w5[t] = pp[t];
            }
#pragma unroll
            for (int t = 0; t < G::NEWF; ++t) w11[kFlen - G::NEWF + t] = pf[kFlen - G::NEWF + t];
#pragma unroll
            for (int t = 0; t < G::NEWP; ++t) w5[5 - G::NEWP + t] = base[off_part[t]];
            uint2 hb = hrow[0];
            // taps of the first pixel; afterwards the taps of pixel i+1 are requested before the
            // FMAs of pixel i (two register sets, even / odd pixel)
            float4 ta[4], tb[4];
            {
                const float4* tp = tab_lane + (hb.x & 0xffu) * (kOctStride / 4);
                ta[0] = tp[0]; ta[1] = tp[8]; ta[2] = tp[16]; ta[3] = tp[24];
            }
#pragma unroll 1
            for (int b0 = 0; b0 < C::IW; b0 += 8) {
                if (oxs + b0 >= p.ow) break;                  // octet-uniform
                const uint2 hnext = hrow[min(b0 / 8 + 1, C::IW / 8 - 1)];   // hash bytes of the next batch
                float acc[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    // request the taps of the next pixel
                    const unsigned nbucket = (b < 7) ? (((b + 1 < 4 ? hb.x : hb.y) >> (8 * ((b + 1) & 3))) & 0xffu) : (hnext.x & 0xffu);
                    const float4* tp = tab_lane + nbucket * (kOctStride / 4);
                    float4 (&tc)[4] = (b & 1) ? tb : ta;
                    float4 (&tn)[4] = (b & 1) ? ta : tb;
                    tn[0] = tp[0]; tn[1] = tp[8]; tn[2] = tp[16]; tn[3] = tp[24];
                    // 16 FMAs of this pixel against its windows
                    constexpr int MP = G::WP - 1;
                    const int o = S * b;
                    float a0 = w11[(o + 0) % G::WF] * tc[0].x, a1 = w11[(o + 1) % G::WF] * tc[0].y;
                    a0 = fmaf(w11[(o + 2) % G::WF], tc[0].z, a0); a1 = fmaf(w11[(o + 3) % G::WF], tc[0].w, a1);
                    a0 = fmaf(w11[(o + 4) % G::WF], tc[1].x, a0); a1 = fmaf(w11[(o + 5) % G::WF], tc[1].y, a1);
                    a0 = fmaf(w11[(o + 6) % G::WF], tc[1].z, a0); a1 = fmaf(w11[(o + 7) % G::WF], tc[1].w, a1);
                    a0 = fmaf(w11[(o + 8) % G::WF], tc[2].x, a0); a1 = fmaf(w11[(o + 9) % G::WF], tc[2].y, a1);
                    a0 = fmaf(w11[(o + 10) % G::WF], tc[2].z, a0); a1 = fmaf(w5[(o + 0) & MP], tc[2].w, a1);
                    a0 = fmaf(w5[(o + 1) & MP], tc[3].x, a0); a1 = fmaf(w5[(o + 2) & MP], tc[3].y, a1);
                    a0 = fmaf(w5[(o + 3) & MP], tc[3].z, a0); a1 = fmaf(w5[(o + 4) & MP], tc[3].w, a1);
                    // fresh patch values of the next pixel overwrite the slots this pixel has just consumed
                    const int npix = b0 + b + 1;
                    if (npix < C::IW && !(p.dbg_flags & 16)) {
                        const int on = S * (b + 1);
#pragma unroll
                        for (int t = 0; t < G::NEWF; ++t) w11[(on + kFlen - G::NEWF + t) % G::WF] = pf[S * npix + kFlen - G::NEWF + t];
#pragma unroll
                        for (int t = 0; t < G::NEWP; ++t) w5[(on + 5 - G::NEWP + t) & MP] = base[S * npix + off_part[t]];
                    }
                    acc[b] = a0 + a1;
                }
                hb = hnext;
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
                if (ox < p.ow && !(p.dbg_flags & 2)) store_px(drow + (S * ox + px), v);
            }
        }
        cur = nxt;
        __syncthreads();   // tile consumed: its buffer may be refilled by the next prefetch
    }
    cp_async_wait<0>();
}

}  // namespace raisr
