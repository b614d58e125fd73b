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
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "packed_f32.cuh"
#include "raisr_filter.cuh"
#include "raisr_prep.cuh"     // eigen_bucket() for the EIG variant

namespace raisr {

constexpr int kOctStride = 128;  // floats per filter record (512 B)
// Rows 8..10 of the filter are split into two 5-runs (columns kRunCol0.. and kRunCol0+5..) and one
// leftover tap (column kSingleCol); this split leaves the fewest shared-memory bank conflicts on the
// per-pixel loads of the 5-runs for the tile pitches in use (searched offline).
constexpr int kRunCol0 = 1, kSingleCol = 0;

// Host-side packing of one 11x11 filter (row-major taps f[i*11+j]) into the lane-major record:
// record float (p + 8n)*4 + c  <->  slot s = 4n + c of lane p.
//   slots 0..10  : filter row p, columns 0..10            (p = 0..7)
//   slots 11..15 : lanes 0..5: row 8 + p/2, columns 1 + 5*(p%2) .. +4
//                  lanes 6,7 : the three leftover taps (8,0) (9,0) (10,0) in the slots that are
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
        for (int t = 0; t < 5; ++t) put(p, 11 + t, f[(8 + p / 2) * kFlen + kRunCol0 + 5 * (p % 2) + t]);
    for (int idx = 0; idx < 3; ++idx) {
        int lane, slot;
        octet_single_slot(S, idx, &lane, &slot);
        put(lane, slot, f[(8 + idx) * kFlen + kSingleCol]);
    }
}

// fp16 tap records ("taps_fp16", the reference multiplies by (half)pf[...], raisr.cl:328): 128 halfs = 256 B,
// two 16-byte chunks per lane; lane p reads chunks p and p+8, slot s of lane p is half (p + 8*(s/8))*8 + s%8.
// `to_half` converts a float to the raw bits of its round-to-nearest fp16 value.
constexpr int kOctStrideH = 128;   // halfs per record
template <typename ToHalf>
inline void octet_pack_filter_h16(const float* f, uint16_t* rec, int S, ToHalf to_half)
{
    float tmp[kOctStride];
    octet_pack_filter_s(f, tmp, S);
    for (int p = 0; p < 8; ++p)
        for (int slot = 0; slot < 16; ++slot)
            rec[(p + 8 * (slot / 8)) * 8 + slot % 8] = to_half(tmp[(p + 8 * (slot / 4)) * 4 + slot % 4]);
}

// 24-bit tap records ("taps_b24"): every tap keeps its sign, its 8 exponent bits and the top 15 mantissa bits
// (3 bytes instead of 4), which cuts the shared-memory tap stream -- the resource that bounds this kernel -- by
// a quarter.  A lane's 16 taps form a circular stream of 48 bytes = three 16-byte chunks; tap k owns stream
// bytes 3k+1 .. 3k+3 (mod 48, most significant last) and is used as the little-endian 32-bit word that starts at
// byte 3k: its low byte is the top byte of tap k-1, a known value, so the packer picks the 24 own bits that make
// the WHOLE word the float nearest to the original tap (|error| <= 2^-16 relative, half of plain truncation's
// worst case) and the kernel needs one PRMT per tap (none for taps 0, 4, 8, 12, which are word-aligned).
// Record = 24 chunks = 384 B; lane p reads chunks p, p+8, p+16: conflict-free like the fp32 record.
constexpr int kOctBytesB24 = 384;

inline float b24_word_to_float(uint32_t w)
{
    float f;
    memcpy(&f, &w, 4);
    return f;
}

// Encodes 16 slot values into the 48-byte circular stream; `eff` receives the values the kernel will decode.
inline void b24_encode_lane(const float v[16], uint8_t stream[48], float eff[16])
{
    uint32_t bits[16], own[16];   // own = the 24 stored bits (sign, exponent, 15 mantissa bits)
    for (int k = 0; k < 16; ++k) { memcpy(&bits[k], &v[k], 4); own[k] = bits[k] >> 8; }
    for (int pass = 0; pass < 4; ++pass) {
        bool changed = false;
        for (int k = 0; k < 16; ++k) {
            const uint32_t g = own[(k + 15) & 15] >> 16;                 // low byte of the word: top byte of tap k-1
            const uint32_t sign = bits[k] & 0x80000000u, mag = bits[k] & 0x7fffffffu;
            uint32_t best = own[k];
            double best_err = 1e300;
            for (int d = -1; d <= 1; ++d) {
                const long long hm = (long long)(mag >> 8) + d;
                if (hm < 0 || hm > 0x7f7fff) continue;                   // stay finite
                const uint32_t w = sign | ((uint32_t)hm << 8) | g;
                const double err = fabs((double)b24_word_to_float(w) - (double)v[k]);
                if (err < best_err) { best_err = err; best = w >> 8; }
            }
            if (best != own[k]) { own[k] = best; changed = true; }
        }
        if (!changed) break;
    }
    for (int k = 0; k < 16; ++k) {
        stream[(3 * k + 1) % 48] = (uint8_t)(own[k] & 0xff);
        stream[(3 * k + 2) % 48] = (uint8_t)((own[k] >> 8) & 0xff);
        stream[(3 * k + 3) % 48] = (uint8_t)(own[k] >> 16);
    }
    for (int k = 0; k < 16; ++k) {
        const uint32_t w = (uint32_t)stream[3 * k] | ((uint32_t)stream[3 * k + 1] << 8) | ((uint32_t)stream[3 * k + 2] << 16) |
                           ((uint32_t)stream[(3 * k + 3) % 48] << 24);
        eff[k] = b24_word_to_float(w);
    }
}

// Packs one 11x11 filter into a 384-byte b24 record; `eff` (121 floats, may be null) receives the tap values the
// kernel will actually multiply by, in the reference's row-major order.
inline void octet_pack_filter_b24(const float* f, uint8_t* rec, int S, float* eff)
{
    float slots[kOctStride], idx[kOctStride], fi[kTaps];
    octet_pack_filter_s(f, slots, S);
    for (int t = 0; t < kTaps; ++t) fi[t] = (float)(t + 1);
    octet_pack_filter_s(fi, idx, S);                                     // which tap sits in which slot (0 = unused)
    for (int p = 0; p < 8; ++p) {
        float v[16], e[16];
        uint8_t stream[48];
        for (int sl = 0; sl < 16; ++sl) v[sl] = slots[(p + 8 * (sl / 4)) * 4 + sl % 4];
        b24_encode_lane(v, stream, e);
        for (int n = 0; n < 48; ++n) rec[(p + 8 * (n / 16)) * 16 + n % 16] = stream[n];
        if (eff)
            for (int sl = 0; sl < 16; ++sl) {
                const int t = (int)idx[(p + 8 * (sl / 4)) * 4 + sl % 4];
                if (t > 0) eff[t - 1] = e[sl];
            }
    }
}

template <int S>
struct OctetCfg;
// OTW x OTH own pixels per tile, one item of IW pixels per octet, NT threads.
template <>
struct OctetCfg<2> { static constexpr int OTW = 64, OTH = 40, IW = 32, NT = 640; };
template <>
struct OctetCfg<3> { static constexpr int OTW = 64, OTH = 16, IW = 16, NT = 512; };
template <>
struct OctetCfg<4> { static constexpr int OTW = 32, OTH = 16, IW = 16, NT = 256; };

// The U tile is kept COLUMN-MAJOR in shared memory: element (row r, column c) of the tile at
// c * PT + r.  The eight lanes of an octet read eight consecutive rows of one column and the four
// octets of a warp sit on consecutive own rows, so every per-pixel LDS.32 of a warp touches 14
// consecutive words: conflict-free, and the tile needs no per-thread copy instructions because
// uext is stored column-major too (raisr_prep.cuh): one TMA box (PT rows x NCOLS columns) per tile.
template <int S>
struct OctetGeom {
    using C = OctetCfg<S>;
    static constexpr int TUH = S * (C::OTH - 1) + kFlen;                 // tile rows needed
    static constexpr int NCOLS = S * (C::OTW - 1) + kFlen;               // tile columns needed
    static constexpr int PT = (TUH + 3 + 3) / 4 * 4;                     // floats per tile column (TMA box inner size; the box starts
                                                                         // at a row multiple of 4: TMA wants 16-byte aligned inner starts)
    static constexpr int NEWF = S;                 // fresh values per pixel in the 11-run
    static constexpr int NEWP = S < 5 ? S : 5;     // fresh values per pixel in the 5-run
    static constexpr int WF = (S == 3) ? 12 : 16;  // circular register window of the 11-run: 8*S % WF == 0
    static constexpr int WP = 8;                   // circular register window of the 5-run:  8*S % 8 == 0
    static constexpr int SEGS = C::OTW / C::IW;
    static constexpr int ITEMS = C::OTH * SEGS;
    static constexpr int NOCT = C::NT / 8;
    static constexpr int TILE_FLOATS = NCOLS * PT;
    static constexpr int HASH_BYTES = C::OTH * C::OTW;   // one byte per own pixel of the tile
    static constexpr int TILE_BYTES = TILE_FLOATS * 4;
    static constexpr int HASH_OFF = (TILE_BYTES + 127) / 128 * 128;      // TMA destinations are 128-byte aligned
    static constexpr int BUF_BYTES = (HASH_OFF + HASH_BYTES + 127) / 128 * 128;
    static_assert(C::IW % 8 == 0 && C::OTW % C::IW == 0 && C::OTW % 16 == 0, "items are whole batches of 8 pixels");
    static_assert((8 * S) % WF == 0 && WF >= kFlen && TILE_BYTES % 16 == 0 && NCOLS <= 256 && PT <= 256, "window period / alignment");
    static_assert(ITEMS == NOCT && (S * C::OTH) % 4 == 0, "one item per octet and tile; tiles start on row quads");
};

__host__ __device__ constexpr int octet_record_bytes(int tf) { return tf == kTapsF16 ? kOctStrideH * 2 : tf == kTapsB24 ? kOctBytesB24 : kOctStride * 4; }

// The patch look-ahead of an item's last pixel reads up to S tile columns past the last tile buffer.
constexpr int kOctetTailPad = 2048;

template <int S, int NBUF = 2>
inline size_t octet_smem_bytes(int n_buckets, int tf = kTapsF32)
{
    using G = OctetGeom<S>;
    const size_t rec = octet_record_bytes(tf);
    return (size_t)n_buckets * rec + NBUF * (size_t)G::BUF_BYTES + 32 + kOctetTailPad;   // + mbarriers / counters + look-ahead slack
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

// Predicated in-place 128-bit shared load: t is overwritten only when `pred` is set (no register
// renaming / move chains around the conditional reload of the tap registers).
__device__ __forceinline__ void lds128_if(float4& t, unsigned saddr, bool pred)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %5, 0;\n\t"
        "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(t.x), "+f"(t.y), "+f"(t.z), "+f"(t.w)
        : "r"(saddr), "r"((unsigned)pred));
}

__device__ __forceinline__ void lds128_if(uint4& t, unsigned saddr, bool pred)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %5, 0;\n\t"
        "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+r"(t.x), "+r"(t.y), "+r"(t.z), "+r"(t.w)
        : "r"(saddr), "r"((unsigned)pred));
}
// two packed halfs -> two floats (exact)
__device__ __forceinline__ float2 h2f2(unsigned v)
{
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Producer side of the pipelined kernel: one thread fills a tile buffer with two TMA boxes -- the U tile
// (as below) and the tile's hash bytes (OTW x OTH box of the planar hash image; rows past the image are
// zero-filled) -- both signalled on the buffer's "full" mbarrier.
template <int S, bool WITH_HASH = true>
__device__ __forceinline__ void octet_issue_tile_tma(const CUtensorMap* tm, const CUtensorMap* hm, unsigned char* buf, unsigned bar,
                                                     const TileCursor& tc, int type, int py, int px)
{
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, G::TILE_BYTES + (WITH_HASH ? G::HASH_BYTES : 0));
    const int r0 = (S * tc.ty * C::OTH + py) & ~3, c0 = S * tc.tx * C::OTW + px;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(sbase), "l"(tm), "r"(r0), "r"(c0), "r"(tc.frame), "r"(bar) : "memory");
    if (WITH_HASH)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(sbase + G::HASH_OFF), "l"(hm), "r"(tc.tx * C::OTW), "r"(tc.ty * C::OTH), "r"(type), "r"(tc.frame), "r"(bar) : "memory");
}

// Asynchronous fill of one tile buffer.  The U tile (rows S*oy0+py .., columns S*ox0+px .. of the
// extended upscaled frame, PT x NCOLS, column-major in HBM and in shared memory) is one TMA box
// issued by a single thread and signalled on an mbarrier; the tile's hash bytes (OTH rows of OTW
// bytes) follow with 16-byte cp.async.
template <int S>
__device__ __forceinline__ void octet_issue_tile(const FilterParams& p, const CUtensorMap* tm, unsigned char* buf, unsigned bar,
                                                 const TileCursor& tc, int type, int py, int px)
{
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this buffer are done
        mbar_expect_tx(bar, G::TILE_BYTES);
        const int r0 = (S * tc.ty * C::OTH + py) & ~3, c0 = S * tc.tx * C::OTW + px;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sbase), "l"(tm), "r"(r0), "r"(c0), "r"(tc.frame), "r"(bar) : "memory");
    }
    const uint8_t* hp = p.hash + (size_t)tc.frame * p.hash_frame_stride + (size_t)type * p.hash_plane_stride;
    constexpr int H16 = C::OTW / 16;
    const int maxh = ((int)p.hash_pitch - tc.tx * C::OTW) / 16 - 1;
    for (int idx = threadIdx.x; idx < C::OTH * H16; idx += C::NT) {
        const int r = idx / H16, c = idx - r * H16;
        const uint8_t* g = hp + (size_t)min(tc.ty * C::OTH + r, p.oh - 1) * p.hash_pitch + tc.tx * C::OTW + 16 * min(c, maxh);
        cp_async16(sbase + G::HASH_OFF + 16u * idx, g);
    }
}

// NBUF = 2: the next tile is fetched while the current one is filtered.  NBUF = 1 (used when the
// prep kernel of the next chunk shares the SM, see raisr_api.cu): one buffer, the fetch of the next
// tile starts when the current one is done and the co-resident kernel fills the gap.
// H16: the resident table holds fp16 taps (256-byte records, half the shared-memory tap stream); they are
// widened to fp32 in registers and the arithmetic is the same fp32 FMA chain.
// PIPE (NBUF = 2 only): no CTA-wide barrier per tile.  Both buffers are filled by TMA (U tile + hash bytes) and
// signalled on a "full" mbarrier each; a warp that has finished a tile bumps a per-buffer counter, and the
// warp that arrives LAST -- at that point nobody reads the buffer any more -- issues the TMA of the tile after
// next into it.  Warps wait only for the tile they need, so they drift apart and the shared-memory pipe no
// longer drains at every tile boundary (filter 11.04 -> 10.4 ms per step).
// TF = kTapsB24: 384-byte records of 24-bit taps (see octet_pack_filter_b24): three predicated LDS.128 per lane and
// pixel instead of four, twelve PRMT on the otherwise idle integer pipe, the same fp32 FMA chain.
// EIG (PIPE, b24 only): "eigen in the filter kernel".  prep2_kernel stops after the structure tensor and stores
// ma, mb, md; here lane q of an octet solves the eigen problem of pixel b0 + q of every batch of eight (the scalar
// sequence eigen_bucket(), bit-identical to the prep kernels), and three shuffles pack the eight buckets into the two
// words the pixel loop reads its hash bytes from.  The eigen-solve is 52 % of the prep kernel's instructions, and the
// packed-FFMA2 filter kernel is bound by shared memory with a third of its issue slots idle: the work moves from an
// issue-bound kernel into the issue shadow of a memory-bound one.
template <int S, typename OutT, int NBUF = 2, int TF = kTapsF32, bool PIPE = false, bool EIG = false>
__global__ void __launch_bounds__(OctetCfg<S>::NT, 1)
    filter_octet_kernel(const FilterParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap hmap)
{
    static_assert(!PIPE || NBUF == 2, "the pipelined kernel is double-buffered");
    static_assert(!EIG || (PIPE && TF == kTapsB24), "the eigen-in-filter variant is built for the pipelined b24 kernel");
    using C = OctetCfg<S>;
    using G = OctetGeom<S>;
    constexpr bool H16 = TF == kTapsF16, B24 = TF == kTapsB24;
    constexpr int REC = octet_record_bytes(TF);                      // bytes per filter record
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);                 // 128-B aligned records
    unsigned char* buf0 = smem_raw + (size_t)p.n_buckets * REC;
    unsigned char* buf1 = buf0 + (NBUF - 1) * G::BUF_BYTES;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(buf1 + G::BUF_BYTES), bar1 = bar0 + 8;
    int* done_cnt = reinterpret_cast<int*>(buf1 + G::BUF_BYTES + 16);   // warps done with buffer 0 / 1
    const int tid = threadIdx.x;
    const int ntypes = S * S;
    const int type = blockIdx.x % ntypes, worker = blockIdx.x / ntypes, nworkers = gridDim.x / ntypes;
    const int py = type / S, px = type % S;
    const int lane8 = tid & 7, octet = tid >> 3;
    if ((__cvta_generic_to_shared(tab) & 127) != 0) __trap();  // records and TMA destinations must be 128-byte aligned
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        done_cnt[0] = 0;
        done_cnt[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int ntiles = p.tiles_x * p.tiles_y * p.n_frames;
    // this octet's item: a warp = 4 consecutive own rows of one segment
    const int seg = octet / C::OTH, row = octet - seg * C::OTH;

    TileCursor cur, nxt;
    cur.init(min(worker, max(ntiles - 1, 0)), p.tiles_x, p.tiles_y);
    nxt = cur;
    if (!PIPE) {
        if (worker < ntiles) octet_issue_tile<S>(p, &tmap, buf0, bar0, cur, type, py, px);   // in flight while the table is copied
        cp_async_commit();
    }
    if (PIPE && tid == 0) {
        // the first two tiles need no release and fly while the table is copied
        TileCursor pc = cur;
        int pit = 0;
        for (int tile = worker; tile < ntiles && pit < 2; tile += nworkers, ++pit) {
            octet_issue_tile_tma<S, !EIG>(&tmap, &hmap, pit ? buf1 : buf0, pit ? bar1 : bar0, pc, type, py, px);
            pc.advance(nworkers, p.tiles_x, p.tiles_y);
        }
    }
    {
        const float4* g = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(p.table) + (size_t)type * p.n_buckets * REC);
        float4* s = reinterpret_cast<float4*>(tab);
        for (int i = tid; i < p.n_buckets * (REC / 16); i += C::NT) s[i] = __ldg(g + i);
    }
    if (PIPE) __syncthreads();            // table slice resident; the only CTA-wide barrier of the pipelined kernel

    // Lane geometry: offsets (floats) from the patch origin of the current pixel in the column-major tile.
    const int off_full = lane8;           // filter row = lane8, column 0
    int off_part[G::NEWP];                // addresses of the freshly loaded 5-run slots
    if (lane8 < 6) {
#pragma unroll
        for (int t = 0; t < G::NEWP; ++t)
            off_part[t] = (8 + lane8 / 2) + (kRunCol0 + 5 * (lane8 % 2) + (5 - G::NEWP) + t) * G::PT;
    } else {
#pragma unroll
        for (int t = 0; t < G::NEWP; ++t) {
            int idx = (lane8 - 6) * G::NEWP + t;   // leftover taps (8,c) (9,c) (10,c), c = kSingleCol
            off_part[t] = (idx < 3 ? (8 + idx) : 10) + kSingleCol * G::PT;
        }
    }
    const float4* tab_lane = reinterpret_cast<const float4*>(tab) + lane8;
    const unsigned tab_lane_s = (unsigned)__cvta_generic_to_shared(tab_lane);
    const unsigned omask = 0xffu << (tid & 24);  // the eight lanes of this octet

    int it = 0;
    for (int tile = worker; tile < ntiles; tile += nworkers, ++it) {
        unsigned char* buf = (NBUF == 2 && (it & 1)) ? buf1 : buf0;
        nxt.advance(nworkers, p.tiles_x, p.tiles_y);
        if (PIPE) {
            mbar_wait((it & 1) ? bar1 : bar0, (it >> 1) & 1);   // U tile and hash bytes of this tile have landed
        } else if (NBUF == 2) {
            if (tile + nworkers < ntiles)
                octet_issue_tile<S>(p, &tmap, (it & 1) ? buf0 : buf1, (it & 1) ? bar0 : bar1, nxt, type, py, px);
            cp_async_commit();
            cp_async_wait<1>();   // hash bytes: everything but the newest group (the prefetch) has landed
            mbar_wait((it & 1) ? bar1 : bar0, (it >> 1) & 1);   // U tile: TMA bytes have landed
        } else {
            cp_async_wait<0>();
            mbar_wait(bar0, it & 1);
        }
        if (!PIPE) __syncthreads();

        const int oy = cur.ty * C::OTH + row;
        const int oxs = cur.tx * C::OTW + seg * C::IW;      // first own column of the item
        if (oy < p.oh && oxs < p.ow) {                      // octet-uniform
            OutT* drow = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(p.dst) + (size_t)cur.frame * p.dst_frame_stride +
                                                 (size_t)(S * oy + py) * p.dst_pitch);
            const float* base = reinterpret_cast<const float*>(buf) + S * (seg * C::IW) * G::PT + S * row + (py & 3);   // S*OTH % 4 == 0
            const float* pf = base + off_full;
            const uint2* hrow = reinterpret_cast<const uint2*>(buf + G::HASH_OFF + row * C::OTW + seg * C::IW);
            // EIG: structure tensor of "my" pixel (lane8) of the next two batches of eight, fetched ahead
            const float* trow = nullptr;
            float e0a = 0, e0b = 0, e0d = 0, e1a = 0, e1b = 0, e1d = 0;
            auto tens_load = [&](int ox, float& a, float& b, float& d) {
                const float* t = trow + min(ox, p.ow - 1);
                a = __ldg(t); b = __ldg(t + p.tens_plane_stride); d = __ldg(t + 2 * p.tens_plane_stride);
            };
            auto eig_pack = [&](float a, float b, float d) {      // my pixel's bucket -> the eight hash bytes of the batch
                const float sq[2] = {p.sq[0], p.sq[1]}, cq[2] = {p.cq[0], p.cq[1]};
                const unsigned mine = (unsigned)eigen_bucket<2>(a, b, d, sq, cq, p.n_angle, p.n_strength, p.n_coherence, p.as_written != 0);
                unsigned v = mine << (8 * (lane8 & 3));
                v |= __shfl_xor_sync(omask, v, 1);
                v |= __shfl_xor_sync(omask, v, 2);
                const unsigned other = __shfl_xor_sync(omask, v, 4);
                return (lane8 & 4) ? make_uint2(other, v) : make_uint2(v, other);
            };
            if (EIG) {
                trow = p.tens + (size_t)cur.frame * p.hash_frame_stride + (size_t)type * p.hash_plane_stride + (size_t)oy * p.hash_pitch;
                tens_load(oxs + lane8, e0a, e0b, e0d);
                tens_load(oxs + 8 + lane8, e1a, e1b, e1d);
            }
            // circular register windows: element j of pixel i lives in slot (S*i + j) % W
            float w11[G::WF], w5[G::WP];
#pragma unroll
            for (int j = 0; j < G::WF; ++j) w11[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < G::WP; ++j) w5[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < kFlen; ++j) w11[j] = pf[j * G::PT];
            if (lane8 < 6) {
                const float* pp = base + (8 + lane8 / 2) + (kRunCol0 + 5 * (lane8 % 2)) * G::PT;
#pragma unroll
                for (int t = 0; t < 5 - G::NEWP; ++t) w5[t] = pp[t * G::PT];
            }
#pragma unroll
            for (int t = 0; t < G::NEWP; ++t) w5[5 - G::NEWP + t] = base[off_part[t]];
            // hash bytes are always < n_buckets: the prep kernel writes valid buckets only and the host zero-fills the
            // scratch when it allocates it (padding bytes never hold anything else), so no clamp is needed here
            uint2 hb = EIG ? eig_pack(e0a, e0b, e0d) : hrow[0];
            unsigned bucket = hb.x & 0xffu;
            const float4* tp = tab_lane + bucket * (REC / 16);
            float4 t0 = {}, t1 = {}, t2 = {}, t3 = {};
            uint4 q0 = {}, q1 = {}, q2 = {};
            if (B24) {
                q0 = *reinterpret_cast<const uint4*>(tp); q1 = *reinterpret_cast<const uint4*>(tp + 8); q2 = *reinterpret_cast<const uint4*>(tp + 16);
            } else if (H16) {
                q0 = *reinterpret_cast<const uint4*>(tp); q1 = *reinterpret_cast<const uint4*>(tp + 8);
            } else {
                t0 = tp[0]; t1 = tp[8]; t2 = tp[16]; t3 = tp[24];
            }
#pragma unroll 1
            for (int b0 = 0; b0 < C::IW; b0 += 8) {
                if (oxs + b0 >= p.ow) break;                  // octet-uniform
                uint2 hnext;                                  // hash bytes of the next batch
                if (EIG) {
                    hnext = make_uint2(0u, 0u);
                    if (b0 + 8 < C::IW) {                         // uniform
                        hnext = eig_pack(e1a, e1b, e1d);
                        if (b0 + 16 < C::IW) tens_load(oxs + b0 + 16 + lane8, e1a, e1b, e1d);
                    }
                } else {
                    hnext = hrow[min(b0 / 8 + 1, C::IW / 8 - 1)];
                }
                float acc[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    // taps of the next pixel: each 16-byte chunk is reloaded as soon as the FMAs of this
                    // pixel have read it, and only if the next pixel hashes to another bucket
                    // byte (b+1)&3 of the hash word, zero-extended, in one PRMT
                    const unsigned nbucket = (b < 7) ? __byte_perm(b + 1 < 4 ? hb.x : hb.y, 0u, 0x4440u | ((b + 1) & 3)) : (hnext.x & 0xffu);
                    const bool reload = nbucket != bucket;
                    bucket = nbucket;
                    const unsigned tpa = tab_lane_s + nbucket * REC;
                    constexpr int MP = G::WP - 1;
                    const int o = S * b;
                    float a0, a1;
                    if (B24) {
                        // tap k = the 32-bit word at stream byte 3k (see octet_pack_filter_b24)
                        auto W = [](unsigned lo, unsigned hi, unsigned sel) { return __uint_as_float(__byte_perm(lo, hi, sel)); };
                        const float c15 = W(q2.w, q0.x, 0x4321);
                        const float c0 = __uint_as_float(q0.x), c1 = W(q0.x, q0.y, 0x6543), c2 = W(q0.y, q0.z, 0x5432), c3 = W(q0.z, q0.w, 0x4321);
                        const float c4 = __uint_as_float(q0.w), c5 = W(q0.w, q1.x, 0x6543);
                        lds128_if(q0, tpa, reload);
                        const float c6 = W(q1.x, q1.y, 0x5432), c7 = W(q1.y, q1.z, 0x4321);
                        const float c8 = __uint_as_float(q1.z), c9 = W(q1.z, q1.w, 0x6543), c10 = W(q1.w, q2.x, 0x5432);
                        lds128_if(q1, tpa + 128, reload);
                        const float c11 = W(q2.x, q2.y, 0x4321), c12 = __uint_as_float(q2.y), c13 = W(q2.y, q2.z, 0x6543), c14 = W(q2.z, q2.w, 0x5432);
                        lds128_if(q2, tpa + 256, reload);
                        // Two chains, summed at the end: chain 0 takes slots 0,2,4,6,8,11,13 and then 10, chain 1 slots
                        // 1,3,5,7,9,12,14 and then 15.  For even S the window offset o = S*b is even, so (slot 2m, slot 2m+1)
                        // of both windows are fixed register pairs and seven of the nine steps are packed FFMA2 (two IEEE
                        // fp32 FMAs per issue slot; the kernel is bound by issue slots, the FMA pipe is a third busy).
                        if (S % 2 == 0) {
                            p2 acc2 = mul2(pk(w11[(o + 0) % G::WF], w11[(o + 1) % G::WF]), pk(c0, c1));
                            acc2 = fma2(pk(w11[(o + 2) % G::WF], w11[(o + 3) % G::WF]), pk(c2, c3), acc2);
                            acc2 = fma2(pk(w11[(o + 4) % G::WF], w11[(o + 5) % G::WF]), pk(c4, c5), acc2);
                            acc2 = fma2(pk(w11[(o + 6) % G::WF], w11[(o + 7) % G::WF]), pk(c6, c7), acc2);
                            acc2 = fma2(pk(w11[(o + 8) % G::WF], w11[(o + 9) % G::WF]), pk(c8, c9), acc2);
                            acc2 = fma2(pk(w5[(o + 0) & MP], w5[(o + 1) & MP]), pk(c11, c12), acc2);
                            acc2 = fma2(pk(w5[(o + 2) & MP], w5[(o + 3) & MP]), pk(c13, c14), acc2);
                            upk(acc2, a0, a1);
                        } else {
                            a0 = w11[(o + 0) % G::WF] * c0; a1 = w11[(o + 1) % G::WF] * c1;
                            a0 = fmaf(w11[(o + 2) % G::WF], c2, a0); a1 = fmaf(w11[(o + 3) % G::WF], c3, a1);
                            a0 = fmaf(w11[(o + 4) % G::WF], c4, a0); a1 = fmaf(w11[(o + 5) % G::WF], c5, a1);
                            a0 = fmaf(w11[(o + 6) % G::WF], c6, a0); a1 = fmaf(w11[(o + 7) % G::WF], c7, a1);
                            a0 = fmaf(w11[(o + 8) % G::WF], c8, a0); a1 = fmaf(w11[(o + 9) % G::WF], c9, a1);
                            a0 = fmaf(w5[(o + 0) & MP], c11, a0); a1 = fmaf(w5[(o + 1) & MP], c12, a1);
                            a0 = fmaf(w5[(o + 2) & MP], c13, a0); a1 = fmaf(w5[(o + 3) & MP], c14, a1);
                        }
                        a0 = fmaf(w11[(o + 10) % G::WF], c10, a0); a1 = fmaf(w5[(o + 4) & MP], c15, a1);
                    } else if (H16) {
                        float2 c0 = h2f2(q0.x), c1 = h2f2(q0.y), c2 = h2f2(q0.z), c3 = h2f2(q0.w);
                        lds128_if(q0, tpa, reload);
                        a0 = w11[(o + 0) % G::WF] * c0.x; a1 = w11[(o + 1) % G::WF] * c0.y;
                        a0 = fmaf(w11[(o + 2) % G::WF], c1.x, a0); a1 = fmaf(w11[(o + 3) % G::WF], c1.y, a1);
                        a0 = fmaf(w11[(o + 4) % G::WF], c2.x, a0); a1 = fmaf(w11[(o + 5) % G::WF], c2.y, a1);
                        a0 = fmaf(w11[(o + 6) % G::WF], c3.x, a0); a1 = fmaf(w11[(o + 7) % G::WF], c3.y, a1);
                        c0 = h2f2(q1.x); c1 = h2f2(q1.y); c2 = h2f2(q1.z); c3 = h2f2(q1.w);
                        lds128_if(q1, tpa + 128, reload);
                        a0 = fmaf(w11[(o + 8) % G::WF], c0.x, a0); a1 = fmaf(w11[(o + 9) % G::WF], c0.y, a1);
                        a0 = fmaf(w11[(o + 10) % G::WF], c1.x, a0); a1 = fmaf(w5[(o + 0) & MP], c1.y, a1);
                        a0 = fmaf(w5[(o + 1) & MP], c2.x, a0); a1 = fmaf(w5[(o + 2) & MP], c2.y, a1);
                        a0 = fmaf(w5[(o + 3) & MP], c3.x, a0); a1 = fmaf(w5[(o + 4) & MP], c3.y, a1);
                    } else {
                    a0 = w11[(o + 0) % G::WF] * t0.x; a1 = w11[(o + 1) % G::WF] * t0.y;
                    a0 = fmaf(w11[(o + 2) % G::WF], t0.z, a0); a1 = fmaf(w11[(o + 3) % G::WF], t0.w, a1);
                    lds128_if(t0, tpa, reload);
                    a0 = fmaf(w11[(o + 4) % G::WF], t1.x, a0); a1 = fmaf(w11[(o + 5) % G::WF], t1.y, a1);
                    a0 = fmaf(w11[(o + 6) % G::WF], t1.z, a0); a1 = fmaf(w11[(o + 7) % G::WF], t1.w, a1);
                    lds128_if(t1, tpa + 128, reload);
                    a0 = fmaf(w11[(o + 8) % G::WF], t2.x, a0); a1 = fmaf(w11[(o + 9) % G::WF], t2.y, a1);
                    a0 = fmaf(w11[(o + 10) % G::WF], t2.z, a0); a1 = fmaf(w5[(o + 0) & MP], t2.w, a1);
                    lds128_if(t2, tpa + 256, reload);
                    a0 = fmaf(w5[(o + 1) & MP], t3.x, a0); a1 = fmaf(w5[(o + 2) & MP], t3.y, a1);
                    a0 = fmaf(w5[(o + 3) & MP], t3.z, a0); a1 = fmaf(w5[(o + 4) & MP], t3.w, a1);
                    lds128_if(t3, tpa + 384, reload);
                    }
                    // fresh patch values of the next pixel overwrite the slots this pixel has just consumed
                    // (for the last pixel of an item these loads run up to S columns past the tile: still inside the
                    // padded shared-memory allocation, and the values are never used)
                    const int npix = b0 + b + 1;
                    {
                        const int on = S * (b + 1);
#pragma unroll
                        for (int t = 0; t < G::NEWF; ++t) w11[(on + kFlen - G::NEWF + t) % G::WF] = pf[(S * npix + kFlen - G::NEWF + t) * G::PT];
#pragma unroll
                        for (int t = 0; t < G::NEWP; ++t) w5[(on + 5 - G::NEWP + t) & MP] = base[S * npix * G::PT + off_part[t]];
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
                if (ox < p.ow) {
                    if (sizeof(OutT) == 4 && p.raw_f32) *reinterpret_cast<float*>(drow + (S * ox + px)) = v;
                    else store_px(drow + (S * ox + px), v);
                }
            }
        }
        cur = nxt;
        if (PIPE) {
            __syncwarp();
            if ((tid & 31) == 0) {
                __threadfence_block();                                   // this warp's reads of the buffer are done
                if (atomicAdd(&done_cnt[it & 1], 1) == C::NT / 32 - 1) {  // last warp out refills the buffer
                    done_cnt[it & 1] = 0;
                    __threadfence_block();
                    if (tile + 2 * nworkers < ntiles) {
                        TileCursor t2 = cur;                             // cur already points at the next tile
                        t2.advance(nworkers, p.tiles_x, p.tiles_y);
                        octet_issue_tile_tma<S, !EIG>(&tmap, &hmap, buf, (it & 1) ? bar1 : bar0, t2, type, py, px);
                    }
                }
            }
            continue;
        }
        __syncthreads();   // tile consumed: its buffer may be refilled by the next prefetch
        if (NBUF == 1) {
            if (tile + nworkers < ntiles) octet_issue_tile<S>(p, &tmap, buf0, bar0, nxt, type, py, px);
            cp_async_commit();
        }
    }
    if (!PIPE) cp_async_wait<0>();
}

}  // namespace raisr
