// raisr_duo.cuh -- kernel B for s = 2 with 24-bit tap records: one CTA filters BOTH pixel types of an output row.
//
// Same contract as filter_octet_kernel (raisr_octet.cuh; raisr.cl:316-337 in fp32): eight lanes per pixel, the
// lane-major 384-byte b24 records, the column-major U tile delivered by TMA, the barrier-free tile pipeline.  What
// changes is the walk.  filter_octet_kernel binds a CTA to ONE pixel type, so an octet steps two tile columns per
// pixel and every lane fetches 2 + 2 fresh patch values per pixel.  The 24-bit records shrink a type's table slice to
// 83 KB, so the slices of the two types of one output row (py fixed, px = 0 / 1) fit together (166 KB); the CTA is
// bound to a row type and an octet walks DENSE output pixels of its row, alternating between the two slices: the
// patch window slides one column per pixel, each lane fetches 1 + 1 fresh values (plus one more on the two lanes that
// hold the three leftover taps), the U tile has half the halo per pixel, and the eight results of a butterfly are
// eight adjacent bytes of the output row.  Shared-memory wavefronts per pixel: 4.16 -> ~3.6.
#pragma once
#include "packed_f32.cuh"
#include "raisr_octet.cuh"

namespace raisr {

struct DuoCfg { static constexpr int S = 2, DW = 128, OTH = 16, IW = 32, NT = 512; };   // dense columns x own rows per tile

struct DuoGeom {
    using C = DuoCfg;
    static constexpr int TUH = 2 * (C::OTH - 1) + kFlen;              // 41 tile rows needed
    static constexpr int NCOLS = C::DW + kFlen - 1;                   // 138 tile columns
    static constexpr int PT = 48;                                     // floats per tile column: >= TUH + 1, a multiple of 4 (TMA), and
                                                                      // 5 * PT = 16 (mod 32): the two 5-runs of a row sit half a bank cycle apart
    static constexpr int SEGS = C::DW / C::IW;                        // 4
    static constexpr int NOCT = C::NT / 8;                            // 64 = OTH * SEGS
    static constexpr int TILE_BYTES = NCOLS * PT * 4;
    static constexpr int HASH_PLANE = C::OTH * (C::DW / 2);           // bytes per pixel type: 16 rows x 64 own columns
    static constexpr int HASH_OFF = (TILE_BYTES + 127) / 128 * 128;
    static constexpr int BUF_BYTES = (HASH_OFF + 2 * HASH_PLANE + 127) / 128 * 128;
    static_assert(C::OTH * SEGS == NOCT && PT >= TUH + 1 && PT % 4 == 0 && (2 * C::OTH) % 4 == 0, "tile geometry");
};

inline size_t duo_smem_bytes(int n_buckets)
{
    return 2 * (size_t)n_buckets * kOctBytesB24 + 2 * (size_t)DuoGeom::BUF_BYTES + 32 + kOctetTailPad;
}

// One thread fills a buffer: the U tile and the hash bytes of both pixel types of the tile, three TMA boxes on one mbarrier.
__device__ __forceinline__ void duo_issue_tile(const CUtensorMap* tm, const CUtensorMap* hm, unsigned char* buf, unsigned bar,
                                               const TileCursor& tc, int py)
{
    using C = DuoCfg;
    using G = DuoGeom;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, G::TILE_BYTES + 2 * G::HASH_PLANE);
    const int r0 = 2 * tc.ty * C::OTH, c0 = tc.tx * C::DW;           // the patch of dense pixel (y, x) starts at extended (y, x)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(sbase), "l"(tm), "r"(r0), "r"(c0), "r"(tc.frame), "r"(bar) : "memory");
#pragma unroll
    for (int px = 0; px < 2; ++px)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(sbase + G::HASH_OFF + px * G::HASH_PLANE), "l"(hm), "r"(tc.tx * (C::DW / 2)), "r"(tc.ty * C::OTH),
                       "r"(2 * py + px), "r"(tc.frame), "r"(bar) : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(DuoCfg::NT, 1)
    filter_duo_kernel(const FilterParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap hmap)
{
    using C = DuoCfg;
    using G = DuoGeom;
    constexpr int REC = kOctBytesB24;
    constexpr int WF = 16, WP = 8, MP = WP - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* tab = smem_raw;                                    // [px][bucket][384 B]
    const unsigned slice = (unsigned)p.n_buckets * REC;
    unsigned char* buf0 = smem_raw + 2 * (size_t)slice;
    unsigned char* buf1 = buf0 + G::BUF_BYTES;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(buf1 + G::BUF_BYTES), bar1 = bar0 + 8;
    int* done_cnt = reinterpret_cast<int*>(buf1 + G::BUF_BYTES + 16);
    const int tid = threadIdx.x;
    const int py = blockIdx.x & 1, worker = blockIdx.x >> 1, nworkers = gridDim.x >> 1;
    const int lane8 = tid & 7, octet = tid >> 3;
    if ((__cvta_generic_to_shared(tab) & 127) != 0) __trap();
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        done_cnt[0] = 0;
        done_cnt[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int ntiles = p.tiles_x * p.tiles_y * p.n_frames;
    const int seg = octet / C::OTH, row = octet - seg * C::OTH;       // a warp = 4 consecutive own rows of one segment
    const int dw = 2 * p.ow;
    TileCursor cur, nxt;
    cur.init(min(worker, max(ntiles - 1, 0)), p.tiles_x, p.tiles_y);
    nxt = cur;
    if (tid == 0) {
        TileCursor pc = cur;
        int pit = 0;
        for (int tile = worker; tile < ntiles && pit < 2; tile += nworkers, ++pit) {
            duo_issue_tile(&tmap, &hmap, pit ? buf1 : buf0, pit ? bar1 : bar0, pc, py);
            pc.advance(nworkers, p.tiles_x, p.tiles_y);
        }
    }
    {   // the table slices of pixel types (py, 0) and (py, 1) are adjacent in the [type][bucket] table
        const float4* g = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(p.table) + (size_t)(2 * py) * slice);
        float4* s = reinterpret_cast<float4*>(tab);
        for (int i = tid; i < 2 * p.n_buckets * (REC / 16); i += C::NT) s[i] = __ldg(g + i);
    }
    __syncthreads();                                                   // the only CTA-wide barrier

    // Lane geometry (floats from the patch origin in the column-major tile).  Slots 0..10 of a lane are filter row
    // lane8; slots 11..15: lanes 0..5 a 5-run of rows 8..10, lanes 6, 7 the leftover taps of column kSingleCol in slots
    // 14 (lane 6: row 8, lane 7: row 10) and 15 (lane 6: row 9) -- the s = 2 record of octet_pack_filter_s.
    int off_b, off_x;                                                  // fresh element of slot 15, and of slot 14 (lanes 6, 7 only)
    if (lane8 < 6) {
        off_b = (8 + lane8 / 2) + (kRunCol0 + 5 * (lane8 % 2) + 4) * G::PT;
        off_x = off_b;
    } else {
        off_x = (lane8 == 6 ? 8 : 10) + kSingleCol * G::PT;
        off_b = (lane8 == 6 ? 9 : 10) + kSingleCol * G::PT;
    }
    const bool two = lane8 >= 6;
    const unsigned tab_lane_s = (unsigned)__cvta_generic_to_shared(tab) + 16u * lane8;
    const unsigned omask = 0xffu << (tid & 24);

    int it = 0;
    for (int tile = worker; tile < ntiles; tile += nworkers, ++it) {
        unsigned char* buf = (it & 1) ? buf1 : buf0;
        nxt.advance(nworkers, p.tiles_x, p.tiles_y);
        mbar_wait((it & 1) ? bar1 : bar0, (it >> 1) & 1);

        const int oy = cur.ty * C::OTH + row;
        const int xs = cur.tx * C::DW + seg * C::IW;                  // first dense column of the item
        if (oy < p.oh && xs < dw) {                                   // octet-uniform
            OutT* drow = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(p.dst) + (size_t)cur.frame * p.dst_frame_stride +
                                                 (size_t)(2 * oy + py) * p.dst_pitch);
            const float* base = reinterpret_cast<const float*>(buf) + (seg * C::IW) * G::PT + 2 * row + py;
            const float* pf = base + lane8;
            const unsigned char* hbase = buf + G::HASH_OFF + row * (C::DW / 2) + seg * (C::IW / 2);
            float w11[WF], w5[WP];
#pragma unroll
            for (int j = 0; j < WF; ++j) w11[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < WP; ++j) w5[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < kFlen; ++j) w11[j] = pf[j * G::PT];
            if (lane8 < 6) {
                const float* pp = base + (8 + lane8 / 2) + (kRunCol0 + 5 * (lane8 % 2)) * G::PT;
#pragma unroll
                for (int t = 0; t < 5; ++t) w5[t] = pp[t * G::PT];
            } else {
                w5[3] = base[off_x];
                w5[4] = base[off_b];
            }
            uint2 h0 = *reinterpret_cast<const uint2*>(hbase), h1 = *reinterpret_cast<const uint2*>(hbase + G::HASH_PLANE);
            // One set of packed taps PER PIXEL TYPE: adjacent dense pixels belong to different types, so a single set would
            // be reloaded for every pixel; with two sets a type's taps are reloaded only when that type's bucket changes
            // (27 % of horizontally adjacent same-type pixels share a bucket, as in filter_octet_kernel).
            unsigned recA = h0.x & 0xffu, recB = (h1.x & 0xffu) + (unsigned)p.n_buckets;   // record index: px * n_buckets + bucket
            uint4 qa0, qa1, qa2, qb0, qb1, qb2;
            {
                const uint4* tp = reinterpret_cast<const uint4*>(tab + recA * REC) + lane8;
                qa0 = tp[0]; qa1 = tp[8]; qa2 = tp[16];
                tp = reinterpret_cast<const uint4*>(tab + recB * REC) + lane8;
                qb0 = tp[0]; qb1 = tp[8]; qb2 = tp[16];
            }
#pragma unroll 1
            for (int b0 = 0; b0 < C::IW; b0 += 16) {
                if (xs + b0 >= dw) break;                             // octet-uniform
                const int nb16 = min(b0 / 16 + 1, C::IW / 16 - 1);    // hash bytes of the next batch of 16 dense pixels (8 per type)
                const uint2 h0n = *reinterpret_cast<const uint2*>(hbase + 8 * nb16), h1n = *reinterpret_cast<const uint2*>(hbase + G::HASH_PLANE + 8 * nb16);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float acc[8];
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const int i = 8 * half + b;                   // pixel of the batch; its type px = i & 1, own column i >> 1
                        uint4& q0 = (i & 1) ? qb0 : qa0;
                        uint4& q1 = (i & 1) ? qb1 : qa1;
                        uint4& q2 = (i & 1) ? qb2 : qa2;
                        unsigned& rec = (i & 1) ? recB : recA;
                        // record of the next pixel of the same type (own column (i >> 1) + 1)
                        unsigned nrec;
                        {
                            const int k = (i >> 1) + 1;
                            const unsigned word = (i & 1) ? (k < 4 ? h1.x : (k < 8 ? h1.y : h1n.x)) : (k < 4 ? h0.x : (k < 8 ? h0.y : h0n.x));
                            nrec = __byte_perm(word, 0u, 0x4440u | (k & 3));
                            if (i & 1) nrec += (unsigned)p.n_buckets;
                        }
                        const bool reload = nrec != rec;
                        rec = nrec;
                        const unsigned tpa = tab_lane_s + nrec * REC;
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
                        // element j of pixel i lives in slot (i + j) % W
                        // the two chains of filter_octet_kernel's b24 path, in its order (scalar FMAs: a dense walk puts the
                        // window pairs of odd pixels on odd register boundaries, and packing them costs more moves than it saves)
                        float a0 = w11[(i + 0) % WF] * c0, a1 = w11[(i + 1) % WF] * c1;
                        a0 = fmaf(w11[(i + 2) % WF], c2, a0); a1 = fmaf(w11[(i + 3) % WF], c3, a1);
                        a0 = fmaf(w11[(i + 4) % WF], c4, a0); a1 = fmaf(w11[(i + 5) % WF], c5, a1);
                        a0 = fmaf(w11[(i + 6) % WF], c6, a0); a1 = fmaf(w11[(i + 7) % WF], c7, a1);
                        a0 = fmaf(w11[(i + 8) % WF], c8, a0); a1 = fmaf(w11[(i + 9) % WF], c9, a1);
                        a0 = fmaf(w5[(i + 0) & MP], c11, a0); a1 = fmaf(w5[(i + 1) & MP], c12, a1);
                        a0 = fmaf(w5[(i + 2) & MP], c13, a0); a1 = fmaf(w5[(i + 3) & MP], c14, a1);
                        a0 = fmaf(w11[(i + 10) % WF], c10, a0); a1 = fmaf(w5[(i + 4) & MP], c15, a1);
                        // fresh patch values of the next pixel overwrite slots this pixel has consumed (for the last pixel
                        // of an item they come from one column past it: inside the padded allocation, never used)
                        const int n = b0 + i + 1;
                        w11[(i + 1 + 10) % WF] = pf[(n + kFlen - 1) * G::PT];
                        if (two) w5[(i + 1 + 3) & MP] = base[off_x + n * G::PT];
                        w5[(i + 1 + 4) & MP] = base[off_b + n * G::PT];
                        acc[b] = a0 + a1;
                    }
                    // transposing butterfly: lane q of the octet ends with the sum of pixel b0 + 8*half + q
                    float r4[4], r2[2];
                    const bool s2 = lane8 & 4, s1 = lane8 & 2, s0 = lane8 & 1;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float send = s2 ? acc[k] : acc[k + 4];
                        const float keep = s2 ? acc[k + 4] : acc[k];
                        r4[k] = keep + __shfl_xor_sync(omask, send, 4);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float send = s1 ? r4[k] : r4[k + 2];
                        const float keep = s1 ? r4[k + 2] : r4[k];
                        r2[k] = keep + __shfl_xor_sync(omask, send, 2);
                    }
                    const float send = s0 ? r2[0] : r2[1];
                    const float keep = s0 ? r2[1] : r2[0];
                    const float v = keep + __shfl_xor_sync(omask, send, 1);
                    const int x = xs + b0 + 8 * half + lane8;
                    if (x < dw) {
                        if (sizeof(OutT) == 4 && p.raw_f32) *reinterpret_cast<float*>(drow + x) = v;
                        else store_px(drow + x, v);
                    }
                }
                h0 = h0n;
                h1 = h1n;
            }
        }
        cur = nxt;
        __syncwarp();
        if ((tid & 31) == 0) {
            __threadfence_block();                                   // this warp's reads of the buffer are done
            if (atomicAdd(&done_cnt[it & 1], 1) == C::NT / 32 - 1) {  // last warp out refills the buffer
                done_cnt[it & 1] = 0;
                __threadfence_block();
                if (tile + 2 * nworkers < ntiles) {
                    TileCursor t2 = cur;
                    t2.advance(nworkers, p.tiles_x, p.tiles_y);
                    duo_issue_tile(&tmap, &hmap, buf, (it & 1) ? bar1 : bar0, t2, py);
                }
            }
        }
    }
}

}  // namespace raisr
