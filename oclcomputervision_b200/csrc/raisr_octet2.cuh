// raisr_octet2.cuh -- two-plane variant of the octet filter kernel for the colour path.
//
// The reference applies the luma-derived filter to all four channels of a pixel (raisr.cl:322-330
// accumulates a half4).  Running filter_octet_kernel once per plane streams the same 484 B of taps per
// pixel through shared memory four times; here one CTA filters TWO planes of the same pixels, so the taps
// are fetched once per pixel pair and only the patch windows, accumulators and the butterfly are doubled:
// 2.93 + 2 x 1.9 shared wavefronts for two plane-pixels instead of 2 x 4.83.  The 110 KB fp32 table slice
// plus two double-buffered plane tiles only fit with 64x20 tiles (227 KB), i.e. 40 octets = 320 threads.
// s = 2, fp32 taps, raw (unclamped) float output -- exactly what the colour path needs; everything else
// (record layout, lane geometry, TMA fill, tap-reload skipping) is shared with raisr_octet.cuh.
// Planes are addressed like frames: plane k of the upscaled image at uext + k * uext_frame_stride, plane k
// of the output at dst + k * dst_frame_stride; blockIdx.y selects the plane pair (2y, 2y+1).
#pragma once
#include "raisr_octet.cuh"

namespace raisr {

struct Octet2Cfg { static constexpr int S = 2, OTW = 64, OTH = 20, IW = 32, NT = 320; };

struct Octet2Geom {
    using C = Octet2Cfg;
    static constexpr int S = C::S;
    static constexpr int TUH = S * (C::OTH - 1) + kFlen;
    static constexpr int NCOLS = S * (C::OTW - 1) + kFlen;
    static constexpr int PT = (TUH + 3 + 3) / 4 * 4;
    static constexpr int WF = 16, WP = 8, NEWF = S, NEWP = S;
    static constexpr int SEGS = C::OTW / C::IW;
    static constexpr int NOCT = C::NT / 8;
    static constexpr int TILE_BYTES = NCOLS * PT * 4;                       // one plane (what TMA delivers)
    static constexpr int PLANE_BYTES = (TILE_BYTES + 127) / 128 * 128;      // plane stride in the buffer: TMA destinations are 128-B aligned
    static constexpr int HASH_BYTES = C::OTH * C::OTW;
    static constexpr int HASH_OFF = 2 * PLANE_BYTES;                        // 128-byte aligned: a TMA destination too
    static constexpr int BUF_BYTES = (HASH_OFF + HASH_BYTES + 127) / 128 * 128;          // two planes + hash
    static_assert(C::OTH * SEGS == NOCT && (S * C::OTH) % 4 == 0, "one item per octet; tiles start on row quads");
};

inline size_t octet2_smem_bytes(int n_buckets) { return (size_t)n_buckets * kOctStride * sizeof(float) + 2 * (size_t)Octet2Geom::BUF_BYTES + 32; }

// One thread fills a buffer: the two plane tiles and the hash bytes of the tile, three TMA boxes on one mbarrier.
__device__ __forceinline__ void octet2_issue_tile(const CUtensorMap* tm, const CUtensorMap* hm, unsigned char* buf, unsigned bar,
                                                  const TileCursor& tc, int type, int py, int px, int plane0)
{
    using C = Octet2Cfg;
    using G = Octet2Geom;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, 2 * G::TILE_BYTES + G::HASH_BYTES);
    const int r0 = (C::S * tc.ty * C::OTH + py) & ~3, c0 = C::S * tc.tx * C::OTW + px;
#pragma unroll
    for (int k = 0; k < 2; ++k)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sbase + k * G::PLANE_BYTES), "l"(tm), "r"(r0), "r"(c0), "r"(plane0 + k), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(sbase + G::HASH_OFF), "l"(hm), "r"(tc.tx * C::OTW), "r"(tc.ty * C::OTH), "r"(type), "r"(0), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(Octet2Cfg::NT, 1)
    filter_octet2_kernel(const FilterParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap hmap)
{
    using C = Octet2Cfg;
    using G = Octet2Geom;
    constexpr int S = C::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);
    unsigned char* buf0 = smem_raw + (size_t)p.n_buckets * kOctStride * sizeof(float);
    unsigned char* buf1 = buf0 + G::BUF_BYTES;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(buf1 + G::BUF_BYTES), bar1 = bar0 + 8;
    int* done_cnt = reinterpret_cast<int*>(buf1 + G::BUF_BYTES + 16);   // warps done with buffer 0 / 1
    const int tid = threadIdx.x;
    const int ntypes = S * S;
    const int type = blockIdx.x % ntypes, worker = blockIdx.x / ntypes, nworkers = gridDim.x / ntypes;
    const int plane0 = 2 * blockIdx.y;
    const int py = type / S, px = type % S;
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

    const int ntiles = p.tiles_x * p.tiles_y;
    const int seg = octet / C::OTH, row = octet - seg * C::OTH;
    TileCursor cur, nxt;
    cur.init(min(worker, max(ntiles - 1, 0)), p.tiles_x, p.tiles_y);
    nxt = cur;
    if (tid == 0) {   // the first two tiles fly while the table is copied
        TileCursor pc = cur;
        int pit = 0;
        for (int tile = worker; tile < ntiles && pit < 2; tile += nworkers, ++pit) {
            octet2_issue_tile(&tmap, &hmap, pit ? buf1 : buf0, pit ? bar1 : bar0, pc, type, py, px, plane0);
            pc.advance(nworkers, p.tiles_x, p.tiles_y);
        }
    }
    {
        const float4* g = reinterpret_cast<const float4*>(p.table + (size_t)type * p.n_buckets * kOctStride);
        float4* s = reinterpret_cast<float4*>(tab);
        for (int i = tid; i < p.n_buckets * (kOctStride / 4); i += C::NT) s[i] = __ldg(g + i);
    }
    __syncthreads();   // table slice resident; no CTA-wide barrier after this one (see filter_octet_kernel<PIPE>)
    const int off_full = lane8;
    int off_part[G::NEWP];
    if (lane8 < 6) {
#pragma unroll
        for (int t = 0; t < G::NEWP; ++t) off_part[t] = (8 + lane8 / 2) + (kRunCol0 + 5 * (lane8 % 2) + (5 - G::NEWP) + t) * G::PT;
    } else {
#pragma unroll
        for (int t = 0; t < G::NEWP; ++t) {
            int idx = (lane8 - 6) * G::NEWP + t;
            off_part[t] = (idx < 3 ? (8 + idx) : 10) + kSingleCol * G::PT;
        }
    }
    const float4* tab_lane = reinterpret_cast<const float4*>(tab) + lane8;
    const unsigned tab_lane_s = (unsigned)__cvta_generic_to_shared(tab_lane);
    const unsigned omask = 0xffu << (tid & 24);
    const unsigned maxb = (unsigned)(p.n_buckets - 1);

    int it = 0;
    for (int tile = worker; tile < ntiles; tile += nworkers, ++it) {
        unsigned char* buf = (it & 1) ? buf1 : buf0;
        nxt.advance(nworkers, p.tiles_x, p.tiles_y);
        mbar_wait((it & 1) ? bar1 : bar0, (it >> 1) & 1);

        const int oy = cur.ty * C::OTH + row;
        const int oxs = cur.tx * C::OTW + seg * C::IW;
        if (oy < p.oh && oxs < p.ow) {
            float* drowA = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(p.dst) + (size_t)plane0 * p.dst_frame_stride + (size_t)(S * oy + py) * p.dst_pitch);
            float* drowB = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(drowA) + p.dst_frame_stride);
            const float* baseA = reinterpret_cast<const float*>(buf) + S * (seg * C::IW) * G::PT + S * row + (py & 3);
            const float* baseB = baseA + G::PLANE_BYTES / 4;
            const float* pfA = baseA + off_full;
            const float* pfB = baseB + off_full;
            const uint2* hrow = reinterpret_cast<const uint2*>(buf + G::HASH_OFF + row * C::OTW + seg * C::IW);
            float wA[G::WF], vA[G::WP], wB[G::WF], vB[G::WP];
#pragma unroll
            for (int j = 0; j < G::WF; ++j) { wA[j] = 0.0f; wB[j] = 0.0f; }
#pragma unroll
            for (int j = 0; j < G::WP; ++j) { vA[j] = 0.0f; vB[j] = 0.0f; }
#pragma unroll
            for (int j = 0; j < kFlen; ++j) { wA[j] = pfA[j * G::PT]; wB[j] = pfB[j * G::PT]; }
            if (lane8 < 6) {
                const int o5 = (8 + lane8 / 2) + (kRunCol0 + 5 * (lane8 % 2)) * G::PT;
#pragma unroll
                for (int t = 0; t < 5 - G::NEWP; ++t) { vA[t] = baseA[o5 + t * G::PT]; vB[t] = baseB[o5 + t * G::PT]; }
            }
#pragma unroll
            for (int t = 0; t < G::NEWP; ++t) { vA[5 - G::NEWP + t] = baseA[off_part[t]]; vB[5 - G::NEWP + t] = baseB[off_part[t]]; }
            uint2 hb = hrow[0];
            unsigned bucket = min(hb.x & 0xffu, maxb);
            const float4* tp = tab_lane + bucket * (kOctStride / 4);
            float4 t0 = tp[0], t1 = tp[8], t2 = tp[16], t3 = tp[24];
#pragma unroll 1
            for (int b0 = 0; b0 < C::IW; b0 += 8) {
                if (oxs + b0 >= p.ow) break;
                const uint2 hnext = hrow[min(b0 / 8 + 1, C::IW / 8 - 1)];
                float accA[8], accB[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    unsigned nbucket = (b < 7) ? (((b + 1 < 4 ? hb.x : hb.y) >> (8 * ((b + 1) & 3))) & 0xffu) : (hnext.x & 0xffu);
                    nbucket = min(nbucket, maxb);
                    const bool reload = nbucket != bucket;
                    bucket = nbucket;
                    const unsigned tpa = tab_lane_s + nbucket * (kOctStride * 4);
                    constexpr int MP = G::WP - 1;
                    const int o = S * b;
                    // same accumulation order per plane as filter_octet_kernel (two interleaved chains, summed at the end)
                    float a0 = wA[(o + 0) % G::WF] * t0.x, a1 = wA[(o + 1) % G::WF] * t0.y;
                    float c0 = wB[(o + 0) % G::WF] * t0.x, c1 = wB[(o + 1) % G::WF] * t0.y;
                    a0 = fmaf(wA[(o + 2) % G::WF], t0.z, a0); a1 = fmaf(wA[(o + 3) % G::WF], t0.w, a1);
                    c0 = fmaf(wB[(o + 2) % G::WF], t0.z, c0); c1 = fmaf(wB[(o + 3) % G::WF], t0.w, c1);
                    lds128_if(t0, tpa, reload);
                    a0 = fmaf(wA[(o + 4) % G::WF], t1.x, a0); a1 = fmaf(wA[(o + 5) % G::WF], t1.y, a1);
                    c0 = fmaf(wB[(o + 4) % G::WF], t1.x, c0); c1 = fmaf(wB[(o + 5) % G::WF], t1.y, c1);
                    a0 = fmaf(wA[(o + 6) % G::WF], t1.z, a0); a1 = fmaf(wA[(o + 7) % G::WF], t1.w, a1);
                    c0 = fmaf(wB[(o + 6) % G::WF], t1.z, c0); c1 = fmaf(wB[(o + 7) % G::WF], t1.w, c1);
                    lds128_if(t1, tpa + 128, reload);
                    a0 = fmaf(wA[(o + 8) % G::WF], t2.x, a0); a1 = fmaf(wA[(o + 9) % G::WF], t2.y, a1);
                    c0 = fmaf(wB[(o + 8) % G::WF], t2.x, c0); c1 = fmaf(wB[(o + 9) % G::WF], t2.y, c1);
                    a0 = fmaf(wA[(o + 10) % G::WF], t2.z, a0); a1 = fmaf(vA[(o + 0) & MP], t2.w, a1);
                    c0 = fmaf(wB[(o + 10) % G::WF], t2.z, c0); c1 = fmaf(vB[(o + 0) & MP], t2.w, c1);
                    lds128_if(t2, tpa + 256, reload);
                    a0 = fmaf(vA[(o + 1) & MP], t3.x, a0); a1 = fmaf(vA[(o + 2) & MP], t3.y, a1);
                    c0 = fmaf(vB[(o + 1) & MP], t3.x, c0); c1 = fmaf(vB[(o + 2) & MP], t3.y, c1);
                    a0 = fmaf(vA[(o + 3) & MP], t3.z, a0); a1 = fmaf(vA[(o + 4) & MP], t3.w, a1);
                    c0 = fmaf(vB[(o + 3) & MP], t3.z, c0); c1 = fmaf(vB[(o + 4) & MP], t3.w, c1);
                    lds128_if(t3, tpa + 384, reload);
                    const int npix = b0 + b + 1;
                    if (npix < C::IW) {
                        const int on = S * (b + 1);
#pragma unroll
                        for (int t = 0; t < G::NEWF; ++t) {
                            wA[(on + kFlen - G::NEWF + t) % G::WF] = pfA[(S * npix + kFlen - G::NEWF + t) * G::PT];
                            wB[(on + kFlen - G::NEWF + t) % G::WF] = pfB[(S * npix + kFlen - G::NEWF + t) * G::PT];
                        }
#pragma unroll
                        for (int t = 0; t < G::NEWP; ++t) {
                            vA[(on + 5 - G::NEWP + t) & MP] = baseA[S * npix * G::PT + off_part[t]];
                            vB[(on + 5 - G::NEWP + t) & MP] = baseB[S * npix * G::PT + off_part[t]];
                        }
                    }
                    accA[b] = a0 + a1;
                    accB[b] = c0 + c1;
                }
                hb = hnext;
                const bool h2 = lane8 & 4, h1 = lane8 & 2, h0 = lane8 & 1;
                auto butterfly = [&](const float (&acc)[8]) {
                    float r4[4], r2[2];
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
                    return keep + __shfl_xor_sync(omask, send, 1);
                };
                const float vAo = butterfly(accA), vBo = butterfly(accB);
                const int ox = oxs + b0 + lane8;
                if (ox < p.ow) {
                    drowA[S * ox + px] = vAo;
                    drowB[S * ox + px] = vBo;
                }
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
                    octet2_issue_tile(&tmap, &hmap, buf, (it & 1) ? bar1 : bar0, t2, type, py, px, plane0);
                }
            }
        }
    }
}

}  // namespace raisr
