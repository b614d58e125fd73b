// raisr_prep2.cuh -- kernel A of the RAISR path, packed-fp32 version ("prep2").
//
// Same contract, parameters and bit-exact results as prep_kernel (raisr_prep.cuh): bilinear upscale on
// the extended domain (raisr.cl:48-61,198-217), Sobel as flipped convolution (:43-46,235-253), separable
// 9-tap Gaussian structure tensor (:258-276, intended semantics), eigen-solve / quantise / hash
// (:278-317).  What changes is how the arithmetic is issued.  prep_kernel is bound by instruction issue
// (81 % of the slots, FMA pipe ~40 %), and on sm_100a one FFMA2 / FADD2 / FMUL2 (PTX fma/add/sub/mul
// .rn.f32x2) does two IEEE fp32 operations for one issue slot (measured, tools/ffma2_test.cu: same
// 72 TFLOP/s at half the issue rate).  So every stage works on PAIRS of pixels that need exactly the same
// instruction stream: rows r and r + D of the tile (D = half the tile height).  All tile arrays in shared
// memory hold such pairs (row n of the pair array = rows n and n + D of the tile), which keeps every
// operand of every stage pair-aligned -- vertical neighbours r-1, r+1 of a pair are again a pair.  Each
// lane of a pair goes through the same operations in the same order as the scalar kernel, so the results
// are identical bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "raisr_prep.cuh"

namespace raisr {

typedef unsigned long long p2;   // two packed fp32 in a 64-bit register pair: .lo = tile row r, .hi = tile row r + P2_D

__device__ __forceinline__ p2 pk(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(p2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ p2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { p2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 sub2(p2 a, p2 b) { p2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 mul2(p2 a, p2 b) { p2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// c - a*b, one rounding (the -a*b + c step of the sqrt / divide refinements)
__device__ __forceinline__ p2 fnma2(p2 a, p2 b, p2 c) { return fma2(mul2(a, bc(-1.0f)), b, c); }

constexpr int P2_W = 64, P2_H = 80, P2_D = P2_H / 2, P2_THREADS = 256;
constexpr int P2_UW = P2_W + 2 * kMargin;   // 74 columns of U
constexpr int P2_UH = P2_H + 2 * kMargin;   // 90 rows of U
constexpr int P2_NU = P2_UH - P2_D;         // 50 pair rows of U: (n, n + 40)
constexpr int P2_UPITCH = 74;               // p2 per pair row: 37 x 16 B, odd -> 8 consecutive rows hit 8 bank groups
constexpr int P2_HH = P2_H + 2 * kGrad;     // 88 rows of horizontally filtered products
constexpr int P2_NH = P2_HH - P2_D;         // 48 pair rows of H: (m, m + 40)
constexpr int P2_HPITCH = 66;               // 33 x 16 B
constexpr int P2W_H = P2_UH / 2 + 3, P2W_W = P2_UW / 2 + 3, P2W_PITCH = P2W_W + 1;   // source window (S >= 2)
constexpr int P2_RPT = P2_D / 4;            // 10 pair rows per thread in the vertical pass / eigen stage

struct Prep2Smem {
    p2 u[P2_NU * P2_UPITCH];
    union {
        p2 h[3][P2_NH * P2_HPITCH];
        struct {
            p2 h01[2][P2_NH * P2_HPITCH];
            float lut[256];
            float win[P2W_H * P2W_PITCH];
        };
    };
    float colu[P2_UW];
    float2 rowv[P2_UH];      // (v, 1-v)
    int2 colx[P2_UW];        // window-relative x0, x1
    int2 rowy[P2_UH];        // window-relative y0*pitch, y1*pitch
};
static_assert(sizeof(Prep2Smem) <= 113 * 1024 && (256 + P2W_H * P2W_PITCH) * 4 <= P2_NH * P2_HPITCH * 8, "two prep2 CTAs per SM");

// sqrt_rn_guarded of both lanes (same sequence as the scalar helper, see raisr_prep.cuh)
__device__ __forceinline__ p2 sqrt2_guarded(float xl, float xh)
{
    const float cl = fmaxf(xl, 1.0e-30f), ch = fmaxf(xh, 1.0e-30f);
    float rl, rh;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(cl));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(ch));
    const p2 x = pk(cl, ch), r = pk(rl, rh);
    const p2 s = mul2(x, r), h = mul2(r, bc(0.5f));
    float sl, sh;
    upk(fma2(fnma2(s, s, x), h, s), sl, sh);
    if (!(xl >= 1.0e-30f)) sl = xl > 0.0f ? __fsqrt_rn(xl) : 0.0f;
    if (!(xh >= 1.0e-30f)) sh = xh > 0.0f ? __fsqrt_rn(xh) : 0.0f;
    return pk(sl, sh);
}

template <int S, bool DBG, int NQ, bool FROM_U = false>
__global__ void __launch_bounds__(P2_THREADS, 2) prep2_kernel(const PrepParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Prep2Smem& sm = *reinterpret_cast<Prep2Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int tiles_per_frame = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_per_frame * p.n_frames;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int frame = tile / tiles_per_frame;
    const int trem = tile - frame * tiles_per_frame;
    const int by = trem / p.tiles_x, bx = trem - by * p.tiles_x;
    const int tx0 = bx * P2_W;   // first output column of the tile
    const int ty0 = by * P2_H;   // first band-local output row of the tile
    const int ext_w = p.dw + 2 * kMargin, ext_h = p.rows + 2 * kMargin;

    if (FROM_U) {
        const float* uin = p.uext_in + (size_t)frame * p.uext_frame_stride;
        for (int idx = tid; idx < P2_NU * P2_UW; idx += P2_THREADS) {
            const int c = idx / P2_NU, n = idx - c * P2_NU;      // rows fastest: coalesced reads of the column-major plane
            const float* col = uin + (size_t)min(tx0 + c, ext_w - 1) * p.uext_pitch;
            sm.u[n * P2_UPITCH + c] = pk(__ldg(col + min(ty0 + n, ext_h - 1)), __ldg(col + min(ty0 + n + P2_D, ext_h - 1)));
        }
        __syncthreads();
    } else {
    // ---- phase 0a: texel LUT and coordinate tables (raisr.cl:209: divide, then multiply)
    sm.lut[tid] = __fdiv_rn((float)tid, 255.0f);
    if (tid < P2_UW) {
        int xe = tx0 - kMargin + tid;
        float fx = __fmul_rn(__fdiv_rn((float)xe, (float)(p.dw - 1)), (float)(p.sw - 1));
        float fl = floorf(fx);
        int xi = (int)fl;
        sm.colu[tid] = __fsub_rn(fx, fl);
        sm.colx[tid] = make_int2(min(max(xi, 0), p.sw - 1), min(max(xi + 1, 0), p.sw - 1));
    } else if (tid >= 128 && tid < 128 + P2_UH) {
        int r = tid - 128;
        int ye = p.y0 + ty0 - kMargin + r;  // global output row
        float fy = __fmul_rn(__fdiv_rn((float)ye, (float)(p.dh_glob - 1)), (float)(p.sh_glob - 1));
        float fl = floorf(fy);
        int yi = (int)fl;
        float v = __fsub_rn(fy, fl);
        sm.rowv[r] = make_float2(v, __fsub_rn(1.0f, v));
        int a = min(max(yi, 0), p.sh_glob - 1) - p.src_row0;
        int b = min(max(yi + 1, 0), p.sh_glob - 1) - p.src_row0;
        sm.rowy[r] = make_int2(min(max(a, 0), p.src_rows - 1), min(max(b, 0), p.src_rows - 1));
    }
    __syncthreads();
    // ---- phase 0b: make the tables window-relative (x0/y0 are monotone, so first/last bound them)
    const int wx0 = sm.colx[0].x, wy0 = sm.rowy[0].x;
    const int ww = sm.colx[P2_UW - 1].y - wx0 + 1, wh = sm.rowy[P2_UH - 1].y - wy0 + 1;
    __syncthreads();
    if (tid < P2_UW) {
        int2 c = sm.colx[tid];
        sm.colx[tid] = make_int2(c.x - wx0, c.y - wx0);
    } else if (tid >= 128 && tid < 128 + P2_UH) {
        int2 r = sm.rowy[tid - 128];
        sm.rowy[tid - 128] = make_int2((r.x - wy0) * P2W_PITCH, (r.y - wy0) * P2W_PITCH);
    }
    // ---- phase 0c: source window -> float texels (read_imagef UNORM8 decode), one LUT hit per texel
    const uint8_t* src = p.src + (size_t)frame * p.src_frame_stride;
    for (int idx = tid; idx < P2W_H * P2W_W; idx += P2_THREADS) {
        int r = idx / P2W_W, c = idx - r * P2W_W;
        if (r < wh && c < ww) sm.win[r * P2W_PITCH + c] = sm.lut[__ldg(src + (size_t)(wy0 + r) * p.src_pitch + wx0 + c)];
    }
    __syncthreads();

    // ---- phase 1: bilinear upscale of the 90x74 extended tile (raisr.cl:48-61), two rows (n, n+40) at a time.
    // Thread = one column and a run of pair quads; a quad (4 vertically adjacent samples = 16 contiguous bytes
    // of the column-major uext) is written with one 128-bit store by the tile that owns it.
    float* uext = p.uext + (size_t)frame * p.uext_frame_stride;
    if (tid < 3 * P2_UW) {
        const int c = tid % P2_UW, rg = tid / P2_UW;
        const int2 cx = sm.colx[c];
        const float u = sm.colu[c], omu = __fsub_rn(1.0f, u);
        const p2 U2 = bc(u), OMU2 = bc(omu);
        const int ge = tx0 + c;  // extended-domain column
        const bool col_owned = ge < ext_w && min(max(ge - kMargin, 0), p.dw - 1) / P2_W == bx;
        const int lo = (by == 0) ? 0 : ty0 + 4;
        const int hi = (by == p.tiles_y - 1) ? ext_h : ty0 + P2_H + 4;
        const int q0 = rg == 0 ? 0 : (rg == 1 ? 5 : 9), q1 = rg == 0 ? 5 : (rg == 1 ? 9 : 13);   // pair quads [4q, 4q+4)
        float* ucol = uext + (size_t)ge * p.uext_pitch;
#pragma unroll 1
        for (int q = q0; q < q1; ++q) {
            float vl[4], vh[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = min(4 * q + k, P2_NU - 1);
                const int2 yl = sm.rowy[n], yh = sm.rowy[n + P2_D];
                const float2 wl = sm.rowv[n], wh2 = sm.rowv[n + P2_D];
                const p2 p00 = pk(sm.win[yl.x + cx.x], sm.win[yh.x + cx.x]), p01 = pk(sm.win[yl.x + cx.y], sm.win[yh.x + cx.y]);
                const p2 p10 = pk(sm.win[yl.y + cx.x], sm.win[yh.y + cx.x]), p11 = pk(sm.win[yl.y + cx.y], sm.win[yh.y + cx.y]);
                const p2 V = pk(wl.x, wh2.x), OMV = pk(wl.y, wh2.y);
                // ptxas contracts a single-use mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 (one rounding instead of
                // two); the additions are therefore scalar, which it leaves alone, and only the products are packed
                float al, ah, tl, th;
                upk(mul2(mul2(OMU2, OMV), p00), al, ah);
                upk(mul2(mul2(U2, OMV), p01), tl, th);
                al = __fadd_rn(al, tl); ah = __fadd_rn(ah, th);
                upk(mul2(mul2(OMU2, V), p10), tl, th);
                al = __fadd_rn(al, tl); ah = __fadd_rn(ah, th);
                upk(mul2(mul2(U2, V), p11), tl, th);
                al = __fadd_rn(al, tl); ah = __fadd_rn(ah, th);
                vl[k] = al; vh[k] = ah;
                const p2 acc = pk(al, ah);
                if (4 * q + k < P2_NU) sm.u[n * P2_UPITCH + c] = acc;
            }
            if (col_owned) {
                const int le_lo = ty0 + 4 * q, le_hi = le_lo + P2_D;   // band-local extended rows of the two quads
                // rows 40..51 exist both as .hi of pairs 0..11 and as .lo of pairs 40..51: the .hi copy stores them
                if (q < P2_D / 4 && le_lo >= lo && le_lo < hi) *reinterpret_cast<float4*>(ucol + le_lo) = make_float4(vl[0], vl[1], vl[2], vl[3]);
                if (le_hi >= lo && le_hi < hi) *reinterpret_cast<float4*>(ucol + le_hi) = make_float4(vh[0], vh[1], vh[2], vh[3]);
                if (DBG && frame == 0 && p.dbg_u && ge >= kMargin && ge < p.dw + kMargin) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (4 * q + k >= P2_NU) continue;
                        const int a = le_lo + k, b = le_hi + k;
                        if (q < P2_D / 4 && a >= lo && a >= kMargin && a < p.rows + kMargin) p.dbg_u[(size_t)(a - kMargin) * p.dbg_pitch + (ge - kMargin)] = vl[k];
                        if (le_hi >= lo && le_hi < hi && b >= kMargin && b < p.rows + kMargin) p.dbg_u[(size_t)(b - kMargin) * p.dbg_pitch + (ge - kMargin)] = vh[k];
                    }
                }
            }
        }
    }
    __syncthreads();
    }

    // ---- phase 2: Sobel, products, horizontal 9-tap Gaussian on pair rows.  One work item = 8 consecutive
    // output columns of one pair row, streamed over the 16 gradient columns it needs: each new gradient
    // column adds its term to every output that uses it, in ascending tap order (the oracle's order).
    // Lanes of a quarter warp take 8 consecutive pair rows: conflict-free 128-bit loads and stores.
    for (int item = tid; item < P2_NH * (P2_W / 8); item += P2_THREADS) {
        const int q = item / P2_NH, m = item - q * P2_NH;
        const ulonglong2* r0 = reinterpret_cast<const ulonglong2*>(&sm.u[(m + 0) * P2_UPITCH + 8 * q]);
        const ulonglong2* r1 = reinterpret_cast<const ulonglong2*>(&sm.u[(m + 1) * P2_UPITCH + 8 * q]);
        const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(&sm.u[(m + 2) * P2_UPITCH + 8 * q]);
        p2 a[18], b[18], c[18];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            ulonglong2 t = r0[i]; a[2 * i] = t.x; a[2 * i + 1] = t.y;
            t = r1[i]; b[2 * i] = t.x; b[2 * i + 1] = t.y;
            t = r2[i]; c[2 * i] = t.x; c[2 * i + 1] = t.y;
        }
        p2 axx[8], axy[8], ayy[8];
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            const p2 d0 = sub2(a[g], a[g + 2]), d1 = sub2(b[g], b[g + 2]), d2 = sub2(c[g], c[g + 2]);
            const p2 gx = add2(add2(d0, add2(d1, d1)), d2);          // 2*d1 == d1+d1 exactly
            const p2 s0 = add2(add2(a[g], add2(a[g + 1], a[g + 1])), a[g + 2]);
            const p2 s2 = add2(add2(c[g], add2(c[g + 1], c[g + 1])), c[g + 2]);
            const p2 gy = sub2(s0, s2);
            const p2 pxx = mul2(gx, gx), pxy = mul2(gx, gy), pyy = mul2(gy, gy);
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int k = g - o;
                if (k == 0) {
                    axx[o] = mul2(bc(g1c(0)), pxx); axy[o] = mul2(bc(g1c(0)), pxy); ayy[o] = mul2(bc(g1c(0)), pyy);
                } else if (k > 0 && k < 9) {
                    axx[o] = fma2(bc(g1c(k)), pxx, axx[o]); axy[o] = fma2(bc(g1c(k)), pxy, axy[o]); ayy[o] = fma2(bc(g1c(k)), pyy, ayy[o]);
                }
            }
        }
        ulonglong2* dxx = reinterpret_cast<ulonglong2*>(&sm.h[0][m * P2_HPITCH + 8 * q]);
        ulonglong2* dxy = reinterpret_cast<ulonglong2*>(&sm.h[1][m * P2_HPITCH + 8 * q]);
        ulonglong2* dyy = reinterpret_cast<ulonglong2*>(&sm.h[2][m * P2_HPITCH + 8 * q]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dxx[i] = make_ulonglong2(axx[2 * i], axx[2 * i + 1]);
            dxy[i] = make_ulonglong2(axy[2 * i], axy[2 * i + 1]);
            dyy[i] = make_ulonglong2(ayy[2 * i], ayy[2 * i + 1]);
        }
    }
    __syncthreads();

    // ---- phase 3a: vertical 9-tap Gaussian.  Thread = one column, 10 consecutive pair rows: per plane its
    // 18 input pairs are pulled into registers, then (after a barrier) the 10 results overwrite the first
    // rows of its own range in place, so phase 3b can run as a rolled loop.
    const int xo = tid & 63, grp = tid >> 6;
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
        p2* plane = &sm.h[ch][(grp * P2_RPT) * P2_HPITCH + xo];
        p2 in[P2_RPT + 8];
#pragma unroll
        for (int k = 0; k < P2_RPT + 8; ++k) in[k] = plane[k * P2_HPITCH];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < P2_RPT; ++j) {
            p2 mm = mul2(bc(g1c(0)), in[j]);
#pragma unroll
            for (int k = 1; k < 9; ++k) mm = fma2(bc(g1c(k)), in[j + k], mm);
            plane[j * P2_HPITCH] = mm;
        }
    }
    // no barrier needed: phase 3b reads back only what this thread wrote

    // ---- phase 3b: 2x2 eigen-solve, quantise, hash (raisr.cl:278-317), both rows of a pair at once
    {
        const int x = tx0 + xo;
        const float PI_F = 3.14159265358979323846f;
        float sq[NQ], cq[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) { sq[i] = p.sq[i]; cq[i] = p.cq[i]; }
        const int xs = x / S, xt = x % S;
        const bool col_ok = x < p.dw;
        uint8_t* hbase = p.hash + (size_t)frame * p.hash_frame_stride + xs;
        const p2* hp = &sm.h[0][(grp * P2_RPT) * P2_HPITCH + xo];
#pragma unroll 1
        for (int j = 0; j < P2_RPT; ++j) {
            const p2 mb = hp[P2_NH * P2_HPITCH];
            const p2 ma = p.as_written ? mb : hp[0];            // raisr.cl:271 accumulates gx*gy into ma
            const p2 md = hp[2 * P2_NH * P2_HPITCH];
            hp += P2_HPITCH;
            const p2 T = add2(ma, md);
            float dal, dah, dbl, dbh;                       // scalar subtraction: see the note on contraction in phase 1
            upk(mul2(ma, md), dal, dah);
            upk(mul2(mb, mb), dbl, dbh);
            const p2 D = pk(__fsub_rn(dal, dbl), __fsub_rn(dah, dbh));
            const p2 rad = sub2(mul2(mul2(T, T), bc(0.25f)), D);     // (a fused T*T*0.25 - D is the same value: *0.25 is exact)
            float radl, radh;
            upk(rad, radl, radh);
            const p2 sqr = sqrt2_guarded(fmaxf(radl, 0.0f), fmaxf(radh, 0.0f));   // radicand clamped at 0
            const p2 ht = mul2(T, bc(0.5f));
            const p2 L1 = add2(ht, sqr);
            float L1l, L1h, L2l, L2h, mbl, mbh, xl, xh;
            upk(L1, L1l, L1h);
            upk(sub2(ht, sqr), L2l, L2h);
            upk(mb, mbl, mbh);
            upk(sub2(L1, md), xl, xh);
            // folded atan2(mb, L1 - md): same octant reduction + polynomial as folded_atan2()
            float thl, thh;
            {
                const float axl = fabsf(xl), ayl = fabsf(mbl), axh = fabsf(xh), ayh = fabsf(mbh);
                const float mxl = fmaxf(axl, ayl), mnl = fminf(axl, ayl), mxh = fmaxf(axh, ayh), mnh = fminf(axh, ayh);
                float rl, rh;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(mxl));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(mxh));
                float zl, zh;
                upk(mul2(pk(mnl, mnh), pk(rl, rh)), zl, zh);
                zl = (mxl > 1.0e-30f) ? zl : 0.0f;
                zh = (mxh > 1.0e-30f) ? zh : 0.0f;
                const p2 z = pk(zl, zh), w = mul2(z, z);
                p2 pz = bc(-0.004054493736475706f);
                pz = fma2(pz, w, bc(0.021862685680389404f));
                pz = fma2(pz, w, bc(-0.055911920964717865f));
                pz = fma2(pz, w, bc(0.09642166644334793f));
                pz = fma2(pz, w, bc(-0.13908617198467255f));
                pz = fma2(pz, w, bc(0.19946563243865967f));
                pz = fma2(pz, w, bc(-0.33329859375953674f));
                pz = fma2(pz, w, bc(0.9999993443489075f));
                upk(mul2(pz, z), thl, thh);
                if (mnl == mxl) thl = (mxl > 0.0f) ? 0.78539816339744830962f : 0.0f;
                if (mnh == mxh) thh = (mxh > 0.0f) ? 0.78539816339744830962f : 0.0f;
                if (ayl > axl) thl = 1.57079632679489661923f - thl;
                if (ayh > axh) thh = 1.57079632679489661923f - thh;
                if (xl < 0.0f) thl = PI_F - thl;
                if (xh < 0.0f) thh = PI_F - thh;
                if (mbl < 0.0f) thl = PI_F - thl;
                if (mbh < 0.0f) thh = PI_F - thh;
            }
            // coherence = (sqrt(L1) - sqrt(L2)) / (sqrt(L1) + sqrt(L2)), 0 when the denominator is 0
            float col, coh;
            {
                const p2 s1 = sqrt2_guarded(L1l, L1h), s2 = sqrt2_guarded(fmaxf(L2l, 0.0f), fmaxf(L2h, 0.0f));   // L2 clamped at 0
                const p2 den = add2(s1, s2), num = sub2(s1, s2);
                float dl, dh, nl, nh;
                upk(den, dl, dh);
                upk(num, nl, nh);
                float rl, rh;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(dl));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(dh));
                p2 r = pk(rl, rh);
                r = fma2(r, fnma2(den, r, bc(1.0f)), r);
                const p2 qq = mul2(num, r);
                upk(fma2(r, fnma2(den, qq, num), qq), col, coh);
                if (!(dl >= 1.0e-15f && dl <= 1.0e15f)) col = dl != 0.0f ? __fdiv_rn(nl, dl) : 0.0f;
                if (!(dh >= 1.0e-15f && dh <= 1.0e15f)) coh = dh != 0.0f ? __fdiv_rn(nh, dh) : 0.0f;
            }
            // theta / pi with the divide's own fast path (theta is 0 or in [1e-8, pi]; exact for theta = 0)
            float tql, tqh;
            {
                const float PI_INV = 0.31830988618379067154f;
                const float r = __fmaf_rn(PI_INV, __fmaf_rn(-PI_F, PI_INV, 1.0f), PI_INV);
                const p2 th = pk(thl, thh);
                const p2 q = mul2(th, bc(r));
                upk(fma2(bc(r), fma2(bc(-PI_F), q, th), q), tql, tqh);
                if (thl != 0.0f && thl < 1.0e-15f) tql = __fdiv_rn(thl, PI_F);
                if (thh != 0.0f && thh < 1.0e-15f) tqh = __fdiv_rn(thh, PI_F);
            }
            float fal, fah;
            upk(mul2(pk(tql, tqh), bc((float)p.n_angle)), fal, fah);   // == (theta / pi) * n_angle of the oracle
            const int al = min(max((int)fal, 0), p.n_angle - 1), ah = min(max((int)fah, 0), p.n_angle - 1);
            // "first i with value < q[i], else last bin" (raisr.cl:301-314); unused q[i] are -inf
            int sil = p.n_strength - 1, cil = p.n_coherence - 1, sih = sil, cih = cil;
#pragma unroll
            for (int i = NQ - 1; i >= 0; --i) {
                if (L1l < sq[i]) sil = i;
                if (L1h < sq[i]) sih = i;
                if ((p.as_written ? L1l : col) < cq[i]) cil = i;     // raisr.cl:310 compares L1
                if ((p.as_written ? L1h : coh) < cq[i]) cih = i;
            }
            if (p.as_written) sil = sih = 0;                         // raisr.cl:316 leaves strength out of the hash
            const int bl = (al * p.n_strength + sil) * p.n_coherence + cil;
            const int bh = (ah * p.n_strength + sih) * p.n_coherence + cih;
            const int yl_lo = ty0 + grp * P2_RPT + j, yl_hi = yl_lo + P2_D;   // band-local output rows
            if (col_ok) {
                if (yl_lo < p.rows) {
                    const int type = (yl_lo % S) * S + xt;
                    hbase[(size_t)type * p.hash_plane_stride + (size_t)(yl_lo / S) * p.hash_pitch] = (uint8_t)bl;
                    if (DBG && frame == 0) {
                        const size_t o = (size_t)yl_lo * p.dbg_pitch + x;
                        if (p.dbg_hash) p.dbg_hash[o] = bl * (S * S) + type;
                        if (p.dbg_angle) p.dbg_angle[o] = thl;
                        if (p.dbg_l1) p.dbg_l1[o] = L1l;
                        if (p.dbg_coh) p.dbg_coh[o] = col;
                    }
                }
                if (yl_hi < p.rows) {
                    const int type = (yl_hi % S) * S + xt;
                    hbase[(size_t)type * p.hash_plane_stride + (size_t)(yl_hi / S) * p.hash_pitch] = (uint8_t)bh;
                    if (DBG && frame == 0) {
                        const size_t o = (size_t)yl_hi * p.dbg_pitch + x;
                        if (p.dbg_hash) p.dbg_hash[o] = bh * (S * S) + type;
                        if (p.dbg_angle) p.dbg_angle[o] = thh;
                        if (p.dbg_l1) p.dbg_l1[o] = L1h;
                        if (p.dbg_coh) p.dbg_coh[o] = coh;
                    }
                }
            }
        }
    }
    __syncthreads();   // shared memory is reused by the next tile
    }
}

}  // namespace raisr
