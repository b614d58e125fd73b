// raisr_prep2.cuh -- kernel A of the RAISR path with packed fp32 arithmetic ("prep2").
//
// Same contract, parameters, tile (64x56, 72.5 KB, three CTAs per SM) and bit-exact results as prep_kernel
// (raisr_prep.cuh): bilinear upscale on the extended domain (raisr.cl:48-61,198-217), Sobel as flipped
// convolution (:43-46,235-253), separable 9-tap Gaussian structure tensor (:258-276, intended semantics),
// eigen-solve / quantise / hash (:278-317).  What changes is how the arithmetic is issued.  prep_kernel is
// bound by instruction issue (81 % of the slots, FMA pipe ~40 %), and on sm_100a one FFMA2 / FADD2 / FMUL2
// (PTX fma/add/sub/mul.rn.f32x2) does two IEEE fp32 operations for one issue slot (measured,
// tools/ffma2_test.cu: the same 72 TFLOP/s at half the issue rate).  Three of the four stages are
// therefore run on PAIRS of pixels that need exactly the same instruction stream:
//   phase 1  (bilinear)        two vertically adjacent samples of a column
//   phase 3a (vertical pass)   two horizontally adjacent columns (one 64-bit shared load = one pair)
//   phase 3b (eigen / hash)    the same two columns
// Each lane of a pair goes through the same operations in the same order as the scalar kernel, so the
// results are identical bit for bit.  Phase 2 (Sobel + horizontal pass) stays scalar: its operands sit
// at odd and even column offsets alike, which no pair layout serves without extra moves.
// (A first version paired rows r and r+40 of a 64x80 tile through every stage; it needed 108 KB, ran two
// CTAs per SM and was slower than the scalar kernel: FFMA2 holds the FMA pipe for two cycles, so a phase
// made only of packed FMAs is pipe-bound unless CTAs in other phases fill the gaps.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "packed_f32.cuh"
#include "raisr_prep.cuh"

namespace raisr {

// sqrt_rn_guarded of both lanes (same values as the scalar helper in raisr_prep.cuh: negative / NaN -> 0, else the
// correctly rounded root).  The refinement is run on x itself with rsqrt(max(x, 1e-30)): that is the scalar fast
// path for x >= 1e-30 and yields exactly 0 for x == 0 (s = 0, residual 0), so the very common zeros of flat
// regions need no branch; only 0 < x < 1e-30 (one unsigned compare per lane) takes the IEEE square root.
__device__ __forceinline__ p2 sqrt2_guarded(float xl, float xh)
{
    xl = fmaxf(xl, 0.0f);
    xh = fmaxf(xh, 0.0f);
    float rl, rh;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(fmaxf(xl, 1.0e-30f)));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(fmaxf(xh, 1.0e-30f)));
    const p2 x = pk(xl, xh), r = pk(rl, rh);
    const p2 s = mul2(x, r), h = mul2(r, bc(0.5f));
    float sl, sh;
    upk(fma2(fnma2(s, s, x), h, s), sl, sh);
    const unsigned tiny = __float_as_uint(1.0e-30f) - 1u;     // (bits - 1) < tiny  <=>  0 < x < 1e-30
    const bool slow_l = (__float_as_uint(xl) - 1u) < tiny, slow_h = (__float_as_uint(xh) - 1u) < tiny;
    if (slow_l | slow_h) {
        if (slow_l) sl = __fsqrt_rn(xl);
        if (slow_h) sh = __fsqrt_rn(xh);
    }
    return pk(sl, sh);
}

// w_k = dot((1, u, u^2, u^3), cubic_matrix[k]) of raisr.cl:63-68,79-98, evaluated left to right
__device__ __forceinline__ float4 cubic_sample_weights(float u)
{
    const float u2 = __fmul_rn(u, u), u3 = __fmul_rn(u2, u);
    auto w = [&](float m0, float m1, float m2, float m3) {
        float acc = __fadd_rn(__fmul_rn(1.0f, m0), __fmul_rn(u, m1));
        acc = __fadd_rn(acc, __fmul_rn(u2, m2));
        return __fadd_rn(acc, __fmul_rn(u3, m3));
    };
    return make_float4(w(0.0f, -0.5f, 1.0f, -0.5f), w(1.0f, 0.0f, -2.5f, 1.5f), w(0.0f, 0.5f, 2.0f, -1.5f), w(0.0f, 0.0f, -0.5f, 0.5f));
}

constexpr int P2_RPT = PT_H / 8;   // 7 rows per thread in the vertical pass / eigen stage (thread = 2 columns)
constexpr int P2_D = PH_H / 2;     // 32: phases 1-2 pair tile rows r and r + 32
constexpr int P2_NU = PU_H - P2_D; // 34 pair rows of U: (n, n + 32), n = 0..33 (rows 32, 33 appear twice)
constexpr int P2_UPITCH = 74;      // p2 per pair row: 37 x 16 B, odd -> 8 consecutive rows hit 8 bank groups

// source window incl. the bicubic ring and up to 3 columns of slack for a 4-byte aligned start; rows of 48 floats
constexpr int P2W_H = PW_H + 2, P2W_W = PW_W + 2 + 3, P2W_PITCH = 48;
static_assert(P2W_W <= P2W_PITCH && P2W_PITCH % 4 == 0, "window rows hold the aligned window and take 16-byte stores");

// Tables of the bicubic stage-1 variant; they live in the first H plane, which is idle until phase 2.
struct CubicTabs {
    float4 colw[PU_W];     // x weights of the 4 taps
    int4 colx[PU_W];       // window-relative x of the 4 taps
    float4 roww[PU_H];
    int4 rowy[PU_H];       // window-relative y * pitch
};

// As PrepSmem, but the upscaled tile is stored as row pairs (n, n + 32) so that phase 2 can run packed.
struct Prep2Smem {
    p2 u2[P2_NU * P2_UPITCH];
    union {
        float h[3][PH_H * PH_PITCH];
        struct {
            float h01[2][PH_H * PH_PITCH];
            float lut[256];
            float win[P2W_H * P2W_PITCH];
        };
    };
    float colu[PU_W];
    float2 rowv[PU_H];      // (v, 1-v)
    int2 colx[PU_W];        // window-relative x0, x1
    int2 rowy[PU_H];        // window-relative y0*PW_PITCH, y1*PW_PITCH
};
static_assert(sizeof(Prep2Smem) <= 75 * 1024 && 256 + P2W_H * P2W_PITCH <= PH_H * PH_PITCH && sizeof(CubicTabs) <= sizeof(float) * PH_H * PH_PITCH,
              "three prep2 CTAs per SM; phase 0-1 scratch fits the idle H planes");

// CUBIC: stage 1 is the reference's cubic_sample instead of linear_sample (a template parameter so that the
// default kernel carries none of its registers).
template <int S, bool DBG, int NQ, bool FROM_U = false, bool CUBIC = false>
__global__ void __launch_bounds__(PT_THREADS, 3) prep2_kernel(const PrepParams p)
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
    const int tx0 = bx * PT_W;   // first output column of the tile
    const int ty0 = by * PT_H;   // first band-local output row of the tile

    if (FROM_U) {
        const float* uin = p.uext_in + (size_t)frame * p.uext_frame_stride;
        const int ext_w = p.dw + 2 * kMargin, ext_h = p.rows + 2 * kMargin;
        for (int idx = tid; idx < P2_NU * PU_W; idx += PT_THREADS) {
            const int c = idx / P2_NU, n = idx - c * P2_NU;      // rows fastest: coalesced reads of the column-major plane
            const float* col = uin + (size_t)min(tx0 + c, ext_w - 1) * p.uext_pitch;
            sm.u2[n * P2_UPITCH + c] = pk(__ldg(col + min(ty0 + n, ext_h - 1)), __ldg(col + min(ty0 + n + P2_D, ext_h - 1)));
        }
        __syncthreads();
    } else {
    // ---- phase 0a: texel LUT and coordinate tables (raisr.cl:209: divide, then multiply)
    CubicTabs& ct = *reinterpret_cast<CubicTabs*>(&sm.h01[0][0]);
    sm.lut[tid] = __fdiv_rn((float)tid, 255.0f);
    if (tid < PU_W) {
        int xe = tx0 - kMargin + tid;
        float fx = __fmul_rn(__fdiv_rn((float)xe, (float)(p.dw - 1)), (float)(p.sw - 1));
        float fl = floorf(fx);
        int xi = (int)fl;
        if (CUBIC) {
            ct.colw[tid] = cubic_sample_weights(__fsub_rn(fx, fl));
            ct.colx[tid] = make_int4(min(max(xi - 1, 0), p.sw - 1), min(max(xi, 0), p.sw - 1), min(max(xi + 1, 0), p.sw - 1), min(max(xi + 2, 0), p.sw - 1));
        } else {
            sm.colu[tid] = __fsub_rn(fx, fl);
            sm.colx[tid] = make_int2(min(max(xi, 0), p.sw - 1), min(max(xi + 1, 0), p.sw - 1));
        }
    } else if (tid >= 128 && tid < 128 + PU_H) {
        int r = tid - 128;
        int ye = p.y0 + ty0 - kMargin + r;  // global output row
        float fy = __fmul_rn(__fdiv_rn((float)ye, (float)(p.dh_glob - 1)), (float)(p.sh_glob - 1));
        float fl = floorf(fy);
        int yi = (int)fl;
        float v = __fsub_rn(fy, fl);
        auto srow = [&](int y) { return min(max(min(max(y, 0), p.sh_glob - 1) - p.src_row0, 0), p.src_rows - 1); };
        if (CUBIC) {
            ct.roww[r] = cubic_sample_weights(v);
            ct.rowy[r] = make_int4(srow(yi - 1), srow(yi), srow(yi + 1), srow(yi + 2));
        } else {
            sm.rowv[r] = make_float2(v, __fsub_rn(1.0f, v));
            sm.rowy[r] = make_int2(srow(yi), srow(yi + 1));
        }
    }
    __syncthreads();
    // ---- phase 0b: make the tables window-relative (x0/y0 are monotone, so first/last bound them)
    // With 4-byte aligned source rows the window starts on a word boundary and is fetched with 32-bit loads (four
    // texels each, SURVEY.md 8(a7): no per-byte global loads); otherwise byte by byte.
    const uint8_t* src = p.src + (size_t)frame * p.src_frame_stride;
    const bool words = ((reinterpret_cast<uintptr_t>(src) | p.src_pitch) & 3) == 0;
    const int wx0 = (CUBIC ? ct.colx[0].x : sm.colx[0].x) & (words ? ~3 : ~0), wy0 = CUBIC ? ct.rowy[0].x : sm.rowy[0].x;
    const int ww = (CUBIC ? ct.colx[PU_W - 1].w : sm.colx[PU_W - 1].y) - wx0 + 1;
    const int wh = (CUBIC ? ct.rowy[PU_H - 1].w : sm.rowy[PU_H - 1].y) - wy0 + 1;
    __syncthreads();
    if (tid < PU_W) {
        // window-relative BYTE offsets: a texel address is then one integer add away (shared base + row + column)
        if (CUBIC) { int4 c = ct.colx[tid]; ct.colx[tid] = make_int4(4 * (c.x - wx0), 4 * (c.y - wx0), 4 * (c.z - wx0), 4 * (c.w - wx0)); }
        else { int2 c = sm.colx[tid]; sm.colx[tid] = make_int2(4 * (c.x - wx0), 4 * (c.y - wx0)); }
    } else if (tid >= 128 && tid < 128 + PU_H) {
        if (CUBIC) {
            int4 r = ct.rowy[tid - 128];
            ct.rowy[tid - 128] = make_int4((r.x - wy0) * (4 * P2W_PITCH), (r.y - wy0) * (4 * P2W_PITCH), (r.z - wy0) * (4 * P2W_PITCH), (r.w - wy0) * (4 * P2W_PITCH));
        } else {
            int2 r = sm.rowy[tid - 128];
            sm.rowy[tid - 128] = make_int2((r.x - wy0) * (4 * P2W_PITCH), (r.y - wy0) * (4 * P2W_PITCH));
        }
    }
    // ---- phase 0c: source window -> float texels (read_imagef UNORM8 decode), one LUT hit per texel
    if (words) {   // 16 words per window row (<= 12 of them active), sixteen rows per pass
        const int c4 = tid & 15;
        if (4 * c4 < ww) {
            const uint8_t* sp = src + (size_t)(wy0 + (tid >> 4)) * p.src_pitch + wx0 + 4 * c4;
            float* wp = &sm.win[(tid >> 4) * P2W_PITCH + 4 * c4];
            for (int r = tid >> 4; r < wh; r += 16, sp += 16 * p.src_pitch, wp += 16 * P2W_PITCH) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(sp));
                *reinterpret_cast<float4*>(wp) = make_float4(sm.lut[w & 0xffu], sm.lut[(w >> 8) & 0xffu], sm.lut[(w >> 16) & 0xffu], sm.lut[w >> 24]);
            }
        }
    } else {       // 64 threads per window row, four rows per pass: no index division
        const int c = tid & 63;
        if (c < ww) {
            const uint8_t* sp = src + (size_t)(wy0 + (tid >> 6)) * p.src_pitch + wx0 + c;
            float* wp = &sm.win[(tid >> 6) * P2W_PITCH + c];
            for (int r = tid >> 6; r < wh; r += 4, sp += 4 * p.src_pitch, wp += 4 * P2W_PITCH) *wp = sm.lut[__ldg(sp)];
        }
    }
    __syncthreads();

    // ---- phase 1: bilinear upscale of the 66x74 extended tile (raisr.cl:48-61), tile rows n and n + 32 at a
    // time.  Thread = one column, a run of pair quads; each quad (4 vertically adjacent samples = 16 contiguous
    // bytes of the column-major uext) is written with one 128-bit store by the tile that owns it.
    // ptxas contracts a single-use mul.rn.f32x2 that feeds add.rn.f32x2 into FFMA2 (one rounding instead of
    // two), so only the products are packed; the additions stay scalar, which it leaves alone.
    float* uext = p.uext + (size_t)frame * p.uext_frame_stride;
    if (tid < 3 * PU_W) {
        const int ext_w = p.dw + 2 * kMargin, ext_h = p.rows + 2 * kMargin;
        const int c = tid % PU_W, rg = tid / PU_W;
        const int2 cx = sm.colx[c];
        const float u = sm.colu[c], omu = __fsub_rn(1.0f, u);
        const p2 U2 = bc(u), OMU2 = bc(omu);
        auto W = [&](int byte_off) { return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(sm.win) + byte_off); };
        const int ge = tx0 + c;  // extended-domain column
        const bool col_owned = ge < ext_w && min(max(ge - kMargin, 0), p.dw - 1) / PT_W == bx;
        const int lo = (by == 0) ? 0 : ty0 + 4;
        const int hi = (by == p.tiles_y - 1) ? ext_h : ty0 + PT_H + 4;
        float* ucol = uext + (size_t)ge * p.uext_pitch;
#pragma unroll 1
        for (int q = 3 * rg; q < 3 * rg + 3; ++q) {   // pair quads: rows [4q, 4q+4) and [4q+32, 4q+36)
            float vl[4], vh[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = min(4 * q + k, P2_NU - 1);
                if (CUBIC) {   // raisr.cl:99-105: acc += (pix * xweight[j]) * yweight[i], i outer, j inner; clamp to [0,1]
                    const float4 xw = ct.colw[c];
                    const int4 xo = ct.colx[c];
                    float res[2];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int r = n + half * P2_D;
                        const float4 yw = ct.roww[r];
                        const int4 yo = ct.rowy[r];
                        const int yoff[4] = {yo.x, yo.y, yo.z, yo.w};
                        const float ywv[4] = {yw.x, yw.y, yw.z, yw.w};
                        float acc = 0.0f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(W(yoff[i] + xo.x), xw.x), ywv[i]));
                            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(W(yoff[i] + xo.y), xw.y), ywv[i]));
                            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(W(yoff[i] + xo.z), xw.z), ywv[i]));
                            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(W(yoff[i] + xo.w), xw.w), ywv[i]));
                        }
                        res[half] = fminf(fmaxf(acc, 0.0f), 1.0f);
                    }
                    vl[k] = res[0]; vh[k] = res[1];
                    if (4 * q + k < P2_NU) sm.u2[n * P2_UPITCH + c] = pk(res[0], res[1]);
                    continue;
                }
                const int2 ya = sm.rowy[n], yb = sm.rowy[n + P2_D];
                const float2 wa = sm.rowv[n], wb = sm.rowv[n + P2_D];
                const p2 p00 = pk(W(ya.x + cx.x), W(yb.x + cx.x)), p01 = pk(W(ya.x + cx.y), W(yb.x + cx.y));
                const p2 p10 = pk(W(ya.y + cx.x), W(yb.y + cx.x)), p11 = pk(W(ya.y + cx.y), W(yb.y + cx.y));
                const p2 V = pk(wa.x, wb.x), OMV = pk(wa.y, wb.y);
                float al, ah, tl, th;
                upk(mul2(mul2(OMU2, OMV), p00), al, ah);
                upk(mul2(mul2(U2, OMV), p01), tl, th);
                al = __fadd_rn(al, tl); ah = __fadd_rn(ah, th);
                upk(mul2(mul2(OMU2, V), p10), tl, th);
                al = __fadd_rn(al, tl); ah = __fadd_rn(ah, th);
                upk(mul2(mul2(U2, V), p11), tl, th);
                al = __fadd_rn(al, tl); ah = __fadd_rn(ah, th);
                vl[k] = al; vh[k] = ah;
                if (4 * q + k < P2_NU) sm.u2[n * P2_UPITCH + c] = pk(al, ah);
            }
            if (col_owned) {
                const int le_lo = ty0 + 4 * q, le_hi = le_lo + P2_D;   // band-local extended rows of the two quads
                // tile rows 32..35 exist both as .hi of pairs 0..3 and as .lo of pairs 32..35: the .hi copy stores them
                const bool st_lo = q < P2_D / 4 && le_lo >= lo && le_lo < hi, st_hi = le_hi >= lo && le_hi < hi;
                if (st_lo) *reinterpret_cast<float4*>(ucol + le_lo) = make_float4(vl[0], vl[1], vl[2], vl[3]);
                if (st_hi) *reinterpret_cast<float4*>(ucol + le_hi) = make_float4(vh[0], vh[1], vh[2], vh[3]);
                if (DBG && frame == 0 && p.dbg_u && ge >= kMargin && ge < p.dw + kMargin) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int a = le_lo + k, b = le_hi + k;
                        if (st_lo && a >= kMargin && a < p.rows + kMargin) p.dbg_u[(size_t)(a - kMargin) * p.dbg_pitch + (ge - kMargin)] = vl[k];
                        if (st_hi && b >= kMargin && b < p.rows + kMargin && 4 * q + k + P2_D < PU_H) p.dbg_u[(size_t)(b - kMargin) * p.dbg_pitch + (ge - kMargin)] = vh[k];
                    }
                }
            }
        }
    }
    __syncthreads();
    }

    // ---- phase 2: Sobel, products, horizontal 9-tap Gaussian on row pairs (m, m + 32).  One work item = 8
    // consecutive output columns of one pair row (256 items = one per thread), streamed over the 16 gradient
    // columns it needs: every new gradient column adds its term to each output that uses it, which visits the
    // taps of an output in ascending order like the oracle.  Lanes of a quarter warp take 8 consecutive pair
    // rows, so the 128-bit loads are bank-conflict-free.
    {
        const int m = tid & (P2_D - 1), q = tid >> 5;
        const ulonglong2* r0 = reinterpret_cast<const ulonglong2*>(&sm.u2[(m + 0) * P2_UPITCH + 8 * q]);
        const ulonglong2* r1 = reinterpret_cast<const ulonglong2*>(&sm.u2[(m + 1) * P2_UPITCH + 8 * q]);
        const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(&sm.u2[(m + 2) * P2_UPITCH + 8 * q]);
        p2 axx[8], axy[8], ayy[8];
        p2 a0, a1, a2, b0, b1, b2, c0, c1, c2;      // columns g, g+1, g+2 of the three rows
        { ulonglong2 t = r0[0]; a0 = t.x; a1 = t.y; t = r1[0]; b0 = t.x; b1 = t.y; t = r2[0]; c0 = t.x; c1 = t.y; }
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            if ((g & 1) == 0) {                      // columns g+2, g+3 arrive together
                ulonglong2 t = r0[g / 2 + 1]; a2 = t.x; const p2 a3 = t.y;
                t = r1[g / 2 + 1]; b2 = t.x; const p2 b3 = t.y;
                t = r2[g / 2 + 1]; c2 = t.x; const p2 c3 = t.y;
                // even step: consume (0,1,2); the odd step that follows uses (1,2,3)
                {
                    const p2 d0 = sub2(a0, a2), d1 = sub2(b0, b2), d2 = sub2(c0, c2);
                    const p2 gx = add2(add2(d0, add2(d1, d1)), d2);          // 2*d1 == d1+d1 exactly
                    const p2 s0 = add2(add2(a0, add2(a1, a1)), a2);
                    const p2 s2 = add2(add2(c0, add2(c1, c1)), c2);
                    const p2 gy = sub2(s0, s2);
                    const p2 pxx = mul2(gx, gx), pxy = mul2(gx, gy), pyy = mul2(gy, gy);
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const int k = g - o;
                        if (k == 0) { axx[o] = mul2(bc(g1c(0)), pxx); axy[o] = mul2(bc(g1c(0)), pxy); ayy[o] = mul2(bc(g1c(0)), pyy); }
                        else if (k > 0 && k < 9) { axx[o] = fma2(bc(g1c(k)), pxx, axx[o]); axy[o] = fma2(bc(g1c(k)), pxy, axy[o]); ayy[o] = fma2(bc(g1c(k)), pyy, ayy[o]); }
                    }
                }
                {
                    const int g1 = g + 1;
                    const p2 d0 = sub2(a1, a3), d1 = sub2(b1, b3), d2 = sub2(c1, c3);
                    const p2 gx = add2(add2(d0, add2(d1, d1)), d2);
                    const p2 s0 = add2(add2(a1, add2(a2, a2)), a3);
                    const p2 s2 = add2(add2(c1, add2(c2, c2)), c3);
                    const p2 gy = sub2(s0, s2);
                    const p2 pxx = mul2(gx, gx), pxy = mul2(gx, gy), pyy = mul2(gy, gy);
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const int k = g1 - o;
                        if (k == 0) { axx[o] = mul2(bc(g1c(0)), pxx); axy[o] = mul2(bc(g1c(0)), pxy); ayy[o] = mul2(bc(g1c(0)), pyy); }
                        else if (k > 0 && k < 9) { axx[o] = fma2(bc(g1c(k)), pxx, axx[o]); axy[o] = fma2(bc(g1c(k)), pxy, axy[o]); ayy[o] = fma2(bc(g1c(k)), pyy, ayy[o]); }
                    }
                }
                a0 = a2; a1 = a3; b0 = b2; b1 = b3; c0 = c2; c1 = c3;
            }
        }
        // the H planes keep the scalar row-major layout (phase 3 pairs columns, not rows)
        float xl[8], xh[8], yl[8], yh[8], zl[8], zh[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) { upk(axx[o], xl[o], xh[o]); upk(axy[o], yl[o], yh[o]); upk(ayy[o], zl[o], zh[o]); }
        float4* d;
        d = reinterpret_cast<float4*>(&sm.h[0][m * PH_PITCH + 8 * q]);           d[0] = make_float4(xl[0], xl[1], xl[2], xl[3]); d[1] = make_float4(xl[4], xl[5], xl[6], xl[7]);
        d = reinterpret_cast<float4*>(&sm.h[0][(m + P2_D) * PH_PITCH + 8 * q]);  d[0] = make_float4(xh[0], xh[1], xh[2], xh[3]); d[1] = make_float4(xh[4], xh[5], xh[6], xh[7]);
        d = reinterpret_cast<float4*>(&sm.h[1][m * PH_PITCH + 8 * q]);           d[0] = make_float4(yl[0], yl[1], yl[2], yl[3]); d[1] = make_float4(yl[4], yl[5], yl[6], yl[7]);
        d = reinterpret_cast<float4*>(&sm.h[1][(m + P2_D) * PH_PITCH + 8 * q]);  d[0] = make_float4(yh[0], yh[1], yh[2], yh[3]); d[1] = make_float4(yh[4], yh[5], yh[6], yh[7]);
        d = reinterpret_cast<float4*>(&sm.h[2][m * PH_PITCH + 8 * q]);           d[0] = make_float4(zl[0], zl[1], zl[2], zl[3]); d[1] = make_float4(zl[4], zl[5], zl[6], zl[7]);
        d = reinterpret_cast<float4*>(&sm.h[2][(m + P2_D) * PH_PITCH + 8 * q]);  d[0] = make_float4(zh[0], zh[1], zh[2], zh[3]); d[1] = make_float4(zh[4], zh[5], zh[6], zh[7]);
    }
    __syncthreads();

    // ---- phase 3a: vertical 9-tap Gaussian on column pairs.  Thread = two adjacent columns (one 64-bit load
    // per row), 7 consecutive rows: per plane its 15 input pairs are pulled into registers, then (after a
    // barrier) the 7 results overwrite the first rows of its own range in place, so phase 3b can run as a
    // rolled loop.
    const int xo = 2 * (tid & 31), grp = tid >> 5;
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
        float* plane = &sm.h[ch][(grp * P2_RPT) * PH_PITCH + xo];
        p2 in[P2_RPT + 8];
#pragma unroll
        for (int k = 0; k < P2_RPT + 8; ++k) in[k] = *reinterpret_cast<const p2*>(plane + k * PH_PITCH);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < P2_RPT; ++j) {
            p2 mm = mul2(bc(g1c(0)), in[j]);
#pragma unroll
            for (int k = 1; k < 9; ++k) mm = fma2(bc(g1c(k)), in[j + k], mm);
            *reinterpret_cast<p2*>(plane + j * PH_PITCH) = mm;
        }
    }
    // no barrier needed: phase 3b reads back only what this thread wrote

    // ---- phase 3b: 2x2 eigen-solve, quantise, hash (raisr.cl:278-317) for the two columns of the thread
    {
        const int x = tx0 + xo;                 // .lo column; .hi is x + 1
        const float PI_F = 3.14159265358979323846f;
        float sq[NQ], cq[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) { sq[i] = p.sq[i]; cq[i] = p.cq[i]; }
        const int xsl = x / S, xtl = x % S, xsh = (x + 1) / S, xth = (x + 1) % S;
        const bool okl = x < p.dw, okh = x + 1 < p.dw;
        // per-thread store cursor: pixel row-type alternates with the row, own row advances every S rows
        const int yl0 = ty0 + grp * P2_RPT;
        int yt = yl0 % S;                  // y0 is a multiple of S, so yl % S == global y % S
        const size_t row_step = (size_t)S * p.hash_plane_stride;                       // next row: next row-type
        const size_t wrap_step = p.hash_pitch - (size_t)(S - 1) * row_step;            // after S rows: first row-type, next own row
        uint8_t* hrow = p.hash + (size_t)frame * p.hash_frame_stride + (size_t)(yt * S) * p.hash_plane_stride + (size_t)(yl0 / S) * p.hash_pitch;
        const size_t offl = (size_t)xtl * p.hash_plane_stride + xsl, offh = (size_t)xth * p.hash_plane_stride + xsh;
        const float* hp = &sm.h[0][(grp * P2_RPT) * PH_PITCH + xo];
#pragma unroll 1
        for (int j = 0; j < P2_RPT; ++j) {
            const p2 mb = *reinterpret_cast<const p2*>(hp + PH_H * PH_PITCH);
            const p2 ma = p.as_written ? mb : *reinterpret_cast<const p2*>(hp);            // raisr.cl:271 accumulates gx*gy into ma
            const p2 md = *reinterpret_cast<const p2*>(hp + 2 * PH_H * PH_PITCH);
            hp += PH_PITCH;
            if (!DBG && p.tens) {    // the filter kernel solves the eigen problem: hand it the tensor (uniform branch)
                if (yl0 + j < p.rows) {
                    float* t0 = p.tens + (hrow - p.hash);
                    float al, ah, bl, bh, dl, dh;
                    upk(ma, al, ah); upk(mb, bl, bh); upk(md, dl, dh);
                    if (okl) { t0[offl] = al; t0[offl + p.tens_plane_stride] = bl; t0[offl + 2 * p.tens_plane_stride] = dl; }
                    if (okh) { t0[offh] = ah; t0[offh + p.tens_plane_stride] = bh; t0[offh + 2 * p.tens_plane_stride] = dh; }
                }
                if (++yt == S) { yt = 0; hrow += wrap_step; } else hrow += row_step;
                continue;
            }
            const p2 T = add2(ma, md);
            float dal, dah, dbl, dbh;                       // scalar subtraction: see the note on contraction in phase 1
            upk(mul2(ma, md), dal, dah);
            upk(mul2(mb, mb), dbl, dbh);
            const p2 D = pk(__fsub_rn(dal, dbl), __fsub_rn(dah, dbh));
            const p2 rad = sub2(mul2(mul2(T, T), bc(0.25f)), D);     // (a fused T*T*0.25 - D is the same value: *0.25 is exact)
            float radl, radh;
            upk(rad, radl, radh);
            const p2 sqr = sqrt2_guarded(radl, radh);   // radicand clamped at 0 inside
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
                const p2 s1 = sqrt2_guarded(L1l, L1h), s2 = sqrt2_guarded(L2l, L2h);   // L2 clamped at 0 inside
                const p2 den = add2(s1, s2), num = sub2(s1, s2);
                float dl, dh, nl, nh;
                upk(den, dl, dh);
                upk(num, nl, nh);
                // den >= 0.  The refinement runs on den itself with rcp(max(den, 1e-30)): the scalar fast path for
                // den in [1e-15, 1e15] and exactly 0 for den == 0 (num is 0 then), the reference's "coherence = 0";
                // anything else (never seen in practice) takes the IEEE divide.
                float rl, rh;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(fmaxf(dl, 1.0e-30f)));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(fmaxf(dh, 1.0e-30f)));
                p2 r = pk(rl, rh);
                r = fma2(r, fnma2(den, r, bc(1.0f)), r);
                const p2 qq = mul2(num, r);
                upk(fma2(r, fnma2(den, qq, num), qq), col, coh);
                const unsigned lo = __float_as_uint(1.0e-15f) - 1u;   // (bits - 1) < lo  <=>  0 < den < 1e-15
                const bool slow_l = ((__float_as_uint(dl) - 1u) < lo) | !(dl <= 1.0e15f), slow_h = ((__float_as_uint(dh) - 1u) < lo) | !(dh <= 1.0e15f);
                if (slow_l | slow_h) {
                    if (slow_l) col = __fdiv_rn(nl, dl);
                    if (slow_h) coh = __fdiv_rn(nh, dh);
                }
            }
            // theta / pi with the divide's own fast path (theta is 0 or in [1e-8, pi]; exact for theta = 0)
            float tql, tqh;
            {
                const float PI_INV = 0.31830988618379067154f;
                const float r = __fmaf_rn(PI_INV, __fmaf_rn(-PI_F, PI_INV, 1.0f), PI_INV);
                const p2 th = pk(thl, thh);
                const p2 q = mul2(th, bc(r));
                upk(fma2(bc(r), fma2(bc(-PI_F), q, th), q), tql, tqh);
                const unsigned lo = __float_as_uint(1.0e-15f) - 1u;   // theta >= 0: (bits - 1) < lo  <=>  0 < theta < 1e-15
                const bool slow_l = (__float_as_uint(thl) - 1u) < lo, slow_h = (__float_as_uint(thh) - 1u) < lo;
                if (slow_l | slow_h) {
                    if (slow_l) tql = __fdiv_rn(thl, PI_F);
                    if (slow_h) tqh = __fdiv_rn(thh, PI_F);
                }
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
            const int yl = yl0 + j;  // band-local output row
            if (yl < p.rows) {
                if (okl) hrow[offl] = (uint8_t)bl;
                if (okh) hrow[offh] = (uint8_t)bh;
                if (DBG && frame == 0) {
                    const size_t o = (size_t)yl * p.dbg_pitch + x;
                    if (okl) {
                        if (p.dbg_hash) p.dbg_hash[o] = bl * (S * S) + yt * S + xtl;
                        if (p.dbg_angle) p.dbg_angle[o] = thl;
                        if (p.dbg_l1) p.dbg_l1[o] = L1l;
                        if (p.dbg_coh) p.dbg_coh[o] = col;
                    }
                    if (okh) {
                        if (p.dbg_hash) p.dbg_hash[o + 1] = bh * (S * S) + yt * S + xth;
                        if (p.dbg_angle) p.dbg_angle[o + 1] = thh;
                        if (p.dbg_l1) p.dbg_l1[o + 1] = L1h;
                        if (p.dbg_coh) p.dbg_coh[o + 1] = coh;
                    }
                }
            }
            if (++yt == S) { yt = 0; hrow += wrap_step; } else hrow += row_step;
        }
    }
    __syncthreads();   // shared memory is reused by the next tile
    }
}

}  // namespace raisr
