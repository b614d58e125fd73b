// packed_f32.cuh -- two IEEE fp32 operations per instruction on sm_100a (PTX *.rn.f32x2 -> SASS FFMA2 / FADD2 /
// FMUL2): same FMA-pipe throughput as the scalar forms at half the issue slots (tools/ffma2_test.cu).
// NOTE: ptxas contracts a single-use mul.rn.f32x2 that feeds add/sub.rn.f32x2 into one FFMA2 (one rounding
// instead of two, regardless of -fmad); where the two roundings matter, do the addition on the unpacked halves.
#pragma once
#include <cuda_runtime.h>

namespace raisr {

typedef unsigned long long p2;   // two packed fp32 in a 64-bit register pair (.lo, .hi)

__device__ __forceinline__ p2 pk(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(p2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ p2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { p2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 sub2(p2 a, p2 b) { p2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 mul2(p2 a, p2 b) { p2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// c - a*b, one rounding (the -a*b + c step of the sqrt / divide refinements)
__device__ __forceinline__ p2 fnma2(p2 a, p2 b, p2 c) { return fma2(mul2(a, bc(-1.0f)), b, c); }

}  // namespace raisr
