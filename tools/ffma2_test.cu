// Packed fp32 (fma.rn.f32x2 -> FFMA2) versus scalar FFMA on sm_100a: throughput of the FMA pipe and
// what the packed form leaves of the issue slots.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   ./ffma2_test    prints JSON lines {variant, tflops, clk_per_warp_instr}
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    unsigned long long d, a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

template <int MODE>   // 0: 16 independent FFMA chains; 1: 8 FFMA2 chains (same flops); 2: FFMA2 + as many integer ops; 3: FFMA + as many integer ops
__global__ void __launch_bounds__(256) k(float* out, int iters, float s)
{
    float a[16];
    int z[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = threadIdx.x + i;
    const float m0 = s, m1 = s * 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0 || MODE == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m0), "f"(m1));
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) ffma2(a[i], a[i + 1], a[i], a[i + 1], m0, m0);
            }
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(it));
            }
            if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(it));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(r));
                }
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += (float)z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int sms, float ghz)
{
    const int blocks = sms * 8, iters = 2000;
    float* out;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 10, 1.0001f);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, iters, 1.0001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fma = (double)blocks * 256 * iters * 4 * 16;
    const double fp_instr_per_warp = (MODE == 0 || MODE == 3) ? 64.0 : 32.0;
    const double int_instr_per_warp = MODE == 2 ? 32.0 : (MODE == 3 ? 64.0 : 0.0);
    const double warps_per_sm = 8.0 * 8;  // 8 blocks x 8 warps resident per SM... each SM runs blocks/sms blocks
    const double clk = best * 1e-3 * ghz * 1e9;
    const double winstr_per_sm = (double)blocks / sms * 8 * iters * (fp_instr_per_warp + int_instr_per_warp);
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"tflops\": %.1f, \"warp_instr_per_clk_per_sm\": %.2f}\n", name, best, 2 * fma / best / 1e9,
           winstr_per_sm / clk);
    (void)warps_per_sm;
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float ghz = khz / 1e6f;
    run<0>("ffma x16", p.multiProcessorCount, ghz);
    run<1>("ffma2 x8 (same flops)", p.multiProcessorCount, ghz);
    run<2>("ffma2 x8 + 8 int ops per 8 ffma2", p.multiProcessorCount, ghz);
    run<3>("ffma x16 + 16 int ops", p.multiProcessorCount, ghz);
    return 0;
}
