// LSU / MIO instruction throughput on B200 with high ILP: R independent loads (or shuffles) are
// issued back to back, then consumed.  Reports cycles per warp-level instruction per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// MODE 0: LDS.32 conflict-free (lane-consecutive)  1: LDS.32 broadcast  2: LDS.64 consecutive
//      3: LDS.128 consecutive  4: SHFL.BFLY  5: LDS.32 + FFMA x4 mix  6: LDS.128 octet pattern (each quarter reads own 128B row)
//      7: LDS.32 2-way conflict  8: mix: 3x LDS.128 octet + 4x LDS.32 + 16 FFMA (the filter inner loop shape)
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, int iters)
{
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    unsigned a32 = base + 4 * lane + 128 * (warp & 7);
    unsigned abc = base + 128 * (warp & 7);
    unsigned a64 = base + 8 * lane + 256 * (warp & 3);
    unsigned a128 = base + 16 * lane + 512 * (warp & 3);
    unsigned aoct = base + 16 * (lane & 7) + 512 * ((lane >> 3) + 4 * (warp & 3));   // each quarter: one 128 B row
    unsigned a2w = base + 8 * lane;   // stride 2 words -> 2-way conflict
    float acc = 0, f0 = lane, f1 = 1, f2 = 2, f3 = 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float r[32];
        const unsigned tog = (it & 1) << 15;   // keep the compiler from hoisting the loads
        a32 ^= tog; abc ^= tog; a64 ^= tog; a128 ^= tog; aoct ^= tog; a2w ^= tog;
        if (MODE == 0 || MODE == 1 || MODE == 7) {
            unsigned a = MODE == 0 ? a32 : MODE == 1 ? abc : a2w;
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[i]) : "r"(a + 1024 * i) : "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += r[i];
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r[2 * i]), "=f"(r[2 * i + 1]) : "r"(a64 + 1024 * i) : "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += r[i];
        } else if (MODE == 3 || MODE == 6) {
            unsigned a = MODE == 3 ? a128 : aoct;
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[4 * i]), "=f"(r[4 * i + 1]), "=f"(r[4 * i + 2]), "=f"(r[4 * i + 3]) : "r"(a + 2048 * i) : "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __shfl_xor_sync(0xffffffffu, f0 + i, 1 + (i & 3));
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += r[i];
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[i]) : "r"(a32 + 1024 * i) : "memory");
                f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += r[i];
        } else if (MODE == 8) {
#pragma unroll
            for (int px = 0; px < 2; ++px) {
#pragma unroll
                for (int i = 0; i < 3; ++i) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[4 * i]), "=f"(r[4 * i + 1]), "=f"(r[4 * i + 2]), "=f"(r[4 * i + 3]) : "r"(aoct + 2048 * i) : "memory");
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[12 + i]) : "r"(a32 + 1024 * i) : "memory");
#pragma unroll
                for (int i = 0; i < 16; i += 4) { f0 = fmaf(f0, r[i], f1); f1 = fmaf(f1, r[i + 1], f2); f2 = fmaf(f2, r[i + 2], f3); f3 = fmaf(f3, r[i + 3], f0); }
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + f0 + f1 + f2 + f3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int sms, float* out, long long* cyc)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int nt : {512, 1024}) {
        int iters = 20000;
        k<MODE><<<sms, nt, 64 * 1024>>>(out, cyc, 100);
        CK(cudaEventRecord(e0));
        k<MODE><<<sms, nt, 64 * 1024>>>(out, cyc, iters);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double m = 0; for (auto v : h) m += v; m /= sms;
        double n_instr = (double)iters * per_iter * (nt / 32);
        printf("{\"bench\": \"%s\", \"threads\": %d, \"warp0_cycles_per_instr\": %.3f, \"event_ns_per_instr_per_sm\": %.4f, \"implied_cycles_at_1965MHz\": %.3f}\n",
               name, nt, m / n_instr, ms * 1e6 / n_instr, ms * 1e6 / n_instr * 1.965);
    }
}

int main()
{
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, sizeof(float) * 1024 * sms));
    long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    run<0>("lds32_conflict_free", 16, sms, out, cyc);
    run<1>("lds32_broadcast", 16, sms, out, cyc);
    run<7>("lds32_2way", 16, sms, out, cyc);
    run<2>("lds64_consecutive", 8, sms, out, cyc);
    run<3>("lds128_consecutive", 8, sms, out, cyc);
    run<6>("lds128_octet_rows", 8, sms, out, cyc);
    run<4>("shfl_bfly", 16, sms, out, cyc);
    run<5>("lds32_plus_4ffma", 16, sms, out, cyc);
    run<8>("filter_shape_3lds128_4lds32_16ffma(per 7 lsu instr)", 14, sms, out, cyc);
    return 0;
}
