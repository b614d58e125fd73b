// Pipe microbenchmarks used to size the RAISR filter kernel on B200 (sm_100a).
//   ffma      register-only FFMA chains -> measured FP32 peak (the roofline denominator candidate)
//   lds       per-lane gathers of 16/8/4-byte chunks from a [n_filters][stride] shared-memory table
//             under several index patterns -> cycles per warp-level LDS at saturation
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    float b = 1.0001f, c = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W> // bytes per lane: 16, 8, 4
__global__ void __launch_bounds__(512) k_lds(const int* __restrict__ idx, float* out, long long* cyc,
                                            int iters, int stride_f, int n_filters, int ffma_per_ld)
{
    extern __shared__ float tab[];
    for (int i = threadIdx.x; i < n_filters * stride_f; i += blockDim.x) tab[i] = (i % 97) * 1e-3f;
    __syncthreads();
    int h = idx[threadIdx.x & 31];
    unsigned base = (unsigned)__cvta_generic_to_shared(tab + (size_t)h * stride_f);
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    float e0 = threadIdx.x, e1 = 1, e2 = 2, e3 = 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 30; ++c) {
            if (W == 16) {
                float x, y, z, w;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(base + c * 16));
                acc0 += x; acc1 += y; acc2 += z; acc3 += w;
            } else if (W == 8) {
                float x, y;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(base + c * 8));
                acc0 += x; acc1 += y;
            } else {
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(base + c * 4));
                acc0 += x;
            }
            for (int f = 0; f < ffma_per_ld; ++f) { e0 = fmaf(e0, 1.0001f, acc0); e1 = fmaf(e1, 1.0001f, acc1); e2 = fmaf(e2, 1.0001f, acc2); e3 = fmaf(e3, 1.0001f, acc3);} 
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3 + e0 + e1 + e2 + e3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main(int argc, char** argv)
{
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int sms = pr.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d}\n", pr.name, sms, clk_khz);
    float* out; CK(cudaMalloc(&out, sizeof(float) * 512 * sms * 16));
    long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // ---- FFMA peak
    {
        int iters = 20000, blocks = sms * 8;
        k_ffma<<<blocks, 256>>>(out, 100); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0)); k_ffma<<<blocks, 256>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        double flops = 2.0 * 128 * iters * (double)blocks * 256;
        printf("{\"bench\": \"ffma\", \"ms\": %.3f, \"tflops\": %.2f, \"implied_mhz_at_128_per_sm\": %.0f}\n", best, flops / best / 1e9,
               flops / 2 / 128 / sms / (best * 1e-3) / 1e6);
    }
    // ---- LDS gathers
    const int NF = 216;
    CK(cudaFuncSetAttribute(k_lds<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_lds<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int* didx; CK(cudaMalloc(&didx, 32 * sizeof(int)));
    struct Pat { const char* name; int id; };
    Pat pats[] = {{"same", 0}, {"consecutive", 1}, {"all_same_group", 2}, {"quarter_same_group", 3}, {"random", 4},
                  {"random_pairs_shared", 5}, {"random_6of8", 6}, {"stride2_consecutive", 7}};
    int strides[] = {132, 124, 128, 136};
    for (int width : {16, 8, 4}) for (int stride : strides) for (auto& pt : pats) for (int nthreads : {256, 512}) for (int ff : {0, 1}) {
        if (ff == 1 && !(width == 16 && stride == 132 && nthreads == 256)) continue;
        if (nthreads == 512 && !(stride == 132)) continue;
        if (width != 16 && stride != 132) continue;
        double avg = 0; int trials = (pt.id >= 4 && pt.id <= 6) ? 8 : 1;
        for (int tr = 0; tr < trials; ++tr) {
            int h[32]; unsigned s = 12345u + tr * 777u;
            int chunkstride = stride / 4;  // in 16B chunks; group = (h*chunkstride) mod 8
            for (int l = 0; l < 32; ++l) {
                switch (pt.id) {
                case 0: h[l] = 7; break;
                case 1: h[l] = l; break;
                case 2: h[l] = (8 * l) % NF; break;   // with odd chunk stride: same bank group, distinct rows
                case 3: h[l] = (l / 8) + 8 * (l % 8); break;
                case 4: h[l] = lcg(s) % NF; break;
                case 5: if (l % 2 == 0) h[l] = lcg(s) % NF; else h[l] = h[l - 1]; break;
                case 6: { int q = l / 8; static int pool[6]; if (l % 8 == 0) for (int i = 0; i < 6; ++i) pool[i] = lcg(s) % NF; (void)q; h[l] = pool[lcg(s) % 6]; } break;
                case 7: h[l] = 2 * l; break;
                }
            }
            (void)chunkstride;
            CK(cudaMemcpy(didx, h, sizeof(h), cudaMemcpyHostToDevice));
            int iters = 2000; size_t smem = (size_t)NF * stride * 4;
            std::vector<long long> hc(sms);
            for (int rep = 0; rep < 2; ++rep) {
                if (width == 16) k_lds<16><<<sms, nthreads, smem>>>(didx, out, cyc, iters, stride, NF, ff * 4);
                else if (width == 8) k_lds<8><<<sms, nthreads, smem>>>(didx, out, cyc, iters, stride, NF, 0);
                else k_lds<4><<<sms, nthreads, smem>>>(didx, out, cyc, iters, stride, NF, 0);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(hc.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            double m = 0; for (int i = 0; i < sms; ++i) m += hc[i]; m /= sms;
            avg += m / ((double)iters * 30 * (nthreads / 32));
        }
        avg /= trials;
        printf("{\"bench\": \"lds\", \"bytes_per_lane\": %d, \"stride_floats\": %d, \"pattern\": \"%s\", \"threads\": %d, \"ffma_per_ld\": %d, \"cycles_per_warp_ld\": %.2f}\n",
               width, stride, pt.name, nthreads, ff * 16, avg);
    }
    return 0;
}
