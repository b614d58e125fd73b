// Does the texture path add load bandwidth on top of the shared-memory (LSU) data pipe on sm_100a?
// The filter kernel is bound by the LSU data pipe (1 wavefront = 128 B / clk / SM).  If TEX fetches that hit
// L1 are served by a separate data stage, the per-pixel patch stream could move there.
//   ./tex_test  -> JSON lines: bytes/clk/SM of (a) LDS.128 only, (b) tex1Dfetch<float4> only, (c) both interleaved
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: LDS only, 1: TEX only, 2: 2 LDS : 1 TEX, 3: LDS + __ldg (LSU global path, L1 hit)
__global__ void __launch_bounds__(256) k(cudaTextureObject_t tex, const float4* __restrict__ g, float* out, int iters)
{
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int si = (w * 64 + lane) & 2047, ti = blockIdx.x * 1024 + ((w * 32 + lane) & 1023);   // 16 KB of texels per CTA: L1-resident
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // independent loads: several in flight per thread
            if (MODE == 0 || MODE == 2 || MODE == 3) {
                float4 a = sm[si], b = sm[(si + 32) & 2047];
                acc.x += a.x + b.x; acc.y += a.y + b.y; acc.z += a.z + b.z; acc.w += a.w + b.w;
                si = (si + 64) & 2047;                  // addresses change every step (nothing to hoist) but do not depend on data
            }
            if (MODE == 1 || MODE == 2) {
                float4 t = tex1Dfetch<float4>(tex, ti);
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                ti = blockIdx.x * 1024 + ((ti + 32) & 1023);
            }
            if (MODE == 3) {
                float4 t = __ldg(g + ti);
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                ti = blockIdx.x * 1024 + ((ti + 32) & 1023);
            }
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>
void run(const char* name, cudaTextureObject_t tex, const float4* g, float* out, int sms, double ghz)
{
    const int blocks = sms * 4, iters = 4000;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    k<MODE><<<blocks, 256, 32768>>>(tex, g, out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256, 32768>>>(tex, g, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double clk = best * 1e-3 * ghz * 1e9;
    const double per_sm_threads = 4.0 * 256, its = (double)iters * 4;
    const double lds_b = (MODE == 0 || MODE == 2 || MODE == 3) ? per_sm_threads * its * 32 : 0;
    const double tex_b = (MODE == 1 || MODE == 2 || MODE == 3) ? per_sm_threads * its * 16 : 0;
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"lds_bytes_per_clk_per_sm\": %.1f, \"tex_or_ldg_bytes_per_clk_per_sm\": %.1f, \"total\": %.1f}\n", name, best,
           lds_b / clk, tex_b / clk, (lds_b + tex_b) / clk);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int n = p.multiProcessorCount * 4 * 1024;
    float4* g; cudaMalloc(&g, n * sizeof(float4));
    float4* h = new float4[n];
    for (int i = 0; i < n; ++i) h[i] = make_float4(i & 7, 1, 2, 3);
    cudaMemcpy(g, h, n * sizeof(float4), cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = g;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = n * sizeof(float4);
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    float* out; cudaMalloc(&out, p.multiProcessorCount * 4 * 256 * sizeof(float));
    const double ghz = khz / 1e6;
    run<0>("LDS.128 only", tex, g, out, p.multiProcessorCount, ghz);
    run<1>("tex1Dfetch<float4> only (L1 hits)", tex, g, out, p.multiProcessorCount, ghz);
    run<2>("2 LDS.128 : 1 tex fetch", tex, g, out, p.multiProcessorCount, ghz);
    run<3>("2 LDS.128 : 1 __ldg float4 (L1 hits)", tex, g, out, p.multiProcessorCount, ghz);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
