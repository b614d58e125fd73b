// Would fetching half of the filter kernel's tap stream through the TEXTURE path relieve the shared-memory
// (LSU) data pipe?  Emulates filter_octet_kernel's per-pixel traffic: every pixel of an octet needs a 512-byte
// record picked by a random bucket (4 x 16 B per lane), four 4-byte patch loads from a shared tile, 16 FMAs.
//   MODE 0: all four chunks from a 110 KB shared-memory table (what the kernel does today)
//   MODE 1: chunks 0,1 from a 55 KB shared half-table, chunks 2,3 via tex1Dfetch<float4> from a 55 KB half-table
//           in global memory that L1 can keep (shared carve-out shrinks accordingly)
//   MODE 2: all four chunks via texture (table entirely in L1/L2)
// Prints clocks per pixel per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int NB = 216, NT = 640;

template <int MODE, int DEPTH>
__global__ void __launch_bounds__(NT, 1) k(cudaTextureObject_t tex, const float4* __restrict__ gtab, float* out, int steps, int tile_floats)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float4* tab = reinterpret_cast<float4*>(smem);                       // MODE 0: NB*32 float4; MODE 1: NB*16 float4
    constexpr int TABV = MODE == 0 ? NB * 32 : (MODE == 1 ? NB * 16 : 0);
    float* tile = reinterpret_cast<float*>(smem + (size_t)TABV * 16);
    for (int i = threadIdx.x; i < TABV; i += NT) tab[i] = MODE == 0 ? gtab[i] : gtab[(i / 16) * 32 + (i % 16)];
    for (int i = threadIdx.x; i < tile_floats; i += NT) tile[i] = 1.0f + i * 1e-6f;
    __syncthreads();
    const int lane8 = threadIdx.x & 7, octet = threadIdx.x >> 3;
    unsigned rng = 12345u + 977u * (blockIdx.x * 80 + octet);
    auto next_bucket = [&]() { rng = rng * 1664525u + 1013904223u; return ((rng >> 16) * NB) >> 16; };
    float acc = 0.0f;
    unsigned b_next = next_bucket();
    float4 c[4], n[4];
    auto fetch = [&](float4 (&t)[4], unsigned b) {
        if (MODE == 0) { const float4* r = tab + b * 32 + lane8; t[0] = r[0]; t[1] = r[8]; t[2] = r[16]; t[3] = r[24]; }
        else if (MODE == 1) {
            const float4* r = tab + b * 16 + lane8; t[0] = r[0]; t[1] = r[8];
            t[2] = tex1Dfetch<float4>(tex, b * 32 + 16 + lane8); t[3] = tex1Dfetch<float4>(tex, b * 32 + 24 + lane8);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) t[q] = tex1Dfetch<float4>(tex, b * 32 + 8 * q + lane8);
        }
    };
    fetch(c, b_next);
    int pos = (octet * 8 + lane8) % (tile_floats - 64);   // a warp's 32 lanes read 32 consecutive words: conflict-free
    auto pixel = [&](float4 (&cur)[4], float4 (&nxt)[4]) {
        fetch(nxt, next_bucket());                     // taps of the next pixel, in flight during this pixel's FMAs
        const float u0 = tile[pos], u1 = tile[pos + 13], u2 = tile[pos + 29], u3 = tile[pos + 47];   // patch refresh
        pos += 64; if (pos >= tile_floats - 64) pos -= tile_floats - 64;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc = fmaf(u0, cur[q].x, acc); acc = fmaf(u1, cur[q].y, acc); acc = fmaf(u2, cur[q].z, acc); acc = fmaf(u3, cur[q].w, acc);
        }
    };
    if (DEPTH == 1) {
        for (int s = 0; s < steps; s += 2) {           // ping-pong: no register copies
            pixel(c, n);
            pixel(n, c);
        }
    } else {                                           // two pixels of taps in flight: ring of three register sets
        float4 m[4];
        fetch(n, next_bucket());
        for (int s = 0; s < steps; s += 3) {
            pixel(c, m);
            pixel(n, c);
            pixel(m, n);
        }
    }
    out[blockIdx.x * NT + threadIdx.x] = acc;
}

template <int MODE, int DEPTH>
void run(const char* name, cudaTextureObject_t tex, const float4* g, float* out, int sms, double ghz)
{
    const int tile_floats = 25 * 1024;                                    // 100 KB of tile buffers, as in the kernel
    const size_t smem = (MODE == 0 ? NB * 512 : (MODE == 1 ? NB * 256 : 0)) + (size_t)tile_floats * 4;
    cudaFuncSetAttribute(k<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int steps = 20000;
    k<MODE, DEPTH><<<sms, NT, smem>>>(tex, g, out, 100, tile_floats);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE, DEPTH><<<sms, NT, smem>>>(tex, g, out, steps, tile_floats);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double clk = best * 1e-3 * ghz * 1e9, px = (double)steps * (NT / 8);
    printf("{\"variant\": \"%s\", \"smem_kb\": %.0f, \"ms\": %.3f, \"clk_per_px_per_sm\": %.2f, \"err\": \"%s\"}\n", name, smem / 1024.0, best, clk / px,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int n = NB * 32;
    float4* g; cudaMalloc(&g, n * sizeof(float4));
    float4* h = new float4[n];
    for (int i = 0; i < n; ++i) h[i] = make_float4(1e-3f * (i % 7), 2e-3f, 3e-3f, 1e-3f);
    cudaMemcpy(g, h, n * sizeof(float4), cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = g;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = n * sizeof(float4);
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    float* out; cudaMalloc(&out, p.multiProcessorCount * NT * sizeof(float));
    const double ghz = khz / 1e6;
    run<0, 1>("4 chunks LDS (110 KB smem table), 1 pixel ahead", tex, g, out, p.multiProcessorCount, ghz);
    run<0, 2>("4 chunks LDS, 2 pixels ahead", tex, g, out, p.multiProcessorCount, ghz);
    run<1, 1>("2 chunks LDS (55 KB) + 2 chunks TEX, 1 pixel ahead", tex, g, out, p.multiProcessorCount, ghz);
    run<1, 2>("2 chunks LDS + 2 chunks TEX, 2 pixels ahead", tex, g, out, p.multiProcessorCount, ghz);
    run<2, 2>("4 chunks TEX, 2 pixels ahead", tex, g, out, p.multiProcessorCount, ghz);
    return 0;
}
