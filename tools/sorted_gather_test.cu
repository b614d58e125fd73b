// Candidate (d) of SURVEY.md 7.2-1, measured: "CTA-local counting sort by bucket so that a warp shares a filter".
// After such a sort an octet's consecutive pixels have the same bucket -- its 16 taps per lane stay in registers, the
// tap stream vanishes -- but they are no longer adjacent in the image, so the sliding register window is gone and every
// pixel has to fetch its whole patch: 16 shared loads per lane and pixel at an arbitrary tile position.
//   MODE 0  sorted walk, BEST case: taps never reloaded (and the sort itself is free); 16 patch loads per lane and pixel
//           from a random position of the column-major tile (137 x 92 floats, the production tile)
//   MODE 1  the production walk for comparison: adjacent pixels, sliding window (2 + 2 fresh patch loads per lane and
//           pixel), three 16-byte tap loads per lane and pixel from a 216 x 384 B table at a random bucket, skipped
//           with probability 0.27 (the measured bucket-run statistics)
// Both do the same 16 FMAs per lane and pixel and a 7-shuffle butterfly per 8 pixels.  Prints clocks per pixel per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/sorted_gather_test tools/sorted_gather_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int NB = 216, NT = 640, PT = 92, NCOLS = 137, REC = 384;

template <int MODE>
__global__ void __launch_bounds__(NT, 1) k(const uint4* __restrict__ gtab, float* out, int steps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint4* tab = reinterpret_cast<uint4*>(smem);                         // NB * 24 chunks
    float* tile = reinterpret_cast<float*>(smem + (size_t)NB * REC);
    for (int i = threadIdx.x; i < NB * (REC / 16); i += NT) tab[i] = gtab[i];
    for (int i = threadIdx.x; i < NCOLS * PT; i += NT) tile[i] = 1.0f + i * 1e-6f;
    __syncthreads();
    const int lane8 = threadIdx.x & 7, octet = threadIdx.x >> 3;
    const unsigned omask = 0xffu << (threadIdx.x & 24);
    unsigned rng = 12345u + 977u * (blockIdx.x * 80 + octet);            // octet-uniform generator
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.01f * (i + lane8);
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = 0.5f;
    uint4 q0 = tab[lane8], q1 = tab[8 + lane8], q2 = tab[16 + lane8];
    float total = 0.0f;
    int col = 0;
    const int row0 = 2 * (octet % 40);                                   // a warp's octets sit on consecutive own rows
    for (int s = 0; s < steps; s += 8) {
        float acc[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (MODE == 0) {
                // a pixel somewhere in the tile: patch origin (r, cc); lane = filter row: 11 columns + a 5-run
                const unsigned r32 = rnd();
                const int r = (r32 & 0xffff) % (PT - 11), cc = (r32 >> 16) % (NCOLS - 11);
                const float* pf = tile + cc * PT + r + lane8;
#pragma unroll
                for (int j = 0; j < 11; ++j) w[j] = pf[j * PT];
                const float* pp = tile + (cc + 1 + 5 * (lane8 & 1)) * PT + r + 8 + (lane8 >> 1) % 3;
#pragma unroll
                for (int j = 0; j < 5; ++j) w[11 + j] = pp[j * PT];
            } else {
                const float* pf = tile + col * PT + row0 + lane8;
                w[(2 * b + 9) & 15] = pf[9 * PT]; w[(2 * b + 10) & 15] = pf[10 * PT];
                const float* pp = tile + (col + 1 + 5 * (lane8 & 1)) * PT + row0 + 8 + (lane8 >> 1) % 3;
                w[(2 * b + 3) & 7 | 8] = pp[3 * PT]; w[(2 * b + 4) & 7 | 8] = pp[4 * PT];
                col += 2; if (col > NCOLS - 16) col = 0;
                const unsigned r32 = rnd();
                const bool reload = (r32 & 0xffff) > 17694;             // 27 % of the pixels keep their bucket
                const unsigned rec = ((r32 >> 16) * NB) >> 8;
                if (reload) {
                    const uint4* t = tab + rec % NB * (REC / 16) + lane8;
                    q0 = t[0]; q1 = t[8]; q2 = t[16];
                }
                const unsigned x[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i] = __uint_as_float(__byte_perm(x[(3 * i) / 4], x[((3 * i) / 4 + 1) % 12], 0x3210 + 0x1111 * ((3 * i) % 4)));
            }
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; i += 2) { a0 = fmaf(w[i], c[i], a0); a1 = fmaf(w[i + 1], c[i + 1], a1); }
            acc[b] = a0 + a1;
        }
        float r4[4], r2[2];
        const bool h2 = lane8 & 4, h1 = lane8 & 2, h0 = lane8 & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) r4[i] = (h2 ? acc[i + 4] : acc[i]) + __shfl_xor_sync(omask, h2 ? acc[i] : acc[i + 4], 4);
#pragma unroll
        for (int i = 0; i < 2; ++i) r2[i] = (h1 ? r4[i + 2] : r4[i]) + __shfl_xor_sync(omask, h1 ? r4[i] : r4[i + 2], 2);
        total += (h0 ? r2[1] : r2[0]) + __shfl_xor_sync(omask, h0 ? r2[0] : r2[1], 1);
    }
    out[blockIdx.x * NT + threadIdx.x] = total;
}

template <int MODE>
void run(const uint4* gtab, float* out, int sms, double clk_hz, const char* what)
{
    const size_t smem = (size_t)NB * REC + (size_t)NCOLS * PT * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int steps = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, NT, smem>>>(gtab, out, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<sms, NT, smem>>>(gtab, out, steps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double px_per_sm = (double)steps * (NT / 8);
    printf("{\"mode\": %d, \"what\": \"%s\", \"ms\": %.4f, \"clk_per_px_per_sm\": %.3f}\n", MODE, what, best, best * 1e-3 * clk_hz / px_per_sm);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint4* gtab; float* out;
    cudaMalloc(&gtab, (size_t)NB * REC);
    cudaMemset(gtab, 0x3c, (size_t)NB * REC);
    cudaMalloc(&out, (size_t)prop.multiProcessorCount * NT * 4);
    run<0>(gtab, out, prop.multiProcessorCount, khz * 1e3, "sorted walk, best case: taps in registers, sort free, 16 scattered patch loads per lane and pixel");
    run<1>(gtab, out, prop.multiProcessorCount, khz * 1e3, "production walk: sliding window (4 patch loads) + 3 x 16-byte tap loads, 27 % skipped");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
