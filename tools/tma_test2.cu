// TMA check with the libcu++ experimental wrappers (known-good reference for the raw PTX version).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#ifndef PT_
#define PT_ 76
#endif
#ifndef NC_
#define NC_ 137
#endif
#ifndef R0_
#define R0_ 7
#endif
#ifndef RANK_
#define RANK_ 3
#endif
constexpr int PT = PT_, NC = NC_;

__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, float* out, int r0, int c0, int f, int mode)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* t = reinterpret_cast<float*>(smem);
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (RANK_ == 3) cde::cp_async_bulk_tensor_3d_global_to_shared(t, &tm, r0, c0, f, bar);
        else cde::cp_async_bulk_tensor_2d_global_to_shared(t, &tm, r0, c0, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, PT * NC * 4);
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < PT * NC; i += blockDim.x) out[i] = t[i];
}

int main()
{
    int pitch = 256, cols = 256, frames = RANK_ == 3 ? 2 : 1;
    std::vector<float> h((size_t)pitch * cols * frames);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, PT * NC * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)cols, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * cols * 4};
    cuuint32_t box[3] = {PT, NC, 1}, es[3] = {1, 1, 1};
    CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, RANK_, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, PT * NC * 4 + 64));
    int r0 = R0_, c0 = 11, f = RANK_ == 3 ? 1 : 0;
    k<<<1, 256, PT * NC * 4 + 64>>>(tm, o, r0, c0, f, 3);
    CK(cudaDeviceSynchronize());
    std::vector<float> res(PT * NC);
    CK(cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < NC; ++c) for (int rr = 0; rr < PT; ++rr) {
        float want = (float)((size_t)f * pitch * cols + (size_t)(c0 + c) * pitch + (r0 + rr));
        if (r0 + rr >= pitch || c0 + c >= cols) want = 0.f;
        if (res[c * PT + rr] != want) { if (bad < 5) printf("mismatch c=%d r=%d got %f want %f\n", c, rr, res[c * PT + rr], want); ++bad; }
    }
    printf("bad=%d\n", bad);
    return 0;
}
