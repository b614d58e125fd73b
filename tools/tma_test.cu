// Stand-alone check of the TMA box load used by the filter kernel (column-major uext, 3-D map).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#ifndef PT_
#define PT_ 76
#endif
#ifndef NC_
#define NC_ 137
#endif
constexpr int PT = PT_, NC = NC_;

__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, float* out, int r0, int c0, int f)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned sb = (unsigned)__cvta_generic_to_shared(smem);
    unsigned bar = sb + PT * NC * 4;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PT * NC * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sb), "l"(&tm), "r"(r0), "r"(c0), "r"(f), "r"(bar) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(bar), "r"(0) : "memory");
    const float* t = reinterpret_cast<const float*>(smem);
    for (int i = threadIdx.x; i < PT * NC; i += blockDim.x) out[i] = t[i];
}

int main()
{
    int pitch = 100, cols = 160, frames = 2;
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
    CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, PT * NC * 4 + 64));
    int r0 = 7, c0 = 11, f = 1;
    k<<<1, 256, PT * NC * 4 + 64>>>(tm, o, r0, c0, f);
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
