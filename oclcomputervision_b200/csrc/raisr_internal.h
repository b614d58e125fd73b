// raisr_internal.h -- state shared by the translation units of libraisr_b200.so (not part of the C-ABI).
#pragma once
#include "../../include/raisr_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "raisr_filter.cuh"   // FilterParams
#include "raisr_prep.cuh"     // PrepParams

int raisr_fail(int code, const char* fmt, ...);   // records the thread-local message of raisr_last_error(), returns code
template <typename... Args>
inline int fail(int code, const char* fmt, Args... args) { return raisr_fail(code, fmt, args...); }

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(e__ == cudaErrorMemoryAllocation ? RAISR_E_NOMEM : RAISR_E_CUDA,       \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,     \
                        __LINE__);                                                             \
    } while (0)


struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) return fail(RAISR_E_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        bytes = need;
        return 0;
    }
    // as ensure(), and a (re)allocated buffer starts out all zero
    int ensure_zero(size_t need)
    {
        if (need <= bytes) return 0;
        if (int rc = ensure(need)) return rc;
        if (cudaMemset(p, 0, bytes) != cudaSuccess) return fail(RAISR_E_CUDA, "cudaMemset(%zu) failed", bytes);
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct ScaleTable {
    DevBuf block;   // [type][bucket][132]  (filter_block_kernel)
    DevBuf octet;   // [type][bucket][128]  (filter_octet_kernel, lane-major chunks)
    DevBuf octet16; // [type][bucket][128 halfs]  (filter_octet_kernel<kTapsF16>, only with taps = fp16)
    DevBuf b24;     // [type][bucket][384 bytes]  (filter_octet_kernel<kTapsB24>, taps = b24 / auto)
    std::vector<float> host;  // the caller's table as given (repacked when the tap precision option changes)
    std::vector<float> eff;   // the tap values the active gray kernel multiplies by, reference layout
    int format = raisr::kTapsF32;    // tap format the gray octet kernel uses for this scale
    float b24_bound = 0;      // max over filters of max(sum of positive, sum of negative tap errors): bounds |out_b24 - out_fp32| for patches in [0,1]
    bool set = false;
};


struct raisr_ctx {
    int device = 0;
    int n_angle = 24, n_strength = 3, n_coherence = 3;
    int n_buckets = 216;
    int sm_count = 0, clock_khz = 0;
    char name[128] = {0};
    float sq[raisr::kMaxQ], cq[raisr::kMaxQ];
    cudaStream_t own_stream = nullptr, h2d_stream = nullptr, d2h_stream = nullptr;
    cudaStream_t user_stream = nullptr;
    bool use_user_stream = false;
    ScaleTable tables[5];  // index = scale (2..4)
    DevBuf uext, hash, dsrc[2], ddst[2], dbg;
    DevBuf uext2, hash2;          // second scratch set of the overlapped pipeline
    DevBuf cplanes;               // colour path: four filtered float planes
    DevBuf tens;                  // eigen_in_filter: three float planes (ma, mb, md) with the geometry of the hash image
    cudaStream_t prep_stream = nullptr, filt_stream = nullptr;
    int overlap = 0;              // 1: prep of chunk c+1 shares the SMs with the filter of chunk c
    std::vector<cudaEvent_t> ev_pool;
    // the uext / hash / cplanes scratch is shared by every entry point: work that uses it on another stream than
    // the previous user first waits for that user's last kernel (an un-synchronised DEVICE call on the caller's
    // stream followed by a HOST call on own_stream would otherwise race)
    cudaEvent_t scratch_ev = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_busy = false;
    void scratch_acquire(cudaStream_t st)
    {
        if (scratch_busy && scratch_stream != st) cudaStreamWaitEvent(st, scratch_ev, 0);
    }
    void scratch_release(cudaStream_t st)
    {
        if (!scratch_ev) cudaEventCreateWithFlags(&scratch_ev, cudaEventDisableTiming);
        cudaEventRecord(scratch_ev, st);
        scratch_stream = st;
        scratch_busy = true;
    }
    cudaEvent_t timer_ev[8] = {};  // raisr_timer_mark slots
    int* flag_err = nullptr;      // device word set by a raisr_flag_wait that timed out (reported by raisr_sync)
    long long launches = 0;
    float last_prep_ms = 0, last_filter_ms = 0;
    int filter_impl = 1;  // 0 = block (v1), 1 = octet
    int prep_impl = 2;    // 2 = packed-fp32 prep2_kernel, 1 = scalar prep_kernel
    int filter_pipe = 1;        // 1 = mbarrier full/empty tile pipeline with a producer warp (default), 0 = CTA-wide barriers per tile
    int color_filter_impl = 2;  // 2 = two planes per CTA (s = 2, fp32 taps), 1 = one launch per plane
    int prep_ctas_per_sm = 0;   // 0 = one prep CTA per tile (default), n > 0 = persistent grid of n CTAs per SM striding over the tiles
    int eig = 0;          // "eigen_in_filter": 1 = prep stores the structure tensor and the s = 2 b24 filter kernel does the eigen-solve / hash
    int duo = 1;          // "filter_duo": 1 = two pixel types per CTA for s = 2 with b24 records (default), 0 = one type per CTA
    int resize_fast = 1;  // 1 = four-pixels-per-thread kernel for gray bilinear resizes that qualify (default), 0 = generic kernel
    int cubic = 0;        // "cheap_upscaler" option: 1 = bicubic stage 1 (gray path, prep2 kernel)
    int as_written = 0;   // "quirks" option
    int taps_fp16 = 0;    // 1 when taps_mode == kTapsF16 (kept for the colour path and the table packer)
    int taps_mode = 3;    // "taps" option: 0 fp32, 1 fp16, 2 b24, 3 auto (b24 where its output bound holds, else fp32)
    size_t chunk_budget = 208u << 20;   // upscaled-image scratch per kernel launch: 6 frames of 1080p->4K

    cudaStream_t stream() const { return use_user_stream ? user_stream : own_stream; }
    cudaEvent_t ev(size_t i)
    {
        while (ev_pool.size() <= i) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[i];
    }
};


// kernel launchers, one translation unit each (raisr_launch_*.cu) so that the library builds in parallel
int raisr_launch_prep(raisr_ctx* h, const raisr::PrepParams& p, int s, cudaStream_t st, bool dbg, int ctas_per_sm);
int raisr_launch_filter_u8(raisr_ctx* h, raisr::FilterParams p, int s, cudaStream_t st, bool single_buffer, bool allow_b24);
int raisr_launch_filter_f32(raisr_ctx* h, raisr::FilterParams p, int s, cudaStream_t st, bool single_buffer, bool allow_b24);
// s = 2 with 24-bit tap records: both pixel types of an output row per CTA (raisr_duo.cuh); p.table = the b24 table.
// Returns 1 when the kernel cannot be used for this call (the caller falls back to the one-type kernel).
int raisr_launch_duo_u8(raisr_ctx* h, raisr::FilterParams p, cudaStream_t st);
int raisr_launch_duo_f32(raisr_ctx* h, raisr::FilterParams p, cudaStream_t st);
// colour path, s = 2, fp32 taps: all four planes in one launch (returns 1 when the kernel does not fit: fall back per plane)
int raisr_launch_filter_octet2(raisr_ctx* h, raisr::FilterParams p, cudaStream_t st);
