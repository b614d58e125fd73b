// raisr_launch_prep.cu -- launches of kernel A (raisr_prep.cuh / raisr_prep2.cuh); see raisr_api.cu for the pipeline.
#include "raisr_internal.h"

#include <algorithm>

#include "raisr_prep2.cuh"

using namespace raisr;

namespace {

template <int S, bool DBG, int NQ, bool FROM_U>
void launch_prep_q(PrepParams p, cudaStream_t st, int max_ctas)
{
    p.tiles_x = (p.dw + PT_W - 1) / PT_W;
    p.tiles_y = (p.rows + PT_H - 1) / PT_H;
    long long total = (long long)p.tiles_x * p.tiles_y * p.n_frames;
    int grid = (int)std::max<long long>(1, std::min<long long>(total, max_ctas));
    size_t smem = sizeof(PrepSmem);
    cudaFuncSetAttribute(prep_kernel<S, DBG, NQ, FROM_U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    prep_kernel<S, DBG, NQ, FROM_U><<<grid, PT_THREADS, smem, st>>>(p);
}

template <int S, bool DBG, int NQ, bool FROM_U, bool CUBIC = false>
void launch_prep2_q(PrepParams p, cudaStream_t st, int max_ctas)
{
    p.tiles_x = (p.dw + PT_W - 1) / PT_W;
    p.tiles_y = (p.rows + PT_H - 1) / PT_H;
    long long total = (long long)p.tiles_x * p.tiles_y * p.n_frames;
    int grid = (int)std::max<long long>(1, std::min<long long>(total, max_ctas));
    size_t smem = sizeof(Prep2Smem);
    cudaFuncSetAttribute(prep2_kernel<S, DBG, NQ, FROM_U, CUBIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    prep2_kernel<S, DBG, NQ, FROM_U, CUBIC><<<grid, PT_THREADS, smem, st>>>(p);
}

// impl 2: packed-fp32 kernel (raisr_prep2.cuh, default); impl 1: scalar kernel (raisr_prep.cuh)
template <int S>
void launch_prep_t(const PrepParams& p, cudaStream_t st, bool dbg, int max_ctas, int impl)
{
    const bool small = p.n_strength <= 3 && p.n_coherence <= 3;   // the reference's 3 x 3 (raisr.cl:9-15)
    if (impl == 2) {
        if (p.cubic) {   // optional stage-1 variant: the general-quantiser instantiation only
            dbg ? launch_prep2_q<S, true, kMaxQ, false, true>(p, st, max_ctas) : launch_prep2_q<S, false, kMaxQ, false, true>(p, st, max_ctas);
            return;
        }
        if (p.uext_in && dbg) launch_prep2_q<S, true, kMaxQ, true>(p, st, max_ctas);   // parity probe of the colour path
        else if (p.uext_in) small ? launch_prep2_q<S, false, 2, true>(p, st, max_ctas) : launch_prep2_q<S, false, kMaxQ, true>(p, st, max_ctas);
        else if (dbg) small ? launch_prep2_q<S, true, 2, false>(p, st, max_ctas) : launch_prep2_q<S, true, kMaxQ, false>(p, st, max_ctas);
        else small ? launch_prep2_q<S, false, 2, false>(p, st, max_ctas) : launch_prep2_q<S, false, kMaxQ, false>(p, st, max_ctas);
        return;
    }
    if (p.uext_in) {   // colour path: hash an existing upscaled plane
        small ? launch_prep_q<S, false, 2, true>(p, st, max_ctas) : launch_prep_q<S, false, kMaxQ, true>(p, st, max_ctas);
        return;
    }
    if (dbg) small ? launch_prep_q<S, true, 2, false>(p, st, max_ctas) : launch_prep_q<S, true, kMaxQ, false>(p, st, max_ctas);
    else small ? launch_prep_q<S, false, 2, false>(p, st, max_ctas) : launch_prep_q<S, false, kMaxQ, false>(p, st, max_ctas);
}

// ctas_per_sm == 0: one CTA per tile (the hardware scheduler balances); > 0: persistent grid of that many CTAs per SM
}  // namespace

int raisr_launch_prep(raisr_ctx* h, const PrepParams& p, int s, cudaStream_t st, bool dbg, int ctas_per_sm)
{
    const int max_ctas = ctas_per_sm > 0 ? h->sm_count * ctas_per_sm : 0x7fffffff;
    if (p.cubic && (h->prep_impl != 2 || p.uext_in))
        return fail(RAISR_E_UNSUPPORTED, "the bicubic cheap upscaler is built for the gray path of prep2_kernel only");
    switch (s) {
    case 2: launch_prep_t<2>(p, st, dbg, max_ctas, h->prep_impl); break;
    case 3: launch_prep_t<3>(p, st, dbg, max_ctas, h->prep_impl); break;
    case 4: launch_prep_t<4>(p, st, dbg, max_ctas, h->prep_impl); break;
    default: return fail(RAISR_E_UNSUPPORTED, "scale %d not supported (2, 3 or 4)", s);
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
