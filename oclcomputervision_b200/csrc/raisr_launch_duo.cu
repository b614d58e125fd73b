// raisr_launch_duo.cu -- launch of the two-types-per-CTA filter kernel (raisr_duo.cuh): s = 2, 24-bit tap records.
#include "raisr_launch_filter.inc"    // tensor-map helpers (templates there are not instantiated here)

#include "raisr_duo.cuh"

namespace {

template <typename OutT>
int launch_duo(raisr_ctx* h, FilterParams p, cudaStream_t st)
{
    using C = DuoCfg;
    using G = DuoGeom;
    p.tiles_x = (2 * p.ow + C::DW - 1) / C::DW;
    p.tiles_y = (p.oh + C::OTH - 1) / C::OTH;
    const size_t smem = duo_smem_bytes(p.n_buckets);
    if (smem > 227 * 1024) return 1;                       // does not fit: the caller uses the one-type kernel
    if ((p.hash_pitch % 16) || (p.hash_plane_stride % 16) || (p.hash_frame_stride % 16) || (reinterpret_cast<uintptr_t>(p.hash) % 16)) return 1;
    CUtensorMap tm, hm;
    if (int rc = make_uext_tmap(&tm, p, G::PT, G::NCOLS)) return rc;
    if (int rc = make_hash_tmap(&hm, p, 4, C::DW / 2, C::OTH)) return rc;
    const long long ntiles = (long long)p.tiles_x * p.tiles_y * p.n_frames;
    const int workers = (int)std::max<long long>(1, std::min<long long>(h->sm_count / 2, ntiles));
    auto kern = filter_duo_kernel<OutT>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<workers * 2, C::NT, smem, st>>>(p, tm, hm);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace

int raisr_launch_duo_u8(raisr_ctx* h, FilterParams p, cudaStream_t st) { return launch_duo<uint8_t>(h, p, st); }
int raisr_launch_duo_f32(raisr_ctx* h, FilterParams p, cudaStream_t st) { return launch_duo<float>(h, p, st); }
