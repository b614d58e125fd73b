// float output (parity probe, colour planes) and the two-plane colour kernel
#define RAISR_FILTER_WITH_OCTET2 1
#include "raisr_launch_filter.inc"

int raisr_launch_filter_f32(raisr_ctx* h, FilterParams p, int s, cudaStream_t st, bool single_buffer, bool allow_b24)
{
    return launch_filter<float>(h, p, s, st, single_buffer, allow_b24);
}

int raisr_launch_filter_octet2(raisr_ctx* h, FilterParams p, cudaStream_t st) { return launch_filter_octet2_impl(h, p, st); }
