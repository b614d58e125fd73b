// gray / band path: 8-bit output
#include "raisr_launch_filter.inc"

int raisr_launch_filter_u8(raisr_ctx* h, FilterParams p, int s, cudaStream_t st, bool single_buffer, bool allow_b24)
{
    return launch_filter<uint8_t>(h, p, s, st, single_buffer, allow_b24);
}
