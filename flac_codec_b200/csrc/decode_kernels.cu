// decode_kernels.cu -- placeholder until the decode pipeline lands (next commit)
#include "../../include/flacb200.h"
#include "common.cuh"
namespace flacb200 {
int decode_impl(flacb200_engine*, const flacb200_stream_params*, const void*, size_t, int, const flacb200_decode_segment*, size_t, void*, size_t,
                int, int, uint64_t, uint64_t*, uint64_t*, uint64_t*)
{
    return FLACB200_E_BAD_ARGUMENT;
}
}   // namespace flacb200
