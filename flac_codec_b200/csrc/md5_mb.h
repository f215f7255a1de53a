// md5_mb.h -- multi-buffer MD5 on the host: eight independent streams in the lanes of one AVX2 register.
//
// The STREAMINFO signature (update_md5, src/encode.rs:1292-1318; verify, src/decode.rs:1291) is MD5 over the little-endian
// PCM of a stream: one strictly serial chain per stream -- 64 dependent steps per 64 bytes -- which neither a GPU thread
// (59 MB/s measured, md5.cu) nor a CPU core (about 0.6 GB/s) can speed up for ONE stream.  A batch of streams is another
// matter: eight of them advance in lock step in the eight 32-bit lanes of AVX2 registers at the latency of one, so the
// whole-file batch paths (flacb200_encode_batch / flacb200_decode_batch) hash at several GB/s per core while the GPU does
// the codec work.  Scalar fallback when the CPU has no AVX2.  Host code of the product (not the oracle).
#pragma once
#include <cstddef>
#include <cstdint>

namespace flacb200 {

// digests[i] = MD5(data[i][0 .. len[i])) for n streams, using up to `threads` host threads
void md5_many(const uint8_t* const* data, const size_t* len, size_t n, uint8_t (*digests)[16], unsigned threads);

}   // namespace flacb200
