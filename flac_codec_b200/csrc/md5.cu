// md5.cu -- k_md5: the STREAMINFO MD5 of many streams at once (update_md5, src/encode.rs:1292-1318; verify,
// src/decode.rs:1291-1309): MD5 over the little-endian interleaved sample bytes, ceil(bps / 8) bytes per sample.
//
// MD5 is a serial chain per stream (64 dependent steps per 64-byte block), so the grain is one THREAD per stream
// and the throughput is streams x ~100 MB/s: worth it only for batches (a C4 shard is 128 tracks per GPU), where it
// runs on a handful of warps beside the encode kernels and takes the hash off the host cores.
//   * packed little-endian input: the message IS the PCM bytes -- aligned 128-bit loads, the 16 message words of a block
//     cut out of them with byte permutes (streams start on sample, not word, boundaries), the next block's loads in
//     flight while the current one is hashed
//   * other layouts (big-endian bytes, int32 interleaved / planar): samples are narrowed to little-endian bytes one at
//     a time into the block (rare path)
#include "common.cuh"
#include "../../include/flacb200.h"

namespace flacb200 {

struct Md5Seg {
    unsigned long long pcm_off, n_pcm;   // inter-channel samples
};

struct Md5Cfg {
    uint32_t channels, bytes_per_sample, pcm_kind, nseg;
    unsigned long long planar_stride;
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int s) { return __funnelshift_l(x, x, s); }

// one 64-byte block (RFC 1321, the same step order as the host copy in stream.cpp)
__device__ __forceinline__ void md5_block(uint32_t (&h)[4], const uint32_t (&m)[16])
{
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
#define F1(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define F2(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define F3(x, y, z) ((x) ^ (y) ^ (z))
#define F4(x, y, z) ((y) ^ ((x) | ~(z)))
#define ST(f, w, x, y, z, k, t, s) w = x + rotl32(w + f(x, y, z) + m[k] + t, s);
    ST(F1, a, b, c, d, 0, 0xd76aa478u, 7) ST(F1, d, a, b, c, 1, 0xe8c7b756u, 12) ST(F1, c, d, a, b, 2, 0x242070dbu, 17) ST(F1, b, c, d, a, 3, 0xc1bdceeeu, 22)
    ST(F1, a, b, c, d, 4, 0xf57c0fafu, 7) ST(F1, d, a, b, c, 5, 0x4787c62au, 12) ST(F1, c, d, a, b, 6, 0xa8304613u, 17) ST(F1, b, c, d, a, 7, 0xfd469501u, 22)
    ST(F1, a, b, c, d, 8, 0x698098d8u, 7) ST(F1, d, a, b, c, 9, 0x8b44f7afu, 12) ST(F1, c, d, a, b, 10, 0xffff5bb1u, 17) ST(F1, b, c, d, a, 11, 0x895cd7beu, 22)
    ST(F1, a, b, c, d, 12, 0x6b901122u, 7) ST(F1, d, a, b, c, 13, 0xfd987193u, 12) ST(F1, c, d, a, b, 14, 0xa679438eu, 17) ST(F1, b, c, d, a, 15, 0x49b40821u, 22)
    ST(F2, a, b, c, d, 1, 0xf61e2562u, 5) ST(F2, d, a, b, c, 6, 0xc040b340u, 9) ST(F2, c, d, a, b, 11, 0x265e5a51u, 14) ST(F2, b, c, d, a, 0, 0xe9b6c7aau, 20)
    ST(F2, a, b, c, d, 5, 0xd62f105du, 5) ST(F2, d, a, b, c, 10, 0x02441453u, 9) ST(F2, c, d, a, b, 15, 0xd8a1e681u, 14) ST(F2, b, c, d, a, 4, 0xe7d3fbc8u, 20)
    ST(F2, a, b, c, d, 9, 0x21e1cde6u, 5) ST(F2, d, a, b, c, 14, 0xc33707d6u, 9) ST(F2, c, d, a, b, 3, 0xf4d50d87u, 14) ST(F2, b, c, d, a, 8, 0x455a14edu, 20)
    ST(F2, a, b, c, d, 13, 0xa9e3e905u, 5) ST(F2, d, a, b, c, 2, 0xfcefa3f8u, 9) ST(F2, c, d, a, b, 7, 0x676f02d9u, 14) ST(F2, b, c, d, a, 12, 0x8d2a4c8au, 20)
    ST(F3, a, b, c, d, 5, 0xfffa3942u, 4) ST(F3, d, a, b, c, 8, 0x8771f681u, 11) ST(F3, c, d, a, b, 11, 0x6d9d6122u, 16) ST(F3, b, c, d, a, 14, 0xfde5380cu, 23)
    ST(F3, a, b, c, d, 1, 0xa4beea44u, 4) ST(F3, d, a, b, c, 4, 0x4bdecfa9u, 11) ST(F3, c, d, a, b, 7, 0xf6bb4b60u, 16) ST(F3, b, c, d, a, 10, 0xbebfbc70u, 23)
    ST(F3, a, b, c, d, 13, 0x289b7ec6u, 4) ST(F3, d, a, b, c, 0, 0xeaa127fau, 11) ST(F3, c, d, a, b, 3, 0xd4ef3085u, 16) ST(F3, b, c, d, a, 6, 0x04881d05u, 23)
    ST(F3, a, b, c, d, 9, 0xd9d4d039u, 4) ST(F3, d, a, b, c, 12, 0xe6db99e5u, 11) ST(F3, c, d, a, b, 15, 0x1fa27cf8u, 16) ST(F3, b, c, d, a, 2, 0xc4ac5665u, 23)
    ST(F4, a, b, c, d, 0, 0xf4292244u, 6) ST(F4, d, a, b, c, 7, 0x432aff97u, 10) ST(F4, c, d, a, b, 14, 0xab9423a7u, 15) ST(F4, b, c, d, a, 5, 0xfc93a039u, 21)
    ST(F4, a, b, c, d, 12, 0x655b59c3u, 6) ST(F4, d, a, b, c, 3, 0x8f0ccc92u, 10) ST(F4, c, d, a, b, 10, 0xffeff47du, 15) ST(F4, b, c, d, a, 1, 0x85845dd1u, 21)
    ST(F4, a, b, c, d, 8, 0x6fa87e4fu, 6) ST(F4, d, a, b, c, 15, 0xfe2ce6e0u, 10) ST(F4, c, d, a, b, 6, 0xa3014314u, 15) ST(F4, b, c, d, a, 13, 0x4e0811a1u, 21)
    ST(F4, a, b, c, d, 4, 0xf7537e82u, 6) ST(F4, d, a, b, c, 11, 0xbd3af235u, 10) ST(F4, c, d, a, b, 2, 0x2ad7d2bbu, 15) ST(F4, b, c, d, a, 9, 0xeb86d391u, 21)
#undef ST
#undef F1
#undef F2
#undef F3
#undef F4
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

// byte i of the stream's message: sample (i / B) narrowed to B little-endian bytes (rare layouts)
__device__ inline uint32_t md5_message_byte(const uint8_t* __restrict__ pcm, const Md5Cfg& cfg, const Md5Seg& sg, unsigned long long i)
{
    const uint32_t B = cfg.bytes_per_sample;
    const unsigned long long smp = i / B;
    const uint32_t k = (uint32_t)(i % B);
    const unsigned long long frame = sg.pcm_off + smp / cfg.channels;
    const uint32_t ch = (uint32_t)(smp % cfg.channels);
    if (cfg.pcm_kind == FLACB200_PCM_BYTES_LE) return pcm[(frame * cfg.channels + ch) * B + k];
    if (cfg.pcm_kind == FLACB200_PCM_BYTES_BE) return pcm[(frame * cfg.channels + ch) * B + (B - 1 - k)];
    const int32_t* p = reinterpret_cast<const int32_t*>(pcm);
    const int32_t v = cfg.pcm_kind == FLACB200_PCM_I32_INTERLEAVED ? p[frame * cfg.channels + ch] : p[(unsigned long long)ch * cfg.planar_stride + frame];
    return ((uint32_t)v >> (8 * k)) & 0xffu;
}

__global__ void __launch_bounds__(32) k_md5(Md5Cfg cfg, const uint8_t* __restrict__ pcm, const Md5Seg* __restrict__ segs, uint32_t* __restrict__ digests)
{
    const uint32_t t = blockIdx.x * 32 + threadIdx.x;
    if (t >= cfg.nseg) return;
    const Md5Seg sg = segs[t];
    const unsigned long long len = sg.n_pcm * cfg.channels * cfg.bytes_per_sample;   // message bytes
    uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    unsigned long long done = 0;
    if (cfg.pcm_kind == FLACB200_PCM_BYTES_LE && len >= 64) {
        // the message is the buffer itself from byte `start` on: aligned 16-byte loads, message word j of a block is
        // bytes [4j + sh, 4j + sh + 4) of the 80 aligned bytes that cover the block (sh = start % 4 after aligning to 16:
        // handled as a word offset wo = (start % 16) / 4 and a byte shift sh)
        const unsigned long long start = sg.pcm_off * cfg.channels * cfg.bytes_per_sample;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(pcm) + start;
        const uint4* q = reinterpret_cast<const uint4*>(addr & ~(uintptr_t)15);
        const uint32_t mis = (uint32_t)(addr & 15), wo = mis >> 2, sh = mis & 3;
        const uint32_t sel = 0x3210u + 0x1111u * sh;
        const unsigned long long nblocks = len / 64;
        // every chunk that is loaded holds at least one byte of the message, so the loads stay inside the (at least
        // 16-byte granular) device allocation the message lives in
        const unsigned long long fast_blocks = nblocks;
        uint4 c0, c1, c2, c3, c4;
        if (fast_blocks) { c0 = q[0]; c1 = q[1]; c2 = q[2]; c3 = q[3]; c4 = mis ? q[4] : make_uint4(0, 0, 0, 0); }
        for (unsigned long long b = 0; b < fast_blocks; b++) {
            const uint32_t w[20] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w, c4.x, c4.y, c4.z, c4.w};
            if (b + 1 < fast_blocks) {   // next block's chunks, in flight while this one is hashed
                q += 4;
                c0 = c4;
                if (mis == 0) c0 = q[0];
                c1 = q[1]; c2 = q[2]; c3 = q[3];
                if (mis) c4 = q[4];
            }
            uint32_t m[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                // words wo + j and wo + j + 1 of w[]; wo is uniform per thread but not a compile-time constant
                const uint32_t lo = wo == 0 ? w[j] : wo == 1 ? w[j + 1] : wo == 2 ? w[j + 2] : w[j + 3];
                const uint32_t hi = wo == 0 ? w[j + 1] : wo == 1 ? w[j + 2] : wo == 2 ? w[j + 3] : w[j + 4];
                m[j] = __byte_perm(lo, hi, sel);
            }
            md5_block(h, m);
        }
        done = fast_blocks * 64;
    }
    // whole blocks of the other layouts, then the tail with the padding (0x80, zeros, 64-bit bit length)
    for (;;) {
        uint32_t m[16];
#pragma unroll
        for (int j = 0; j < 16; j++) m[j] = 0;
        const unsigned long long left = len - done;
        const uint32_t take = left >= 64 ? 64u : (uint32_t)left;
        for (uint32_t i = 0; i < take; i++) {
            const uint32_t byte = md5_message_byte(pcm, cfg, sg, done + i);
#pragma unroll
            for (int j = 0; j < 16; j++)
                if ((i >> 2) == (uint32_t)j) m[j] |= byte << (8 * (i & 3));
        }
        done += take;
        if (take == 64) {
            md5_block(h, m);
            continue;
        }
#pragma unroll
        for (int j = 0; j < 16; j++)
            if ((take >> 2) == (uint32_t)j) m[j] |= 0x80u << (8 * (take & 3));
        if (take >= 56) {   // no room for the length: it goes into one more block
            md5_block(h, m);
#pragma unroll
            for (int j = 0; j < 16; j++) m[j] = 0;
        }
        const unsigned long long bits = len * 8;
        m[14] = (uint32_t)bits;
        m[15] = (uint32_t)(bits >> 32);
        md5_block(h, m);
        break;
    }
    digests[t * 4 + 0] = h[0];
    digests[t * 4 + 1] = h[1];
    digests[t * 4 + 2] = h[2];
    digests[t * 4 + 3] = h[3];
}

cudaError_t launch_md5(const Md5Cfg& cfg, const uint8_t* pcm, const Md5Seg* segs, uint32_t* digests, cudaStream_t st)
{
    if (cfg.nseg == 0) return cudaSuccess;
    count_launch(), k_md5<<<(cfg.nseg + 31) / 32, 32, 0, st>>>(cfg, pcm, segs, digests);
    return cudaGetLastError();
}

}   // namespace flacb200
