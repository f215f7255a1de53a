// pack_bits.cuh -- bit emission and CRC-16 pieces shared by the frame packers (k_pack3 in encode_frame.cu, k_frame4 in
// encode_analyze.cu): OR-ing codes into a big-endian bit image in shared memory, and the wide-folding CRC-16 over that image.
// The CRC tables live in device memory once per translation unit that includes this header (no relocatable device code
// in this build); each unit's tables are built at engine creation (init_*_tables).
#pragma once
#include "common.cuh"
#include "crc.cuh"

namespace flacb200 {

static __device__ Crc16Fold g_crc16_tabs;   // built once per device by k_crc16_tables_init

static __device__ uint16_t g_crc16_xblk[1024];   // x^(1024 j) mod P: shifts a CRC over j blocks of 32 words

static __global__ void k_crc16_tables_init()
{
    crc16_fold_init(g_crc16_tabs);
    for (uint32_t j = threadIdx.x; j < 1024; j += blockDim.x) g_crc16_xblk[j] = (uint16_t)gf16_xpow8(128 * j);
}

// OR the low nbits (1..32) of v into the big-endian bit image `words` at bit position pos
__device__ inline void p3_put(uint32_t words_sa, uint32_t pos, uint32_t nbits, uint32_t v)
{
    const uint32_t w = pos >> 5, off = pos & 31;
    const unsigned long long wide = ((unsigned long long)v) << (64u - nbits - off);
    const uint32_t hi = (uint32_t)(wide >> 32), lo = (uint32_t)wide;
    // reductions without a return value; the second word is touched only when the code straddles (predicated, no branch);
    // words_sa: shared-window address of the image (converted once per kernel, not per code)
    const uint32_t a = words_sa + 4u * w;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "red.shared.or.b32 [%0], %1;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p red.shared.or.b32 [%0+4], %2;\n\t}"
        ::"r"(a), "r"(hi), "r"(lo) : "memory");
}

__device__ inline void p3_put_masked(uint32_t words_sa, uint32_t pos, uint32_t nbits, uint32_t v)
{
    if (nbits == 0) return;
    if (nbits < 32) v &= (1u << nbits) - 1u;
    p3_put(words_sa, pos, nbits, v);
}

// CRC-16 of the message bytes held in words[w0 .. w0 + nw) (big-endian words: the first message byte is the top byte;
// w0 even).  One warp, all lanes call and get the result.  A lane folds every 32nd PAIR of words (crc.cuh); an odd last
// word is appended by the classic byte steps.
__device__ inline uint32_t p3_crc_words(const Crc16Fold& t, const uint32_t* words, uint32_t w0, uint32_t nw)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t np = nw >> 1;
    const uint2* pairs = reinterpret_cast<const uint2*>(words + w0);
    uint32_t acc = 0, last = 0;
    bool any = false;
    for (uint32_t i = lane; i < np; i += 32) {
        const uint2 v = pairs[i];
        const uint32_t f = crc16_f8(t, v.x, v.y);
        acc = (t.m_hi[acc >> 8] ^ t.m_lo[acc & 0xff]) ^ f;   // acc * x^2048 + F8: 32 pairs lie between two pairs of a lane
        last = i;
        any = true;
    }
    uint32_t part = any ? gf16_mulmod(acc, t.xd2[np - 1 - last]) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
    part &= 0xffffu;
    if (nw & 1u) {   // the odd word: shift by four bytes, add its CRC
        const uint32_t v = words[w0 + nw - 1];
        part = gf16_mulmod(part, t.xd[1]) ^ t.T[3][v >> 24] ^ t.T[2][(v >> 16) & 0xff] ^ t.T[1][(v >> 8) & 0xff] ^ t.T[0][v & 0xff];
    }
    return part & 0xffffu;
}

}   // namespace flacb200
