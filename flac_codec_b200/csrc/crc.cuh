// crc.cuh -- warp-cooperative CRC-16 (poly 0x8005, init 0, MSB first; src/crc.rs:144-188).
//
// A frame is read as coalesced 32-bit words: lane l owns words l, l+32, l+64, ... of the frame body and
// folds them with Horner steps  acc = acc * x^1024 + F(word)  (x^1024: the 128 bytes between two words of
// one lane), F(word) = word(x) * x^16 mod P.  The 32 lane results are shifted to the end of the message
// with x^(32 d) (d = words that follow the lane's last word) and XOR-reduced.  All constants live in a
// small shared-memory block built once per CTA.
#pragma once
#include "common.cuh"

namespace flacb200 {

struct Crc16Tables {
    uint16_t byte_tab[256];   // classic table: crc of one byte
    uint16_t mul_lo[256];     // (i) * x^1024 mod P
    uint16_t mul_hi[256];     // (i << 8) * x^1024 mod P
    uint16_t xd[33];          // x^(32 d) mod P, d = 0..32
    uint16_t xb[4];           // x^(8 t) mod P, t = 0..3
};

// x^(8 n) mod P by square-and-multiply
__device__ inline uint32_t gf16_xpow8(uint32_t n)
{
    uint32_t r = 1, b = 0x0100;   // x^8
    while (n) {
        if (n & 1u) r = gf16_mulmod(r, b);
        b = gf16_mulmod(b, b);
        n >>= 1;
    }
    return r;
}

// all threads of the CTA call this once (followed by __syncthreads by the caller)
__device__ inline void crc16_tables_init(Crc16Tables& t)
{
    const uint32_t x1024 = gf16_xpow8(128);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        t.byte_tab[i] = crc16_table_entry(i);
        t.mul_lo[i] = (uint16_t)gf16_mulmod(i, x1024);
        t.mul_hi[i] = (uint16_t)gf16_mulmod(i << 8, x1024);
    }
    for (uint32_t i = threadIdx.x; i < 33; i += blockDim.x) t.xd[i] = (uint16_t)gf16_xpow8(4 * i);
    if (threadIdx.x < 4) t.xb[threadIdx.x] = (uint16_t)gf16_xpow8(threadIdx.x);
}

__device__ inline uint32_t crc16_byte(const Crc16Tables& t, uint32_t crc, uint32_t b)
{
    return (t.byte_tab[((crc >> 8) ^ b) & 0xff] ^ (crc << 8)) & 0xffffu;
}

// CRC-16 of bytes[start, start + len) computed by one warp (all 32 lanes must call; all get the result).
// `bytes` must be 4-byte aligned (the buffer base); start and len are arbitrary.
__device__ inline uint32_t crc16_warp(const Crc16Tables& t, const uint8_t* __restrict__ bytes, unsigned long long start,
                                      unsigned long long len)
{
    const uint32_t lane = threadIdx.x & 31;
    const unsigned long long end = start + len;
    unsigned long long a = (start + 3) & ~3ull;   // first aligned byte
    if (a > end) a = end;
    const uint32_t head = (uint32_t)(a - start);
    const unsigned long long nwords = (end - a) >> 2;
    const uint32_t tail = (uint32_t)((end - a) & 3);
    uint32_t acc = 0;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(bytes + a);
    unsigned long long last = 0;
    bool any = false;
    for (unsigned long long i = lane; i < nwords; i += 32) {
        const uint32_t v = w[i];   // memory order = message order: byte 0 is the low byte
        uint32_t f = t.byte_tab[v & 0xff];
        f = crc16_byte(t, f, (v >> 8) & 0xff);
        f = crc16_byte(t, f, (v >> 16) & 0xff);
        f = crc16_byte(t, f, v >> 24);
        acc = (t.mul_hi[acc >> 8] ^ t.mul_lo[acc & 0xff]) ^ f;
        last = i;
        any = true;
    }
    uint32_t part = 0;
    if (any) {
        const unsigned long long after = nwords - 1 - last;   // words that follow this lane's last word (0..31)
        part = gf16_mulmod(acc, t.xd[(uint32_t)after]);
        if (tail) part = gf16_mulmod(part, t.xb[tail]);
    }
    if (lane == 0) {
        // head bytes sit before everything else: shift them over body and tail
        uint32_t h = 0;
        for (uint32_t i = 0; i < head; i++) h = crc16_byte(t, h, bytes[start + i]);
        if (head && (nwords || tail)) h = gf16_mulmod(h, gf16_xpow8((uint32_t)(nwords * 4 + tail)));
        uint32_t tl = 0;
        for (uint32_t i = 0; i < tail; i++) tl = crc16_byte(t, tl, bytes[a + nwords * 4 + i]);
        part ^= h ^ tl;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
    return part & 0xffffu;
}

// ---- wide folding: 8 message bytes per step, look-ups independent of one another -----------------------------------
// T[j][b] = b(x) * x^(16 + 8 j) mod P: the CRC of byte b followed by j zero bytes.  For 8 message bytes b0..b7 (b0 first)
// F8 = T[7][b0] ^ T[6][b1] ^ ... ^ T[0][b7] is the CRC of those 8 bytes; a lane that owns every 32nd group of 8 bytes folds
// acc = acc * x^2048 + F8 (two more look-ups, the only ones that depend on the previous step).
struct Crc16Fold {
    uint16_t T[8][256];
    uint16_t m_lo[256], m_hi[256];   // (i) * x^2048, (i << 8) * x^2048
    uint16_t xd2[33];                // x^(64 d), d = 0..32
    uint16_t xd[33];                 // x^(32 d), d = 0..32
};

__device__ inline void crc16_fold_init(Crc16Fold& t)
{
    const uint32_t x2048 = gf16_xpow8(256);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = crc16_table_entry(i);
        for (int j = 0; j < 8; j++) {
            t.T[j][i] = (uint16_t)c;
            c = gf16_mulmod(c, 0x0100);   // one more zero byte
        }
        t.m_lo[i] = (uint16_t)gf16_mulmod(i, x2048);
        t.m_hi[i] = (uint16_t)gf16_mulmod(i << 8, x2048);
    }
    for (uint32_t i = threadIdx.x; i < 33; i += blockDim.x) {
        t.xd2[i] = (uint16_t)gf16_xpow8(8 * i);
        t.xd[i] = (uint16_t)gf16_xpow8(4 * i);
    }
}

// CRC of 8 message bytes held as two big-endian words (hi = the first four bytes)
__device__ inline uint32_t crc16_f8(const Crc16Fold& t, uint32_t hi, uint32_t lo)
{
    const uint32_t a = t.T[7][hi >> 24] ^ t.T[6][(hi >> 16) & 0xff], b = t.T[5][(hi >> 8) & 0xff] ^ t.T[4][hi & 0xff];
    const uint32_t c = t.T[3][lo >> 24] ^ t.T[2][(lo >> 16) & 0xff], d = t.T[1][(lo >> 8) & 0xff] ^ t.T[0][lo & 0xff];
    return (a ^ b) ^ (c ^ d);
}


// CRC-16 of bytes[start, start + len) in global memory by one warp with the wide folding above (all 32 lanes call, all get
// the result).  `bytes` must be 8-byte aligned (the buffer base); start and len are arbitrary (len < 2^34).  Lane l folds every
// 32nd group of eight bytes; the bytes in front of the first aligned group ride on lane 0's first term, the bytes behind the
// last group are appended by classic byte steps.
// F8 of a group straight from the two little-endian words of the load (message byte 0 = low byte of lo): one PRMT per byte
__device__ __forceinline__ uint32_t crc16_f8_le(const Crc16Fold& t, uint32_t w0, uint32_t w1)
{
    const uint32_t a = t.T[7][__byte_perm(w0, 0, 0x4440)] ^ t.T[6][__byte_perm(w0, 0, 0x4441)];
    const uint32_t b = t.T[5][__byte_perm(w0, 0, 0x4442)] ^ t.T[4][__byte_perm(w0, 0, 0x4443)];
    const uint32_t c = t.T[3][__byte_perm(w1, 0, 0x4440)] ^ t.T[2][__byte_perm(w1, 0, 0x4441)];
    const uint32_t d = t.T[1][__byte_perm(w1, 0, 0x4442)] ^ t.T[0][__byte_perm(w1, 0, 0x4443)];
    return (a ^ b) ^ (c ^ d);
}

__device__ inline uint32_t crc16_warp_fold(const Crc16Fold& t, const uint8_t* __restrict__ bytes, unsigned long long start, unsigned long long len)
{
    const uint32_t lane = threadIdx.x & 31;
    const unsigned long long end = start + len;
    auto byte_step = [&](uint32_t crc, uint32_t b) { return (t.T[0][((crc >> 8) ^ b) & 0xff] ^ (crc << 8)) & 0xffffu; };
    auto fold = [&](uint32_t acc, uint32_t f) { return (uint32_t)(t.m_hi[acc >> 8] ^ t.m_lo[acc & 0xff]) ^ f; };
    unsigned long long a = (start + 7) & ~7ull;
    if (a > end) a = end;
    const uint32_t head = (uint32_t)(a - start);
    const uint32_t npairs = (uint32_t)((end - a) >> 3);
    const uint32_t tail = (uint32_t)((end - a) & 7);
    const uint2* pairs = reinterpret_cast<const uint2*>(bytes + a);
    // this lane's groups: lane, lane + 32, ...
    const uint32_t mine = npairs > lane ? (npairs - 1 - lane) / 32 + 1 : 0;
    uint32_t acc = 0;
    uint32_t i = lane;
    if (mine) {   // the first group, with the head bytes on lane 0: CRC(head || group 0) = CRC(head) * x^64 + F8(group 0)
        const uint2 v = pairs[i];
        acc = crc16_f8_le(t, v.x, v.y);
        if (lane == 0 && head) {
            uint32_t h = 0;
            for (uint32_t k = 0; k < head; k++) h = byte_step(h, bytes[start + k]);
            for (uint32_t k = 0; k < 8; k++) h = byte_step(h, 0);
            acc ^= h;
        }
        i += 32;
    }
    uint32_t left = mine ? mine - 1 : 0;
    for (; left >= 2; left -= 2, i += 64) {   // two groups per turn: the loads and the sixteen look-ups are independent
        const uint2 v = pairs[i], w = pairs[i + 32];
        const uint32_t f0 = crc16_f8_le(t, v.x, v.y), f1 = crc16_f8_le(t, w.x, w.y);
        acc = fold(fold(acc, f0), f1);
    }
    if (left) {
        const uint2 v = pairs[i];
        acc = fold(acc, crc16_f8_le(t, v.x, v.y));
    }
    const uint32_t last = lane + 32 * (mine - 1);   // (only used when mine != 0)
    uint32_t part = mine ? gf16_mulmod(acc, t.xd2[npairs - 1 - last]) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
    part &= 0xffffu;
    if (npairs == 0)   // no aligned group at all: the head bytes still have to go in
        for (uint32_t k = 0; k < head; k++) part = byte_step(part, bytes[start + k]);
    for (uint32_t k = 0; k < tail; k++) part = byte_step(part, bytes[a + (unsigned long long)npairs * 8 + k]);
    return part;
}

}   // namespace flacb200
