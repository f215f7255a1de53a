// md5_mb.cpp -- see md5_mb.h
#include "md5_mb.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define FLACB200_HAVE_X86 1
#endif

namespace flacb200 {
namespace {

const uint32_t K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
    0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
    0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
    0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
    0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
    0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
const uint8_t S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                       4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

inline uint32_t msg_index(int i) { return i < 16 ? i : i < 32 ? (5 * i + 1) & 15 : i < 48 ? (3 * i + 5) & 15 : (7 * i) & 15; }

struct State {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
};

void block_scalar(State& s, const uint8_t* p)
{
    uint32_t m[16];
    memcpy(m, p, 64);
    uint32_t a = s.a, b = s.b, c = s.c, d = s.d;
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        if (i < 16) f = d ^ (b & (c ^ d));
        else if (i < 32) f = c ^ (d & (b ^ c));
        else if (i < 48) f = b ^ c ^ d;
        else f = c ^ (b | ~d);
        const uint32_t t = a + f + K[i] + m[msg_index(i)];
        a = d; d = c; c = b;
        b = b + ((t << S[i]) | (t >> (32 - S[i])));
    }
    s.a += a; s.b += b; s.c += c; s.d += d;
}

// the tail of one stream: remaining bytes, padding, length
void finish(State s, const uint8_t* rest, size_t nrest, uint64_t total, uint8_t out[16])
{
    while (nrest >= 64) {
        block_scalar(s, rest);
        rest += 64;
        nrest -= 64;
    }
    uint8_t buf[128] = {0};
    memcpy(buf, rest, nrest);
    buf[nrest] = 0x80;
    const size_t padded = nrest < 56 ? 64 : 128;
    const uint64_t bits = total * 8;
    for (int i = 0; i < 8; i++) buf[padded - 8 + i] = (uint8_t)(bits >> (8 * i));
    block_scalar(s, buf);
    if (padded == 128) block_scalar(s, buf + 64);
    const uint32_t v[4] = {s.a, s.b, s.c, s.d};
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(v[i] >> (8 * k));
}

#ifdef FLACB200_HAVE_X86
// eight streams, `nblocks` 64-byte blocks each, in lock step
__attribute__((target("avx2"))) void blocks_avx2(State st[8], const uint8_t* const p[8], size_t nblocks)
{
    alignas(32) uint32_t ta[8], tb[8], tc[8], td[8];
    for (int l = 0; l < 8; l++) { ta[l] = st[l].a; tb[l] = st[l].b; tc[l] = st[l].c; td[l] = st[l].d; }
    __m256i A = _mm256_load_si256((const __m256i*)ta), B = _mm256_load_si256((const __m256i*)tb);
    __m256i C = _mm256_load_si256((const __m256i*)tc), D = _mm256_load_si256((const __m256i*)td);
    for (size_t blk = 0; blk < nblocks; blk++) {
        // message words: 8 streams x 16 words -> 16 registers of 8 lanes (two 8x8 transposes)
        __m256i M[16];
        for (int half = 0; half < 2; half++) {
            __m256i r[8];
            for (int l = 0; l < 8; l++) r[l] = _mm256_loadu_si256((const __m256i*)(p[l] + blk * 64 + half * 32));
            const __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);
            const __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);
            const __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);
            const __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);
            const __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
            const __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
            const __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
            const __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
            M[half * 8 + 0] = _mm256_permute2x128_si256(u0, u4, 0x20);
            M[half * 8 + 1] = _mm256_permute2x128_si256(u1, u5, 0x20);
            M[half * 8 + 2] = _mm256_permute2x128_si256(u2, u6, 0x20);
            M[half * 8 + 3] = _mm256_permute2x128_si256(u3, u7, 0x20);
            M[half * 8 + 4] = _mm256_permute2x128_si256(u0, u4, 0x31);
            M[half * 8 + 5] = _mm256_permute2x128_si256(u1, u5, 0x31);
            M[half * 8 + 6] = _mm256_permute2x128_si256(u2, u6, 0x31);
            M[half * 8 + 7] = _mm256_permute2x128_si256(u3, u7, 0x31);
        }
        __m256i a = A, b = B, c = C, d = D;
        for (int i = 0; i < 64; i++) {   // (unrolling changes nothing: the step is one dependent chain a -> t -> rotate -> b)
            __m256i f;
            if (i < 16) f = _mm256_xor_si256(d, _mm256_and_si256(b, _mm256_xor_si256(c, d)));
            else if (i < 32) f = _mm256_xor_si256(c, _mm256_and_si256(d, _mm256_xor_si256(b, c)));
            else if (i < 48) f = _mm256_xor_si256(_mm256_xor_si256(b, c), d);
            else f = _mm256_xor_si256(c, _mm256_or_si256(b, _mm256_xor_si256(d, _mm256_set1_epi32(-1))));
            __m256i t = _mm256_add_epi32(_mm256_add_epi32(a, f), _mm256_add_epi32(_mm256_set1_epi32((int)K[i]), M[msg_index(i)]));
            a = d; d = c; c = b;
            b = _mm256_add_epi32(b, _mm256_or_si256(_mm256_slli_epi32(t, S[i]), _mm256_srli_epi32(t, 32 - S[i])));
        }
        A = _mm256_add_epi32(A, a); B = _mm256_add_epi32(B, b); C = _mm256_add_epi32(C, c); D = _mm256_add_epi32(D, d);
    }
    _mm256_store_si256((__m256i*)ta, A); _mm256_store_si256((__m256i*)tb, B);
    _mm256_store_si256((__m256i*)tc, C); _mm256_store_si256((__m256i*)td, D);
    for (int l = 0; l < 8; l++) { st[l].a = ta[l]; st[l].b = tb[l]; st[l].c = tc[l]; st[l].d = td[l]; }
}
#endif

// sixteen streams in the lanes of AVX-512 registers: native rotate, one ternary-logic instruction per round function
__attribute__((target("avx512f"))) void blocks_avx512(State st[16], const uint8_t* const p[16], size_t nblocks)
{
    alignas(64) uint32_t ta[16], tb[16], tc[16], td[16];
    for (int l = 0; l < 16; l++) { ta[l] = st[l].a; tb[l] = st[l].b; tc[l] = st[l].c; td[l] = st[l].d; }
    __m512i A = _mm512_load_si512(ta), B = _mm512_load_si512(tb), C = _mm512_load_si512(tc), D = _mm512_load_si512(td);
    for (size_t blk = 0; blk < nblocks; blk++) {
        // 16 streams x 16 message words -> 16 registers of 16 lanes (a 16 x 16 transpose of 32-bit words)
        __m512i r[16], t[16];
        for (int l = 0; l < 16; l++) r[l] = _mm512_loadu_si512(p[l] + blk * 64);
        for (int i = 0; i < 16; i += 2) {
            t[i] = _mm512_unpacklo_epi32(r[i], r[i + 1]);
            t[i + 1] = _mm512_unpackhi_epi32(r[i], r[i + 1]);
        }
        for (int i = 0; i < 16; i += 4) {
            r[i] = _mm512_unpacklo_epi64(t[i], t[i + 2]);
            r[i + 1] = _mm512_unpackhi_epi64(t[i], t[i + 2]);
            r[i + 2] = _mm512_unpacklo_epi64(t[i + 1], t[i + 3]);
            r[i + 3] = _mm512_unpackhi_epi64(t[i + 1], t[i + 3]);
        }
        // r[4 g + k] holds, per 128-bit lane q, word 4 q + k of streams 4 g .. 4 g + 3
        for (int k = 0; k < 4; k++) {
            t[k] = _mm512_shuffle_i32x4(r[k], r[4 + k], 0x88);        // lanes 0, 2 of streams 0-3 | 4-7
            t[4 + k] = _mm512_shuffle_i32x4(r[k], r[4 + k], 0xdd);    // lanes 1, 3
            t[8 + k] = _mm512_shuffle_i32x4(r[8 + k], r[12 + k], 0x88);
            t[12 + k] = _mm512_shuffle_i32x4(r[8 + k], r[12 + k], 0xdd);
        }
        __m512i M[16];
        for (int k = 0; k < 4; k++) {
            M[k] = _mm512_shuffle_i32x4(t[k], t[8 + k], 0x88);            // word k      (128-bit lane 0 of every stream group)
            M[8 + k] = _mm512_shuffle_i32x4(t[k], t[8 + k], 0xdd);        // word 8 + k  (lane 2)
            M[4 + k] = _mm512_shuffle_i32x4(t[4 + k], t[12 + k], 0x88);   // word 4 + k  (lane 1)
            M[12 + k] = _mm512_shuffle_i32x4(t[4 + k], t[12 + k], 0xdd);  // word 12 + k (lane 3)
        }
        __m512i a = A, b = B, c = C, d = D;
        for (int i = 0; i < 64; i++) {
            __m512i f;
            if (i < 16) f = _mm512_ternarylogic_epi32(b, c, d, 0xCA);        // (b & c) | (~b & d)
            else if (i < 32) f = _mm512_ternarylogic_epi32(b, c, d, 0xE4);   // (b & d) | (c & ~d)
            else if (i < 48) f = _mm512_ternarylogic_epi32(b, c, d, 0x96);   // b ^ c ^ d
            else f = _mm512_ternarylogic_epi32(b, c, d, 0x39);               // c ^ (b | ~d)
            const __m512i tt = _mm512_add_epi32(_mm512_add_epi32(a, f), _mm512_add_epi32(_mm512_set1_epi32((int)K[i]), M[msg_index(i)]));
            a = d; d = c; c = b;
            b = _mm512_add_epi32(b, _mm512_rolv_epi32(tt, _mm512_set1_epi32(S[i])));
        }
        A = _mm512_add_epi32(A, a); B = _mm512_add_epi32(B, b); C = _mm512_add_epi32(C, c); D = _mm512_add_epi32(D, d);
    }
    _mm512_store_si512(ta, A); _mm512_store_si512(tb, B); _mm512_store_si512(tc, C); _mm512_store_si512(td, D);
    for (int l = 0; l < 16; l++) { st[l].a = ta[l]; st[l].b = tb[l]; st[l].c = tc[l]; st[l].d = td[l]; }
}

bool have_avx512()
{
    static const bool v = __builtin_cpu_supports("avx512f");
    return v;
}

bool have_avx2()
{
#ifdef FLACB200_HAVE_X86
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
#else
    return false;
#endif
}

// one group of up to sixteen (AVX-512) / eight (AVX2) streams
void md5_group(const uint8_t* const* data, const size_t* len, const size_t* idx, size_t n, uint8_t (*digests)[16])
{
    State st[16];
    size_t done = 0;
#ifdef FLACB200_HAVE_X86
    if (n > 8 && have_avx512()) {
        size_t common = ~(size_t)0;
        for (size_t l = 0; l < n; l++) common = std::min(common, len[idx[l]] / 64);
        const uint8_t* p[16];
        for (size_t l = 0; l < 16; l++) p[l] = data[idx[l < n ? l : 0]];
        const size_t CH = 4096;
        for (size_t b0 = 0; b0 < common; b0 += CH) {
            const uint8_t* q[16];
            for (int l = 0; l < 16; l++) q[l] = p[l] + b0 * 64;
            blocks_avx512(st, q, std::min(CH, common - b0));
        }
        done = common * 64;
    } else if (n > 1 && have_avx2()) {
        size_t common = ~(size_t)0;
        for (size_t l = 0; l < n; l++) common = std::min(common, len[idx[l]] / 64);
        const uint8_t* p[8];
        for (size_t l = 0; l < 8; l++) p[l] = data[idx[l < n ? l : 0]];   // idle lanes repeat the first stream
        const size_t CH = 4096;   // blocks per call: the lanes walk their streams in step, a few pages at a time
        for (size_t b0 = 0; b0 < common; b0 += CH) {
            const uint8_t* q[8];
            for (int l = 0; l < 8; l++) q[l] = p[l] + b0 * 64;
            blocks_avx2(st, q, std::min(CH, common - b0));
        }
        done = common * 64;
    }
#endif
    for (size_t l = 0; l < n; l++) {
        const size_t i = idx[l];
        finish(st[l], data[i] + done, len[i] - done, len[i], digests[i]);
    }
}

}   // namespace

void md5_many(const uint8_t* const* data, const size_t* len, size_t n, uint8_t (*digests)[16], unsigned threads)
{
    if (n == 0) return;
    // streams of similar length share a group: the lock-step part is the shortest stream of the group
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return len[a] > len[b]; });
    // enough groups for every thread: groups narrower than eight when streams are few (a lane costs nothing, a thread does)
    threads = std::max(1u, threads);
    size_t width = 8;
#ifdef FLACB200_HAVE_X86
    if (have_avx512()) width = 16;
#endif
    while (width > 1 && (n + width - 1) / width < threads) width /= 2;
    const size_t ngroups = (n + width - 1) / width;
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (;;) {
            const size_t g = next.fetch_add(1);
            if (g >= ngroups) return;
            const size_t a = g * width, b = std::min(n, a + width);
            md5_group(data, len, order.data() + a, b - a, digests);
        }
    };
    const unsigned nt = (unsigned)std::min<size_t>(threads, ngroups);
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

}   // namespace flacb200
