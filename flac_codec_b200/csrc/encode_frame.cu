// encode_frame.cu -- k_pack3: one CTA assembles one whole frame in shared memory (same shapes as the other
// register-tiled kernels: block <= 4096 samples, samples <= 28 bits, LPC order <= 16).
//
// Bit emission of encode_frame (src/encode.rs:2282-2436): frame header, then per subframe the header, warm-up samples,
// LPC parameters and the partitioned Rice residual block (:2982-3136, :3834-3863, :3944-3961), zero padding to a byte,
// CRC-16 (src/crc.rs:144-188).  One WARP per subframe: in round r lane l owns the 16-sample tile 32 r + l (history of the
// fixed differences / the FIR comes from the left neighbour by shuffle, as in k_analyze), computes its residuals and code
// lengths, a warp scan turns the lengths into bit positions, and every code is OR-ed into the frame image with one or two
// shared-memory atomics -- no per-thread word state machine, no divergent flushes.  The finished image is CRC-ed by the
// same warps (ranges of words, combined with x^(8 len) mod P) and copied out with coalesced 32-bit stores; only the up to
// three bytes at either end that share a word with the neighbouring frames are stored bytewise.  Nothing else writes the
// output, so it needs neither zeroing (k_zero) nor atomics, and no separate CRC kernel.

#include "common.cuh"
#include "crc.cuh"
#include "pack_bits.cuh"
#include "tiles.cuh"

namespace flacb200 {

bool analyze_fast_ok(const EncCfg& cfg);   // encode_kernels.cu

struct P3Smem {
    uint32_t crc_part[MAX_CH];
    FrameRec fr;
    CandRec cr[MAX_CH];
};

// the residual block of one FIXED / LPC subframe; pos = bit position of the first residual partition header
template <int HB, bool STEREO>
__device__ inline void p3_residuals(const EncCfg& cfg, const FrameDesc& d, const uint8_t* __restrict__ pcm, uint32_t slot, const CandRec& cr,
                                    uint32_t words_sa, uint32_t pos, const uint4* __restrict__ res16)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = d.n, wasted = cr.wasted, order = cr.order, shift = cr.shift;
    const bool lpc = cr.type == 3;
    const uint32_t rounds = ((n + 15) / 16 + 31) / 32;
    const uint32_t cp = n >> cr.porder_g;                      // samples per partition (the first one holds cp - order residuals)
    const uint32_t j0 = (1u << cr.porder_g) - cr.nparts;
    const bool cp16 = (cp & 15u) == 0;
    const UDiv dcp = udiv_make(cp);
    const uint32_t hb = cr.method ? 5u : 4u, escape_code = cr.method ? 31u : 15u;
    int32_t q[HB];
#pragma unroll
    for (int j = 0; j < HB; j++) q[j] = (lpc && (uint32_t)j < order) ? (int32_t)cr.q[j] : 0;
    int32_t carry[16];
#pragma unroll
    for (int e = 0; e < 16; e++) carry[e] = 0;
    uint32_t base = pos;
    for (uint32_t rd = 0; rd < rounds; rd++) {
        const uint32_t i0 = (rd * 32 + lane) * 16;
        const bool live = i0 < n;
        int32_t r[16];
        if (res16 != nullptr) {   // the residuals k_analyze3 left behind (int16 pairs, two 16-byte chunks per tile)
            uint4 lo4 = make_uint4(0, 0, 0, 0), hi4 = lo4;
            if (live) {
                lo4 = res16[(rd * 2 + 0) * 32 + lane];
                hi4 = res16[(rd * 2 + 1) * 32 + lane];
            }
            const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                r[2 * k] = (int32_t)(w[k] << 16) >> 16;
                r[2 * k + 1] = (int32_t)w[k] >> 16;
            }
        } else {
        int32_t x[16], h[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = 0;
        if (live) aw_load_tile<STEREO>(cfg, d, pcm, slot, i0, x);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            x[e] >>= wasted;   // :2891
            const int32_t up = __shfl_up_sync(0xffffffffu, x[e], 1);
            h[e] = lane == 0 ? carry[e] : up;
            carry[e] = __shfl_sync(0xffffffffu, x[e], 31);
        }
        if (lpc) {   // LpcSubframeParameters::encode_residuals (:3174-3203)
#pragma unroll
            for (int e = 0; e < 16; e++) {
                long long sum = 0;
#pragma unroll
                for (int j = 0; j < HB; j++) sum = mad_wide_s32(e - 1 - j >= 0 ? x[e - 1 - j >= 0 ? e - 1 - j : 0] : h[16 + e - 1 - j >= 0 ? 16 + e - 1 - j : 0], q[j], sum);
                r[e] = (int32_t)((uint32_t)x[e] - (uint32_t)(unsigned long long)(sum >> shift));
            }
        } else {     // fixed differences (:3039-3060)
            int32_t x1 = h[15], x2 = h[14], x3 = h[13], x4 = h[12];
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int32_t x0 = x[e];
                r[e] = order == 0 ? x0 : order == 1 ? x0 - x1 : order == 2 ? x0 - 2 * x1 + x2 : order == 3 ? x0 - 3 * x1 + 3 * x2 - x3
                                                                                                           : x0 - 4 * x1 + 6 * x2 - 4 * x3 + x4;
                x4 = x3; x3 = x2; x2 = x1; x1 = x0;
            }
        }
        }
        // ---- code lengths of this lane's tile ----
        const uint32_t lo_i = max(i0, order), hi_i = min(i0 + 16u, n);   // residuals exist for [lo_i, hi_i)
        const uint32_t pj = live ? udiv(i0, dcp) : 0u;
        // the common tile: one Rice parameter for all its residuals (:3845-3851).  The block's first tile belongs here too --
        // its first `order` samples are warm-up and carry no code (skip) -- or its lane would hold the warp in the general
        // path for the length of five ordinary rounds.
        const bool uniform = live && (i0 >= order || (i0 == 0 && order <= 16)) && i0 + 16 <= n && (cp16 || udiv(i0 + 15, dcp) == pj);
        const uint32_t cc0 = cr.rice[live ? min(pj - j0, (uint32_t)MAX_PARTS - 1) : 0u];
        const uint32_t first_res = max(pj * cp, order);                         // first residual of the tile's partition
        const uint32_t skip = first_res > i0 ? min(first_res - i0, 16u) : 0u;   // nonzero only in the first tile
        const bool hdr_here = first_res >= i0 && first_res < i0 + 16;          // ResidualPartitionHeader (src/stream.rs:1603-1619) rides in front
        uint32_t tsum = 0;
        uint32_t len[16];
        if (uniform && cc0 < 0x40) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                len[e] = (uint32_t)e >= skip ? (zigzag32(r[e]) >> cc0) + 1u + cc0 : 0u;
                tsum += len[e];
            }
            if (hdr_here) tsum += hb;
        } else if (live) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const uint32_t i = i0 + e;
                uint32_t l = 0;
                if (i >= lo_i && i < hi_i) {
                    const uint32_t p = udiv(i, dcp);
                    const uint32_t cc = cr.rice[p - j0];
                    if (cc < 0x40) l = (zigzag32(r[e]) >> cc) + 1u + cc;
                    else if (cc & 0x40) l = cc & 31u;
                    if (i == max(p * cp, order)) l += (cc < 0x40) ? hb : hb + 5u;
                }
                len[e] = l;
                tsum += l;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) len[e] = 0;
        }
        // ---- bit positions: exclusive scan over the lanes, running base over the rounds ----
        uint32_t incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        uint32_t p = base + incl - tsum;
        base += __shfl_sync(0xffffffffu, incl, 31);
        // ---- emission ----
        if (uniform && cc0 < 0x40) {
            if (hdr_here) {
                p3_put(words_sa, p, hb, cc0);
                p += hb;
            }
            const uint32_t stop = 1u << cc0, mask = stop - 1u;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const uint32_t u = zigzag32(r[e]);
                // unary zeros, stop bit, cc0 LSBs; a warm-up sample ORs nothing (its length is 0: p stays)
                p3_put(words_sa, p + ((uint32_t)e >= skip ? u >> cc0 : 0u), cc0 + 1u, (uint32_t)e >= skip ? stop | (u & mask) : 0u);
                p += len[e];
            }
        } else if (live) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const uint32_t i = i0 + e;
                if (i < lo_i || i >= hi_i) continue;
                const uint32_t pi = udiv(i, dcp);
                const uint32_t cc = cr.rice[pi - j0];
                uint32_t at = p;
                if (i == max(pi * cp, order)) {
                    if (cc < 0x40) { p3_put(words_sa, at, hb, cc); at += hb; }
                    else { p3_put(words_sa, at, hb, escape_code); p3_put_masked(words_sa, at + hb, 5, (cc & 0x40) ? (cc & 31u) : 0u); at += hb + 5; }
                }
                if (cc < 0x40) {
                    const uint32_t u = zigzag32(r[e]);
                    p3_put(words_sa, at + (u >> cc), cc + 1u, (1u << cc) | (u & ((1u << cc) - 1u)));
                } else if (cc & 0x40) {
                    p3_put_masked(words_sa, at, cc & 31u, (uint32_t)r[e]);   // escaped: raw two's complement (:3857)
                }
                p += len[e];
            }
        }
    }
}

// grid = frames of the launch group; block = 32 * (subframes per frame); dynamic smem = image words
// Two launches share the work: the first with a shared-memory image sized for ordinary frames (min_words = 0, cap_words =
// P3_SMALL_WORDS: more CTAs per SM), the second with room for the worst case, for the few frames that need more
// (min_words = P3_SMALL_WORDS); a CTA whose frame belongs to the other launch exits at once.
#ifndef FLACB200_P3_PF
#define FLACB200_P3_PF 1
#endif
constexpr uint32_t P3_PF_DIST = 1480;   // frames ahead: the CTAs resident at once (10 x 148 SMs)
template <int HB, bool STEREO>
#ifndef FLACB200_P3_MINB
#define FLACB200_P3_MINB 9
#endif
#ifndef FLACB200_P3_SMALL
#define FLACB200_P3_SMALL 6
#endif
__global__ void __launch_bounds__(STEREO ? 64 : 256, STEREO ? FLACB200_P3_MINB : 2)
    k_pack3(EncCfg cfg, uint32_t min_words, uint32_t cap_words, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
            const CandRec* __restrict__ cands, const FrameRec* __restrict__ frecs, uint8_t* __restrict__ out, const uint4* __restrict__ gres16)
{
    extern __shared__ __align__(16) uint32_t p3_words[];
    __shared__ P3Smem sm;
    const uint32_t f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const uint32_t words_sa = (uint32_t)__cvta_generic_to_shared(p3_words);
    {
        const uint32_t need = (frecs[f].frame_bytes + 3) / 4 + 2;
        if (need <= min_words || (need > cap_words && min_words == 0)) return;   // the other launch's frame
    }
    // ---- stage the frame record, its subframes' candidate records and the CRC tables; clear the image ----
    for (uint32_t i = tid; i < sizeof(FrameRec) / 4; i += nthreads) reinterpret_cast<uint32_t*>(&sm.fr)[i] = reinterpret_cast<const uint32_t*>(frecs + f)[i];
    __syncthreads();
    const FrameRec& fr = sm.fr;
    const uint32_t nsub = fr.nsub;
    const uint32_t frame_bytes = fr.frame_bytes;
    const uint32_t img_words = min((frame_bytes + 3) / 4 + 2, cap_words);
    for (uint32_t i = tid; i < img_words / 4; i += nthreads) reinterpret_cast<uint4*>(p3_words)[i] = make_uint4(0, 0, 0, 0);
    if (tid < (img_words & 3u)) p3_words[(img_words & ~3u) + tid] = 0;
    for (uint32_t c = wid; c < nsub; c += nwarps) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(cands + (size_t)f * cfg.nslots + fr.slot[c]);
        for (uint32_t i = lane; i < sizeof(CandRec) / 4; i += 32) reinterpret_cast<uint32_t*>(&sm.cr[c])[i] = src[i];
    }
    __syncthreads();
    if (fr.err || (frame_bytes + 3) / 4 + 2 > cap_words) return;   // k_scan has already raised the sticky error word
    const FrameDesc d = descs[f];
    const uint32_t n = d.n;
    // ---- frame header (src/stream.rs:242-276; bytes prepared by k_decide) ----
    if (tid < fr.hdr_len) p3_put(words_sa, 8 * tid, 8, fr.hdr[tid]);
    // ---- subframes: one warp each ----
    for (uint32_t c = wid; c < nsub; c += nwarps) {
        const CandRec& cr = sm.cr[c];
        const uint32_t slot = fr.slot[c];
        const uint32_t wasted = cr.wasted, bps = cr.bps, type = cr.type, order = type >= 2 ? cr.order : 0u;
        uint32_t pos = fr.sub_bit[c];
        if (lane == 0) {   // SubframeHeader (src/stream.rs:1397-1413): pad, 6-bit type, wasted flag, unary(wasted - 1)
            const uint32_t code = type == 0 ? 0u : type == 1 ? 1u : type == 2 ? 8u + order : 31u + order;
            p3_put_masked(words_sa, pos, 8, (code << 1) | (wasted ? 1u : 0u));
            if (wasted) p3_put(words_sa, pos + 8 + (wasted - 1), 1, 1);
        }
        pos += 8 + wasted;
        if (type == 0) {   // CONSTANT (:2982-2998): the first sample
            if (lane == 0) {
                int32_t x[16];
                aw_load_tile<STEREO>(cfg, d, pcm, slot, 0, x);
                p3_put_masked(words_sa, pos, bps, (uint32_t)(x[0] >> wasted));
            }
            continue;
        }
        if (type == 1) {   // VERBATIM (:3000-3018)
            for (uint32_t i0 = lane * 16; i0 < n; i0 += 32 * 16) {
                int32_t x[16];
                aw_load_tile<STEREO>(cfg, d, pcm, slot, i0, x);
#pragma unroll
                for (int e = 0; e < 16; e++)
                    if (i0 + e < n) p3_put_masked(words_sa, pos + (i0 + e) * bps, bps, (uint32_t)(x[e] >> wasted));
            }
            continue;
        }
        if (lane == 0) {   // warm-up samples (:3083, :3118); order <= 16 = one tile
            int32_t x[16];
            aw_load_tile<STEREO>(cfg, d, pcm, slot, 0, x);
#pragma unroll
            for (int e = 0; e < 16; e++)
                if ((uint32_t)e < order) p3_put_masked(words_sa, pos + e * bps, bps, (uint32_t)(x[e] >> wasted));
        }
        pos += order * bps;
        if (type == 3) {   // :3122-3133
            const uint32_t prec = cr.precision;
            if (lane == 0) {
                p3_put_masked(words_sa, pos, 4, prec - 1);
                p3_put_masked(words_sa, pos + 4, 5, cr.shift);
            }
            if (lane < order) p3_put_masked(words_sa, pos + 9 + lane * prec, prec, (uint32_t)(int32_t)cr.q[lane]);
            pos += 9 + order * prec;
        }
        if (lane == 0) {   // residual block header (:3944-3961)
            p3_put_masked(words_sa, pos, 2, cr.method);
            p3_put_masked(words_sa, pos + 2, 4, cr.porder_w);
        }
        pos += 6;
        // (CandRec::pad0: k_analyze3 stored this LPC subframe's residuals; 512 16-byte chunks per candidate)
        const uint4* r16 = (gres16 != nullptr && cr.type == 3 && cr.pad0 == 1) ? gres16 + ((size_t)f * cfg.nslots + slot) * 512 : nullptr;
        p3_residuals<HB, STEREO>(cfg, d, pcm, slot, cr, words_sa, pos, r16);
    }
#if FLACB200_P3_PF
    // the int16 residuals of the frame that will be packed in this CTA's place (ten CTAs per SM) are asked into L2 now: they were
    // written by k_analyze3 a gigabyte ago and are this kernel's only DRAM read
    if (STEREO && gres16 != nullptr && f + P3_PF_DIST < cfg.nframes) {
        const FrameRec* fnext = frecs + f + P3_PF_DIST;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint4* q = gres16 + ((size_t)(f + P3_PF_DIST) * cfg.nslots + fnext->slot[c]) * 512 + tid * 8;
            if (tid < 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
    }
#endif
    __syncthreads();
    // ---- CRC-16 over everything but the last two bytes (src/encode.rs:2408-2409) ----
    const uint32_t body = frame_bytes - 2;
    const uint32_t bw = body >> 2, btail = body & 3;
    {
        // warp w takes words [w * per, (w + 1) * per), per a multiple of 64 (whole rounds of 32 pairs)
        const uint32_t per = (((bw + nwarps - 1) / nwarps) + 63u) & ~63u;
        const uint32_t a = min(wid * per, bw), b = min(a + per, bw);
        const uint32_t part = p3_crc_words(g_crc16_tabs, p3_words, a, b - a);
        if (lane == 0) sm.crc_part[wid] = part;
        __syncthreads();
        if (tid == 0) {
            uint32_t crc = 0;
            for (uint32_t w = 0; w < nwarps; w++) {
                const uint32_t wa = min(w * per, bw), wb = min(wa + per, bw);
                if (wb > wa) {   // ranges are whole blocks of 32 words except the last: shift by the table, then by the odd words
                    const uint32_t len = wb - wa;
                    crc = gf16_mulmod(crc, g_crc16_xblk[len >> 5]);
                    if (len & 31u) crc = gf16_mulmod(crc, g_crc16_tabs.xd[len & 31u]);
                    crc ^= sm.crc_part[w];
                }
            }
            for (uint32_t t = 0; t < btail; t++) crc = (g_crc16_tabs.T[0][((crc >> 8) ^ (p3_words[bw] >> (24 - 8 * t))) & 0xff] ^ (crc << 8)) & 0xffffu;
            p3_put(words_sa, body * 8, 16, crc);
        }
        __syncthreads();
    }
    // ---- copy out: bytes up to the first 16-byte boundary of the output, 128-bit stores, the rest by bytes ----
    const unsigned long long o = fr.out_off;
    const unsigned long long A = (o + 15) & ~15ull, B = (o + frame_bytes) & ~15ull;
    auto frame_byte = [&](uint32_t b) -> uint8_t { return (uint8_t)(p3_words[b >> 2] >> (24 - 8 * (b & 3))); };
    if (A >= B) {
        for (uint32_t b = tid; b < frame_bytes; b += nthreads) out[o + b] = frame_byte(b);
        return;
    }
    const uint32_t head = (uint32_t)(A - o), tail0 = (uint32_t)(B - o);
    if (tid < head) out[o + tid] = frame_byte(tid);
    if (tid < frame_bytes - tail0) out[B + tid] = frame_byte(tail0 + tid);
    uint4* gq = reinterpret_cast<uint4*>(out + A);
    const uint32_t nq = (uint32_t)((B - A) >> 4);
    const uint32_t sh = 8 * (head & 3);   // the image's words are big-endian bit order; a 16-byte run starts `head` bytes into it
    for (uint32_t j = tid; j < nq; j += nthreads) {
        const uint32_t idx = (head >> 2) + 4 * j;
        const uint32_t w0 = p3_words[idx], w1 = p3_words[idx + 1], w2 = p3_words[idx + 2], w3 = p3_words[idx + 3];
        uint4 v;
        if (sh) {
            const uint32_t w4 = p3_words[idx + 4];   // (within the image: at least one byte of the frame follows this run's last word)
            v.x = __byte_perm(__funnelshift_l(w1, w0, sh), 0, 0x0123);
            v.y = __byte_perm(__funnelshift_l(w2, w1, sh), 0, 0x0123);
            v.z = __byte_perm(__funnelshift_l(w3, w2, sh), 0, 0x0123);
            v.w = __byte_perm(__funnelshift_l(w4, w3, sh), 0, 0x0123);
        } else {
            v.x = __byte_perm(w0, 0, 0x0123); v.y = __byte_perm(w1, 0, 0x0123); v.z = __byte_perm(w2, 0, 0x0123); v.w = __byte_perm(w3, 0, 0x0123);
        }
        gq[j] = v;
    }
}

// once per device, from flacb200_engine_create (synchronised there: no kernel can see half-built tables)
void init_encode_tables(cudaStream_t st) { k_crc16_tables_init<<<1, 256, 0, st>>>(); }

uint32_t pack3_cap_words(const EncCfg& cfg)
{
    const uint32_t nsub = cfg.mode == MODE_INDEPENDENT ? cfg.channels : 2;
    // header <= 16 bytes; a subframe is never larger than VERBATIM + 1 bit (8 + wasted + n * bps bits, src/encode.rs:2971); CRC-16
    const size_t bits = 16 * 8 + (size_t)nsub * (40 + (size_t)cfg.block_size * (cfg.bps + 1)) + 16;
    return (uint32_t)((bits + 31) / 32 + 4);
}

bool pack3_ok(const EncCfg& cfg) { return analyze_fast_ok(cfg) && (size_t)pack3_cap_words(cfg) * 4 <= 200 * 1024; }

cudaError_t launch_pack3(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const CandRec* cands, const FrameRec* frecs, uint8_t* out,
                         const uint4* gres16, cudaStream_t st)
{
    const uint32_t nsub = cfg.mode == MODE_INDEPENDENT ? cfg.channels : 2;
    const uint32_t cap_words = pack3_cap_words(cfg);
    // ordinary frames compress to well under 3/4 of the raw size: a smaller image lets more CTAs share an SM
    const uint32_t small_words = (cap_words * FLACB200_P3_SMALL / 8u) & ~1u;
    const uint32_t hb = cfg.max_lpc_order ? (cfg.max_lpc_order + 3u) >> 2 : 1u;
#define FLACB200_P3(HBV, ST)                                                                                                         \
    do {                                                                                                                             \
        cudaError_t e_ = cudaFuncSetAttribute(k_pack3<HBV, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);            \
        if (e_ != cudaSuccess) return e_;                                                                                            \
        count_launch(), k_pack3<HBV, ST><<<cfg.nframes, 32 * nsub, (size_t)small_words * 4, st>>>(cfg, 0u, small_words, descs, pcm, cands, frecs, out, gres16);    \
        count_launch(), k_pack3<HBV, ST><<<cfg.nframes, 32 * nsub, (size_t)cap_words * 4, st>>>(cfg, small_words, cap_words, descs, pcm, cands, frecs, out, gres16); \
    } while (0)
    if (cfg.mode != MODE_INDEPENDENT) {
        switch (hb) {
        case 1: FLACB200_P3(4, true); break;
        case 2: FLACB200_P3(8, true); break;
        case 3: FLACB200_P3(12, true); break;
        default: FLACB200_P3(16, true); break;
        }
    } else {
        switch (hb) {
        case 1: FLACB200_P3(4, false); break;
        case 2: FLACB200_P3(8, false); break;
        case 3: FLACB200_P3(12, false); break;
        default: FLACB200_P3(16, false); break;
        }
    }
#undef FLACB200_P3
    return cudaGetLastError();
}

}   // namespace flacb200
