// common.cuh -- shared device structures and helpers of the B200 FLAC frame engine.
//
// Data layout in HBM (per launch group of F frames, S candidate slots per frame):
//   pcm        packed interleaved PCM as the caller supplied it (read by k_planes only)
//   planes     int32 [F][S][bpad]   candidate channels: stereo -> L, R, M=(L+R)>>1, S=L-R; else one per channel
//   ormask     u32   [F][S]         OR of all samples of a candidate (wasted bits / all-zero detection)
//   abssum     u64   [F][4]         sum |x| per stereo candidate (only for non-exhaustive channel correlation)
//   lpc        LpcRec[F][S]         quantised LPC parameters chosen by k_lpc
//   cand       CandRec[F][S]        the winning subframe encoding of each candidate and its exact bit size
//   frec       FrameRec[F]          channel assignment, header bytes, frame size, subframe bit offsets
//   out        bytes                frames back to back at frec[f].out_off
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace flacb200 {

// every kernel launch of the library passes through here: flacb200_timings::launches is a count, not an estimate
// (per host thread: an engine call runs on its caller's thread, so the difference across a call is that call's count)
extern thread_local unsigned long long g_kernel_launches;
inline void count_launch() { g_kernel_launches++; }

constexpr int MAX_LPC = 32;
constexpr int MAX_PARTS = 64;   // MAX_PARTITIONS, src/encode.rs:3756
constexpr int MAX_PORDER = 6;   // the reference overflows its 64-entry ArrayVec above this
constexpr int MAX_CH = 8;

// stereo handling modes
enum : uint32_t {
    MODE_INDEPENDENT = 0,     // mono / >2 channels / 32-bit stereo: one candidate per channel
    MODE_EXH_MID_SIDE = 1,    // correlate_channels_exhaustive with mid_side: L R M S all encoded   (:2716)
    MODE_EXH_SIDE = 2,        // correlate_channels_exhaustive without mid_side: L R S              (:2788)
    MODE_FAST_MID_SIDE = 3,   // correlate_channels (abs-sum heuristic) with mid_side               (:2474)
    MODE_FAST_SIDE = 4,       // correlate_channels without mid_side                                (:2582)
};

struct EncCfg {
    uint32_t channels, bps, sample_rate, subset;
    uint32_t block_size;      // nominal block size of the launch group
    uint32_t bpad;            // plane stride in samples (multiple of 32)
    uint32_t nslots;          // candidate slots per frame
    uint32_t mode;
    uint32_t max_lpc_order, max_porder, use_rice2;
    uint32_t pcm_kind, bytes_per_sample;
    unsigned long long planar_stride;
    uint32_t nframes;
};

struct FrameDesc {
    unsigned long long pcm_off;   // first inter-channel sample of the block in the pcm buffer
    unsigned long long fnum;      // FrameNumber
    uint32_t n;                   // block length in samples
    uint32_t win_off;             // offset (in doubles) of this length's window in the window pool
};

struct LpcRec {
    int32_t ok;                   // 0: LPC unavailable (reference returned Err), 1: parameters valid
    uint8_t order, precision, shift, pad;
    int16_t q[MAX_LPC];
};

// rice[] encoding: 0..30 standard parameter, 0x40|w escaped with w raw bits, 0x80 all-zero partition
struct CandRec {
    uint32_t bits;                // exact subframe size in bits (BitRecorder::written())
    uint8_t type;                 // 0 CONSTANT 1 VERBATIM 2 FIXED 3 LPC, 0xFF error
    uint8_t order, wasted, bps;   // bps = effective subframe sample width
    uint8_t precision, shift, method, porder_w;   // porder_w: partition order as written (ilog2(count))
    uint8_t porder_g, nparts, pad0, pad1;          // porder_g: geometry (chunk = n >> porder_g)
    int16_t q[MAX_LPC];
    uint8_t rice[MAX_PARTS];
};

struct FrameRec {
    unsigned long long out_off;   // byte offset of the frame in the output
    uint32_t frame_bytes;
    uint32_t err;
    uint32_t sub_bit[MAX_CH];     // bit offset of each subframe from the frame start
    uint8_t slot[MAX_CH];         // candidate slot used for each subframe
    uint8_t assignment, hdr_len, nsub, pad;
    uint8_t hdr[20];
};

// ---- candidate geometry ------------------------------------------------------------------
__host__ __device__ inline uint32_t cand_bps(const EncCfg& c, uint32_t slot)
{
    return (c.mode != MODE_INDEPENDENT && slot == 3) ? c.bps + 1 : c.bps;   // side channel: bps + 1 (:2715)
}

// correlate_channels (src/encode.rs:2463-2674): assignment from the abs sums, first minimum wins.
// Returns the frame-header channel assignment code (1 independent, 8 L/S, 9 S/R, 10 M/S).
__host__ __device__ inline uint32_t fast_assignment(const unsigned long long* s, bool mid_side)
{
    unsigned long long l = s[0], r = s[1], m = s[2], d = s[3];
    if (mid_side) {   // [Independent, LeftSide, SideRight, MidSide]  :2506-2517
        unsigned long long t[4] = {l + r, l + d, d + r, m + d};
        int b = 0;
        for (int k = 1; k < 4; k++) if (t[k] < t[b]) b = k;
        return b == 0 ? 1u : (b == 1 ? 8u : (b == 2 ? 9u : 10u));
    }
    // [LeftSide, SideRight, Independent]  :2600-2607
    unsigned long long t[3] = {l + d, d + r, l + r};
    int b = 0;
    for (int k = 1; k < 3; k++) if (t[k] < t[b]) b = k;
    return b == 0 ? 8u : (b == 1 ? 9u : 1u);
}

__host__ __device__ inline void assignment_slots(uint32_t assignment, uint8_t* s0, uint8_t* s1)
{
    switch (assignment) {
    case 8: *s0 = 0; *s1 = 3; break;    // left, side
    case 9: *s0 = 3; *s1 = 1; break;    // side, right
    case 10: *s0 = 2; *s1 = 3; break;   // mid, side
    default: *s0 = 0; *s1 = 1; break;
    }
}

// is candidate `slot` of a frame encoded at all?
__device__ inline bool slot_active(const EncCfg& c, const unsigned long long* abssum4, uint32_t slot)
{
    switch (c.mode) {
    case MODE_INDEPENDENT:
    case MODE_EXH_MID_SIDE: return true;
    case MODE_EXH_SIDE: return slot != 2;
    default: {
        uint8_t a, b;
        assignment_slots(fast_assignment(abssum4, c.mode == MODE_FAST_MID_SIDE), &a, &b);
        return slot == a || slot == b;
    }
    }
}

// ---- warp / block reductions -------------------------------------------------------------
__device__ inline unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ inline uint32_t warp_sum_u32(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

// block-wide sum; scratch must hold blockDim.x / 32 entries; all threads get the result
__device__ inline unsigned long long block_sum_u64(unsigned long long v, unsigned long long* scratch)
{
    v = warp_sum_u64(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); w++) t += scratch[w];
    return t;
}

__device__ inline uint32_t block_or_u32(uint32_t v, uint32_t* scratch)
{
    v = __reduce_or_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); w++) t |= scratch[w];
    return t;
}

// ---- bit output ----------------------------------------------------------------------------
// The stream is a big-endian bit string.  Bit b lives in 32-bit word b/32 at bit 31-(b%32) of the
// word's big-endian value.  `words` may be shared or (zero-initialised) global memory; words are kept
// in LOGICAL big-endian value when in shared staging and byte-swapped when they are stored to memory
// that the host reads as bytes.  put_bits ORs the low `nbits` (1..32) of v at bit position `pos`.
template <bool SWAP>
__device__ inline void put_bits(uint32_t* words, unsigned long long pos, uint32_t nbits, uint32_t v)
{
    if (nbits == 0) return;
    if (nbits < 32) v &= (1u << nbits) - 1u;
    unsigned long long w = pos >> 5;
    uint32_t off = (uint32_t)(pos & 31);
    unsigned long long wide = ((unsigned long long)v) << (64 - nbits - off);
    uint32_t hi = (uint32_t)(wide >> 32), lo = (uint32_t)wide;
    if (SWAP) {
        hi = __byte_perm(hi, 0, 0x0123);
        lo = __byte_perm(lo, 0, 0x0123);
    }
    if (hi) atomicOr(words + w, hi);
    if (lo) atomicOr(words + w + 1, lo);
}

// ---- CRC ---------------------------------------------------------------------------------
__host__ __device__ inline uint8_t crc8_update(uint8_t c, uint8_t byte)   // poly 0x07, src/crc.rs:104
{
    c ^= byte;
    for (int b = 0; b < 8; b++) c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
    return c;
}

__host__ __device__ inline uint16_t crc16_table_entry(uint32_t i)   // poly 0x8005, src/crc.rs:154
{
    uint16_t d = (uint16_t)(i << 8);
    for (int b = 0; b < 8; b++) d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1));
    return d;
}

// (a * b) mod P over GF(2), P = x^16 + x^15 + x^2 + 1
__host__ __device__ inline uint32_t gf16_mulmod(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 15; i >= 0; i--) {
        r <<= 1;
        if (r & 0x10000u) r ^= 0x18005u;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

// ---- wide integer helpers ----
// d = a * b + c with 32-bit signed a, b and a 64-bit accumulator: one IMAD.WIDE (the C++ form (long long)a * b
// compiles to a four-instruction 64 x 64 multiply)
__device__ inline long long mad_wide_s32(int32_t a, int32_t b, long long c)
{
    long long d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

// acc + v for a 32-bit v and a 64-bit accumulator in one IMAD.WIDE.U32
__device__ inline unsigned long long acc_u32(unsigned long long acc, uint32_t v)
{
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(d) : "r"(v), "l"(acc));
    return d;
}

// ---- misc ------------------------------------------------------------------------------------
// division by a warp-uniform runtime divisor that is nearly always a power of two (partition sizes n >> p)
struct UDiv {
    uint32_t d, sh;   // sh = log2(d) when d is a power of two, else 32
};
__device__ inline UDiv udiv_make(uint32_t d)
{
    UDiv u;
    u.d = d;
    u.sh = (d != 0 && (d & (d - 1)) == 0) ? 31u - (uint32_t)__clz((int)d) : 32u;
    return u;
}
__device__ inline uint32_t udiv(uint32_t x, UDiv u) { return u.sh < 32 ? x >> u.sh : x / u.d; }

// f64::total_cmp ordering key
__device__ inline long long total_key(double v)
{
    long long x = __double_as_longlong(v);
    x ^= (long long)((unsigned long long)(x >> 63) >> 1);
    return x;
}

__device__ inline int32_t f64_as_i32_sat(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int32_t)v;
}

__host__ __device__ inline uint32_t lpc_precision_for(uint32_t n)   // src/encode.rs:3305-3315
{
    return n <= 192 ? 7 : n <= 384 ? 8 : n <= 576 ? 9 : n <= 1152 ? 10 : n <= 2304 ? 11 : n <= 4608 ? 12 : 13;
}

__device__ inline uint32_t uabs32(int32_t v) { return v < 0 ? 0u - (uint32_t)v : (uint32_t)v; }

// zig-zag exactly as src/encode.rs:3845-3849 (u32 arithmetic)
// (for s < 0: ((-s - 1) << 1) + 1 == ~(2 s); for s >= 0: 2 s)
__device__ inline uint32_t zigzag32(int32_t s) { return ((uint32_t)s << 1) ^ (uint32_t)(s >> 31); }

}   // namespace flacb200
