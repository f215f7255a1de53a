// decode_kernels.cu -- sm_100a kernels of the FLAC frame decoder.
//
// FLAC frames carry no length field, and the reference reads them strictly serially
// (Decoder::read_frame, src/decode.rs:1388).  Here every byte position is tested for a frame header
// in parallel, every candidate is decoded speculatively, and the serial walk is re-established
// afterwards as a linked list (frame -> candidate that starts where it ends) that is ranked by
// pointer jumping.  Pipeline (one stream):
//   k_find<0>   tile scan for sync code + full header validation + CRC-8, per-tile counts   [HBM-bound]
//   k_scan_u32  exclusive scan of the tile counts
//   k_find<1>   same test, ordered compaction of FrameCand records
//   then per group of candidates (as many as the scratch budget allows, so that one thread per frame fills the GPU):
//   k_decode    thread per candidate: subframe headers, Rice/escape residuals, fixed/LPC
//               restoration (history ring in shared memory), wasted-bit shift -> planar scratch
//   k_crc16f    warp per candidate: CRC-16 over [start, end) must leave residue 0
//   k_chain     single CTA, slices of 16384 candidates: links, pointer jumping from the segment heads, ShortBlock/total-sample
//               rules, output positions (segmented scan of block sizes), first error in stream order
//   k_emit      stereo restoration + interleave + narrowing to the caller's PCM layout (Frame::to_buf)
#include <cooperative_groups.h>

#include "common.cuh"
#include "crc.cuh"
#include "decode.cuh"

namespace flacb200 {


// ------------------------------------------------------------------------------------------------
// FrameHeader::parse + STREAMINFO cross-checks + CRC-8  (src/stream.rs:214-240, :279-313, :151-163)
// Returns 0 and fills the fields, or the Error ordinal.  `avail` = bytes from d to the segment end.
// ------------------------------------------------------------------------------------------------
__device__ uint32_t parse_frame_header(const uint8_t* __restrict__ d, unsigned long long avail, const DecCfg& cfg, uint32_t* block_size,
                                       uint32_t* hdr_len, uint32_t* assignment)
{
    if (avail < 4) return 1;   // Io (UnexpectedEof)
    const uint32_t b0 = d[0], b1 = d[1], b2 = d[2], b3 = d[3];
    if (b0 != 0xFF || (b1 & 0xFE) != 0xF8) return 23;   // InvalidSyncCode
    const uint32_t bsc = b2 >> 4, src = b2 & 15, ca = b3 >> 4, bpc = (b3 >> 1) & 7;
    if (bsc == 0) return 24;   // InvalidBlockSize
    uint32_t rate = 0, rate_kind = 0;
    switch (src) {
    case 0:
        if (cfg.subset) return 27;   // NonSubsetSampleRate
        rate = cfg.sample_rate;
        break;
    case 1: rate = 88200; break;   case 2: rate = 176400; break;  case 3: rate = 192000; break;  case 4: rate = 8000; break;
    case 5: rate = 16000; break;   case 6: rate = 22050; break;   case 7: rate = 24000; break;   case 8: rate = 32000; break;
    case 9: rate = 44100; break;   case 10: rate = 48000; break;  case 11: rate = 96000; break;
    case 12: rate_kind = 1; break; case 13: rate_kind = 2; break; case 14: rate_kind = 3; break;
    default: return 26;   // InvalidSampleRate
    }
    if (ca > 10) return 31;   // InvalidChannels
    uint32_t bps;
    switch (bpc) {
    case 0:
        if (cfg.subset) return 28;   // NonSubsetBitsPerSample
        bps = cfg.bps;
        break;
    case 1: bps = 8; break;  case 2: bps = 12; break; case 3: return 33;   // InvalidBitsPerSample
    case 4: bps = 16; break; case 5: bps = 20; break; case 6: bps = 24; break; default: bps = 32; break;
    }
    // frame number (src/stream.rs:1246-1264); the reserved bit before it is skipped unchecked (:226)
    unsigned long long n = 4;
    if (avail < 5) return 1;
    const uint32_t f0 = d[4];
    uint32_t ones = 0;
    while (ones < 8 && (f0 & (0x80u >> ones))) ones++;
    if (ones == 0) n += 1;
    else {
        if (ones == 1 || ones > 7) return 36;   // InvalidFrameNumber
        if (n + ones > avail) return 1;
        for (uint32_t i = 1; i < ones; i++)
            if ((d[4 + i] & 0xC0) != 0x80) return 36;
        n += ones;
    }
    uint32_t bs;
    if (bsc == 6) {
        if (n + 1 > avail) return 1;
        bs = (uint32_t)d[n] + 1;
        n += 1;
    } else if (bsc == 7) {
        if (n + 2 > avail) return 1;
        const uint32_t v = ((uint32_t)d[n] << 8) | d[n + 1];
        if (v == 0xFFFF) return 24;
        bs = v + 1;
        n += 2;
    } else {
        bs = bsc == 1 ? 192u : bsc <= 5 ? (576u << (bsc - 2)) : (256u << (bsc - 8));
    }
    if (rate_kind == 1) {
        if (n + 1 > avail) return 1;
        rate = (uint32_t)d[n] * 1000;
        n += 1;
    } else if (rate_kind) {
        if (n + 2 > avail) return 1;
        rate = ((uint32_t)d[n] << 8) | d[n + 1];
        if (rate_kind == 3) rate *= 10;
        n += 2;
    }
    if (n + 1 > avail) return 1;
    n += 1;   // CRC-8
    const uint32_t channels = ca <= 7 ? ca + 1 : 2;
    if (!cfg.subset) {   // src/stream.rs:291-312, in this order
        if (cfg.max_block_size && bs > cfg.max_block_size) return 25;   // BlockSizeMismatch
        if (rate != cfg.sample_rate) return 29;                         // SampleRateMismatch
        if (channels != cfg.channels) return 32;                        // ChannelsMismatch
        if (bps != cfg.bps) return 35;                                  // BitsPerSampleMismatch
    } else {
        // the batch API decodes into one buffer of fixed shape: subset frames must agree with it too
        if (channels != cfg.channels) return 32;
        if (bps != cfg.bps) return 35;
    }
    uint8_t crc = 0;
    for (uint32_t i = 0; i < (uint32_t)n; i++) crc = crc8_update(crc, d[i]);
    if (crc != 0) return 39;   // Crc8Mismatch
    *block_size = bs;
    *hdr_len = (uint32_t)n;
    *assignment = ca;
    return 0;
}

__device__ inline uint32_t find_segment(const DecSeg* __restrict__ segs, uint32_t nseg, unsigned long long off)
{
    // last segment with byte_off <= off (segments are sorted and disjoint); nseg if none contains off
    uint32_t lo = 0, hi = nseg;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (segs[mid].byte_off <= off) lo = mid + 1;
        else hi = mid;
    }
    if (lo == 0) return nseg;
    const uint32_t s = lo - 1;
    return off < segs[s].byte_end ? s : nseg;
}

// ------------------------------------------------------------------------------------------------
// k_find: candidate frame starts.  Tile of FIND_TILE bytes per CTA, 32 bytes per thread.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t FIND_THREADS = 256;
constexpr uint32_t FIND_TILE = FIND_THREADS * 32;
constexpr uint32_t FIND_SLOTS = 8;   // candidates a tile (8 KB of frames) can park; ordinary streams have one or two

// WRITE = 0: count per tile.  1: write the candidates at their ordered rank (tile_base from the scan of the counts).
// 2: count AND park the tile's first FIND_SLOTS candidates in its slots (cands = the slot array): the bytes are read once;
// k_find_compact then moves the parked records to their ordered places.  A tile with more candidates raises max_bs[1] and
// the host falls back to the second pass (streams of very small blocks).
template <int WRITE>
__global__ void __launch_bounds__(FIND_THREADS) k_find(DecCfg cfg, const uint8_t* __restrict__ bytes, const DecSeg* __restrict__ segs,
                                                       uint32_t* __restrict__ tile_counts, const uint32_t* __restrict__ tile_base,
                                                       FrameCand* __restrict__ cands, uint32_t* __restrict__ max_bs)
{
    __shared__ uint32_t wsum[FIND_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned long long p0 = (unsigned long long)blockIdx.x * FIND_TILE + (unsigned long long)tid * 32;
    uint32_t w[9];
#pragma unroll
    for (int k = 0; k < 9; k++) w[k] = 0;
    if (p0 + 36 <= cfg.nbytes) {
        const uint4 a = *reinterpret_cast<const uint4*>(bytes + p0), b = *reinterpret_cast<const uint4*>(bytes + p0 + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = *reinterpret_cast<const uint32_t*>(bytes + p0 + 32);
    } else {
        for (uint32_t i = 0; i < 33; i++)
            if (p0 + i < cfg.nbytes) w[i >> 2] |= (uint32_t)bytes[p0 + i] << (8 * (i & 3));
    }
    // bit i of `hits`: bytes p0+i, p0+i+1 look like a sync code (0xFF, 0b1111100x).  Four byte positions per step: d has a
    // zero byte exactly where the byte is 0xFF and its successor is 0xF8 or 0xF9; an exact zero-byte test marks those bytes
    // with 0x80 and a multiplication gathers the four marks into a nibble
    uint32_t hits = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t u = __funnelshift_r(w[k], w[k + 1], 8);                                   // the bytes one position on
        const uint32_t d = ~w[k] | ((u & 0xfefefefeu) ^ 0xf8f8f8f8u);
        const uint32_t z = ~(((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d | 0x7f7f7f7fu);               // 0x80 in exactly the zero bytes of d
        hits |= ((((z >> 7) * 0x00204081u) >> 21) & 0xfu) << (4 * k);                            // bits 0, 8, 16, 24 -> a nibble
    }
    uint32_t valid = 0;
    FrameCand found[4];   // at most 4 candidates per 32 bytes are kept (a header is >= 6 bytes; both passes apply the same cap)
    uint32_t nfound = 0;
    while (hits && valid < 4) {
        const uint32_t i = (uint32_t)__ffs((int)hits) - 1u;
        hits &= hits - 1;
        const unsigned long long off = p0 + i;
        const uint32_t s = find_segment(segs, cfg.nseg, off);
        if (s == cfg.nseg) continue;
        uint32_t bs, hl, ca;
        if (parse_frame_header(bytes + off, segs[s].byte_end - off, cfg, &bs, &hl, &ca) != 0) continue;
        valid++;
        if (WRITE) {
            FrameCand c;
            c.off = off; c.block_size = bs; c.seg = s; c.hdr_len = (uint8_t)hl; c.assignment = (uint8_t)ca;
            c.pad0 = c.pad1 = 0; c.pad2 = 0;
            found[nfound++] = c;
        }
        if (WRITE != 1) atomicMax(max_bs, bs);
    }
    // ordered rank of this thread's candidates inside the tile
    uint32_t incl = valid;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t k = 0; k < FIND_THREADS / 32; k++) {
        if (k < wid) before += wsum[k];
        total += wsum[k];
    }
    if (WRITE != 1) {
        if (tid == 0) {
            tile_counts[blockIdx.x] = total;
            if (WRITE == 2 && total > FIND_SLOTS) max_bs[1] = 1;
        }
        if (WRITE == 0) return;
    }
    const uint32_t rank = (WRITE == 1 ? tile_base[blockIdx.x] : 0u) + before + incl - valid;
    for (uint32_t k = 0; k < nfound; k++) {
        if (WRITE == 1) cands[rank + k] = found[k];
        else if (rank + k < FIND_SLOTS) cands[(size_t)blockIdx.x * FIND_SLOTS + rank + k] = found[k];
    }
}

// parked candidates -> their ordered places
__global__ void __launch_bounds__(256) k_find_compact(uint32_t tiles, const uint32_t* __restrict__ tile_counts, const uint32_t* __restrict__ tile_base,
                                                      const FrameCand* __restrict__ slots, FrameCand* __restrict__ cands)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const uint32_t tile = (uint32_t)(i / FIND_SLOTS), s = (uint32_t)(i % FIND_SLOTS);
    if (tile < tiles && s < tile_counts[tile]) cands[tile_base[tile] + s] = slots[i];
}

// exclusive scan of n counts (single CTA of 1024 threads, eight consecutive counts per thread and pass: the loop is a chain of
// block-wide barriers, so its time is the number of passes); total written to out[n]
constexpr uint32_t SCAN_PER = 8;

__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t n, const uint32_t* __restrict__ in, uint32_t* __restrict__ out)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t tile_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += 1024 * SCAN_PER) {
        const uint32_t i0 = base + tid * SCAN_PER;
        uint32_t v[SCAN_PER];
        uint32_t mine = 0;
#pragma unroll
        for (uint32_t k = 0; k < SCAN_PER; k++) {
            v[k] = i0 + k < n ? in[i0 + k] : 0u;
            mine += v[k];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const uint32_t ws = wsum[lane];
            uint32_t wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o) wi += t;
            }
            wsum[lane] = wi - ws;
            if (lane == 31) tile_total = wi;
        }
        __syncthreads();
        uint32_t run = carry + wsum[wid] + incl - mine;
#pragma unroll
        for (uint32_t k = 0; k < SCAN_PER; k++) {
            if (i0 + k < n) out[i0 + k] = run;
            run += v[k];
        }
        carry += tile_total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

// ------------------------------------------------------------------------------------------------
// Bit reader over the frames buffer: 128-bit loads, 64-bit MSB-aligned window with >= 32 valid bits
// ------------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* bytes;
    unsigned long long nbytes;   // bytes that may be read
    unsigned long long widx;     // next 32-bit word to fetch
    unsigned long long endbit;   // bit position of the segment end
    unsigned long long buf;
    uint4 q;
    uint32_t cnt;
    uint32_t err;

    uint4 qn;   // the vector after q, requested one vector (128 bits of stream) before it is needed
    __device__ inline uint4 read_vec(unsigned long long vidx) const
    {
        const unsigned long long b = vidx << 4;
        if (b + 16 <= nbytes) return *reinterpret_cast<const uint4*>(bytes + b);
        uint32_t t[4] = {0, 0, 0, 0};
        for (uint32_t i = 0; i < 16; i++)
            if (b + i < nbytes) t[i >> 2] |= (uint32_t)bytes[b + i] << (8 * (i & 3));
        return make_uint4(t[0], t[1], t[2], t[3]);
    }
    __device__ inline void load_vec(unsigned long long vidx)   // q = vector vidx (already in flight), qn = the next one
    {
        q = qn;
        qn = read_vec(vidx + 1);
    }
    __device__ inline uint32_t fetch()
    {
        const uint32_t sub = (uint32_t)(widx & 3);
        if (sub == 0) load_vec(widx >> 2);
        const uint32_t v = sub == 0 ? q.x : sub == 1 ? q.y : sub == 2 ? q.z : q.w;
        widx++;
        return __byte_perm(v, 0, 0x0123);
    }
    __device__ inline void init(unsigned long long bitpos)
    {
        err = 0;
        widx = bitpos >> 5;
        qn = read_vec(widx >> 2);
        load_vec(widx >> 2);
        const uint32_t hi = fetch_noload(), skip = (uint32_t)(bitpos & 31);
        const uint32_t lo = fetch();
        buf = (((unsigned long long)hi << 32) | lo) << skip;
        cnt = 64 - skip;
        if (cnt < 32) refill();
    }
    __device__ inline uint32_t fetch_noload()
    {
        const uint32_t sub = (uint32_t)(widx & 3);
        const uint32_t v = sub == 0 ? q.x : sub == 1 ? q.y : sub == 2 ? q.z : q.w;
        widx++;
        return __byte_perm(v, 0, 0x0123);
    }
    __device__ inline void refill()
    {
        const uint32_t w = fetch();
        buf |= (unsigned long long)w << (32 - cnt);
        cnt += 32;
    }
    __device__ inline unsigned long long position() const { return (widx << 5) - cnt; }
    __device__ inline void consume(uint32_t n)   // n <= 32, cnt >= 32
    {
        buf <<= n;
        cnt -= n;
        if (cnt < 32) refill();
    }
    __device__ inline uint32_t get(uint32_t n)   // n in 0..=32
    {
        if (n == 0) return 0;
        const uint32_t v = (uint32_t)(buf >> (64 - n));
        consume(n);
        return v;
    }
    __device__ inline int32_t get_signed(uint32_t n)   // n in 1..=32
    {
        const uint32_t v = get(n);
        return (int32_t)(v << (32 - n)) >> (32 - n);
    }
    __device__ inline long long get_signed64(uint32_t n)   // n in 1..=33
    {
        if (n <= 32) return get_signed(n);
        const unsigned long long hi = get(n - 32);
        const unsigned long long v = (hi << 32) | get(32);
        return (long long)(v << (64 - n)) >> (64 - n);
    }
    __device__ inline uint32_t unary()   // read_unary::<1>: zeros up to the next one bit
    {
        uint32_t qv = 0;
        for (;;) {
            const uint32_t top = (uint32_t)(buf >> 32);
            if (top) {
                const uint32_t lz = (uint32_t)__clz((int)top);
                consume(lz + 1);
                return qv + lz;
            }
            qv += 32;
            consume(32);
            if (position() > endbit) {
                err = 1;
                return qv;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// k_decode: read_subframes (src/decode.rs:1494-1856), one thread per candidate frame
// ------------------------------------------------------------------------------------------------
constexpr uint32_t DEC_THREADS = 128;

struct PlaneWriter {
    int32_t* plane;
    uint32_t i, wasted;
    int32_t o0, o1, o2;
    __device__ inline void push(int32_t v)
    {
        v = (int32_t)((uint32_t)v << wasted);   // `<<= wasted_bps`  src/decode.rs:1671
        const uint32_t ph = i & 3;
        if (ph == 0) o0 = v;
        else if (ph == 1) o1 = v;
        else if (ph == 2) o2 = v;
        else *reinterpret_cast<int4*>(plane + plane_off(i - 3)) = make_int4(o0, o1, o2, v);
        i++;
    }
    __device__ inline void flush()
    {
        const uint32_t ph = i & 3, b = i - ph;
        if (ph >= 1) plane[plane_off(b)] = o0;
        if (ph >= 2) plane[plane_off(b + 1)] = o1;
        if (ph >= 3) plane[plane_off(b + 2)] = o2;
    }
};

struct ResidualSpec {
    uint32_t n, order, nres, chunk, porder, hb, esc_code, shift;
};

// read_residuals (src/decode.rs:1800-1856) + predict (:1738-1752) for one subframe.
// HB > 0: predictor order <= HB, coefficients and the last HB samples live in registers; samples are produced in blocks
// of 8 with static register indexing.  HB == 0: no predictor.  HB < 0: any order, rings in shared memory.
// rare paths of the residual loop
// Rice code that does not fit the 32-bit window: returns msb, lsb through u
__device__ inline uint32_t rice_slow(BitReader& br, uint32_t k, uint32_t* u)
{
    const uint32_t msb = br.unary();
    if (br.err) return br.err;
    *u = (msb << k) | br.get(k);   // src/decode.rs:1827
    return 0;
}

// ResidualPartitionHeader (src/stream.rs:1586-1600): returns (mode << 8) | k, or 0xFFFFFFFF past the segment end
__device__ inline uint32_t partition_header(BitReader& br, uint32_t hb, uint32_t esc_code)
{
    uint32_t k = br.get(hb), mode = 0;
    if (k == esc_code) {
        k = br.get(5);
        mode = k ? 1 : 2;
    }
    if (br.position() > br.endbit) return 0xFFFFFFFFu;
    return (mode << 8) | k;
}

template <int HB>
__device__ __forceinline__ uint32_t residual_loop_body(BitReader& br, PlaneWriter& pw, const ResidualSpec& rs, int32_t* hist, int16_t* coef);

template <int HB>
static __device__ __noinline__ uint32_t residual_loop(BitReader& br_io, PlaneWriter& pw_io, const ResidualSpec& rs, int32_t* hist, int16_t* coef)
{
    // the reader/writer state is copied into locals whose address never escapes (registers) and written back at the end;
    // through the references every access would be a local-memory round trip
    BitReader br = br_io;
    PlaneWriter pw = pw_io;
    const uint32_t rc = residual_loop_body<HB>(br, pw, rs, hist, coef);
    br_io = br;
    pw_io = pw;
    return rc;
}

template <int HB>
__device__ __forceinline__ uint32_t residual_loop_body(BitReader& br, PlaneWriter& pw, const ResidualSpec& rs, int32_t* hist, int16_t* coef)
{
    constexpr int W = HB > 0 ? HB : 1;
    constexpr int BLK = 4;   // samples per unrolled block = one 128-bit store; the window shifts by BLK registers after each
    int32_t q[W], w[W + BLK];   // w[W - 1] is the newest sample at the start of a block
    if (HB > 0) {
#pragma unroll
        for (int j = 0; j < W; j++) {
            q[j] = (uint32_t)j < rs.order ? (int32_t)coef[j * DEC_THREADS] : 0;
            w[W - 1 - j] = (uint32_t)j < rs.order ? hist[((rs.order - 1 - j) & 31) * DEC_THREADS] : 0;   // warm-up samples
        }
    }
    pw.flush();   // warm-up samples that did not fill a group of four; from here on samples are stored directly
    int32_t* const plane = pw.plane;
    const uint32_t wasted = pw.wasted;
    uint32_t part_end = 0;       // residual index where the next partition starts
    uint32_t next_len = rs.nres - ((1u << rs.porder) - 1) * rs.chunk;   // the first partition is short by `order`
    uint32_t mode = 0, k = 0;    // mode 0 rice(k), 1 escaped(k bits), 2 all zero
    uint32_t idx = 0;
    uint32_t fail = 0;
    // one residual (read_residuals, src/decode.rs:1800-1856)
    auto residual = [&]() -> int32_t {
        if (idx == part_end) {
            const uint32_t ph = partition_header(br, rs.hb, rs.esc_code);
            if (ph == 0xFFFFFFFFu) fail = 1;
            k = ph & 0x1f;
            mode = (ph >> 8) & 3;
            part_end += next_len;
            next_len = rs.chunk;
        }
        int32_t r;
        if (mode == 0) {
            uint32_t u;
            const uint32_t top = (uint32_t)(br.buf >> 32);
            const uint32_t lz = (uint32_t)__clz((int)top);
            if (top != 0 && lz + 1 + k <= 32) {   // whole code inside the 32-bit window
                const uint32_t lsb = k ? ((top << (lz + 1)) >> (32 - k)) : 0u;
                u = (lz << k) | lsb;               // src/decode.rs:1827
                br.consume(lz + 1 + k);
            } else {
                u = 0;
                const uint32_t e2 = rice_slow(br, k, &u);
                if (e2) fail = e2;
            }
            r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
        } else if (mode == 1) {
            r = br.get_signed(k);
        } else {
            r = 0;
        }
        idx++;
        return r;
    };
    // predict (src/decode.rs:1738-1752) for one sample whose W predecessors are w[base .. base + W)
    // single samples: until the sample index is a multiple of four, and the tail
    auto single = [&]() {
        const uint32_t sidx = rs.order + idx;
        const int32_t r = residual();
        int32_t x;
        if (HB > 0) {
            long long sum = 0;
#pragma unroll
            for (int j = 0; j < W; j++) sum = mad_wide_s32(w[W - 1 - j], q[j], sum);
            x = (int32_t)((uint32_t)r + (uint32_t)(unsigned long long)(sum >> rs.shift));
#pragma unroll
            for (int j = 0; j + 1 < W; j++) w[j] = w[j + 1];
            w[W - 1] = x;
        } else if (HB == 0) {
            x = r;
        } else {
            long long sum = 0;
            for (uint32_t j = 0; j < rs.order; j++)
                sum = mad_wide_s32(hist[((sidx - 1 - j) & 31) * DEC_THREADS], coef[j * DEC_THREADS], sum);
            x = (int32_t)((uint32_t)r + (uint32_t)(unsigned long long)(sum >> rs.shift));
            hist[(sidx & 31) * DEC_THREADS] = x;
        }
        plane[plane_off(sidx)] = (int32_t)((uint32_t)x << wasted);   // `<<= wasted_bps`  src/decode.rs:1671
    };
    while (idx < rs.nres && ((rs.order + idx) & 3u) != 0) {
        single();
        if (fail) return fail;
    }
    while (idx + BLK <= rs.nres) {
        const uint32_t sidx = rs.order + idx;
        int32_t xs[BLK];
#pragma unroll
        for (int e = 0; e < BLK; e++) {
            const int32_t r = residual();
            int32_t x;
            if (HB > 0) {
                long long sum = 0;
#pragma unroll
                for (int j = 0; j < W; j++) sum = mad_wide_s32(w[W + e - 1 - j], q[j], sum);
                x = (int32_t)((uint32_t)r + (uint32_t)(unsigned long long)(sum >> rs.shift));
                w[W + e] = x;
            } else if (HB == 0) {
                x = r;
            } else {
                long long sum = 0;
                for (uint32_t j = 0; j < rs.order; j++)
                    sum = mad_wide_s32(hist[((sidx + e - 1 - j) & 31) * DEC_THREADS], coef[j * DEC_THREADS], sum);
                x = (int32_t)((uint32_t)r + (uint32_t)(unsigned long long)(sum >> rs.shift));
                hist[((sidx + e) & 31) * DEC_THREADS] = x;
            }
            xs[e] = (int32_t)((uint32_t)x << wasted);
        }
        if (fail) return fail;
        *reinterpret_cast<int4*>(plane + plane_off(sidx)) = make_int4(xs[0], xs[1], xs[2], xs[3]);
        if (HB > 0) {
#pragma unroll
            for (int j = 0; j < W; j++) w[j] = w[j + BLK];
        }
    }
    while (idx < rs.nres) {
        single();
        if (fail) return fail;
    }
    pw.i = 0;   // everything is in the plane; nothing left for flush()
    return 0;
}

// One subframe whose samples fit 32 bits.  hist/coef are this thread's columns of the shared rings.
static __device__ __noinline__ uint32_t decode_subframe(BitReader& br, uint32_t bps, uint32_t n, int32_t* __restrict__ plane, int32_t* hist, int16_t* coef)
{
    // SubframeHeader (src/stream.rs:1382-1395, :1537-1553)
    const uint32_t h = br.get(8);
    if (h & 0x80) return 41;   // InvalidSubframeHeader
    const uint32_t type = (h >> 1) & 0x3f;
    uint32_t wasted = 0;
    if (h & 1) wasted = br.unary() + 1;
    if (br.err) return br.err;
    uint32_t kind, order = 0;
    if (type == 0) kind = 0;
    else if (type == 1) kind = 1;
    else if (type >= 8 && type <= 12) { kind = 2; order = type - 8; }
    else if (type >= 32) { kind = 3; order = type - 31; }
    else return 42;   // InvalidSubframeHeaderType
    if (wasted >= bps) return 43;   // ExcessiveWastedBits (src/decode.rs:1644)
    const uint32_t ebps = bps - wasted;
    PlaneWriter pw;
    pw.plane = plane; pw.i = 0; pw.wasted = wasted; pw.o0 = pw.o1 = pw.o2 = 0;
    if (kind == 0) {
        const int32_t v = br.get_signed(ebps);
        for (uint32_t i = 0; i < n; i++) pw.push(v);
        pw.flush();
        return br.position() > br.endbit ? 1u : 0u;
    }
    if (kind == 1) {
        for (uint32_t i = 0; i < n; i++) pw.push(br.get_signed(ebps));
        pw.flush();
        return br.position() > br.endbit ? 1u : 0u;
    }
    if (order > n) return kind == 2 ? 47u : 48u;   // InvalidFixedOrder / InvalidLpcOrder
    for (uint32_t i = 0; i < order; i++) {         // warm-up samples
        const int32_t v = br.get_signed(ebps);
        hist[(i & 31) * DEC_THREADS] = v;
        pw.push(v);
    }
    uint32_t shift = 0;
    if (kind == 2) {
        const int16_t fc[4][4] = {{1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};   // src/stream.rs:1534
        for (uint32_t j = 0; j < order; j++) coef[j * DEC_THREADS] = fc[order - 1][j];
    } else {
        const uint32_t prec = br.get(4) + 1;
        if (prec > 15) return 49;   // InvalidQlpPrecision
        const int32_t sh = br.get_signed(5);
        if (sh < 0) return 50;      // NegativeLpcShift
        shift = (uint32_t)sh;
        for (uint32_t j = 0; j < order; j++) coef[j * DEC_THREADS] = (int16_t)br.get_signed(prec);
    }
    if (br.position() > br.endbit) return 1;
    // read_residuals (src/decode.rs:1800-1856)
    const uint32_t method = br.get(2);
    if (method > 1) return 45;   // InvalidCodingMethod
    const uint32_t hb = method ? 5u : 4u, esc_code = method ? 31u : 15u;
    const uint32_t porder = br.get(4);
    const uint32_t nres = n - order;
    const uint32_t chunk = n >> porder;
    if (chunk == 0) return 46;   // InvalidPartitionOrder (rchunks_mut(0) panics in the reference)
    if ((nres + chunk - 1) / chunk != (1u << porder)) return 46;
    // predict (src/decode.rs:1738-1752): orders up to 16 keep the coefficients and a sliding window of samples in
    // registers (blocks of 8 samples, static indexing); longer predictors use the shared-memory rings
    const ResidualSpec rs = {n, order, nres, chunk, porder, hb, esc_code, shift};
    uint32_t rc;
    switch (order == 0 ? 0u : (order + 3u) >> 2) {
    case 0: rc = residual_loop<0>(br, pw, rs, hist, coef); break;
    case 1: rc = residual_loop<4>(br, pw, rs, hist, coef); break;
    case 2: rc = residual_loop<8>(br, pw, rs, hist, coef); break;
    case 3: rc = residual_loop<12>(br, pw, rs, hist, coef); break;
    case 4: rc = residual_loop<16>(br, pw, rs, hist, coef); break;
    default: rc = residual_loop<-1>(br, pw, rs, hist, coef); break;
    }
    if (rc) return rc;
    pw.flush();
    return br.position() > br.endbit ? 1u : 0u;
}

// The 33-bit side channel of a 32-bit stereo-decorrelated stream (src/decode.rs:1528-1546): i64 samples,
// low words to `plane`, high words to `plane_hi`.  Rare: plain loops over global memory.
__device__ uint32_t decode_subframe_wide(BitReader& br, uint32_t bps, uint32_t n, int32_t* __restrict__ plane, int32_t* __restrict__ plane_hi)
{
    const uint32_t h = br.get(8);
    if (h & 0x80) return 41;
    const uint32_t type = (h >> 1) & 0x3f;
    uint32_t wasted = 0;
    if (h & 1) wasted = br.unary() + 1;
    if (br.err) return br.err;
    uint32_t kind, order = 0;
    if (type == 0) kind = 0;
    else if (type == 1) kind = 1;
    else if (type >= 8 && type <= 12) { kind = 2; order = type - 8; }
    else if (type >= 32) { kind = 3; order = type - 31; }
    else return 42;
    if (wasted >= bps) return 43;
    const uint32_t ebps = bps - wasted;
    auto put = [&](uint32_t i, long long v) {
        plane[plane_off(i)] = (int32_t)(uint32_t)(unsigned long long)v;
        plane_hi[plane_off(i)] = (int32_t)(v >> 32);
    };
    auto at = [&](uint32_t i) -> long long { return (long long)(((unsigned long long)(uint32_t)plane_hi[plane_off(i)] << 32) | (uint32_t)plane[plane_off(i)]); };
    if (kind == 0) {
        const long long v = br.get_signed64(ebps);
        for (uint32_t i = 0; i < n; i++) put(i, v);
    } else if (kind == 1) {
        for (uint32_t i = 0; i < n; i++) put(i, br.get_signed64(ebps));
    } else {
        if (order > n) return kind == 2 ? 47u : 48u;
        for (uint32_t i = 0; i < order; i++) put(i, br.get_signed64(ebps));
        long long cf[32];
        uint32_t shift = 0;
        if (kind == 2) {
            const long long fc[4][4] = {{1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
            for (uint32_t j = 0; j < order; j++) cf[j] = fc[order - 1][j];
        } else {
            const uint32_t prec = br.get(4) + 1;
            if (prec > 15) return 49;
            const int32_t sh = br.get_signed(5);
            if (sh < 0) return 50;
            shift = (uint32_t)sh;
            for (uint32_t j = 0; j < order; j++) cf[j] = br.get_signed(prec);
        }
        if (br.position() > br.endbit) return 1;
        const uint32_t method = br.get(2);
        if (method > 1) return 45;
        const uint32_t hb = method ? 5u : 4u, esc_code = method ? 31u : 15u;
        const uint32_t porder = br.get(4);
        const uint32_t nres = n - order;
        const uint32_t chunk = n >> porder;
        if (chunk == 0) return 46;
        if ((nres + chunk - 1) / chunk != (1u << porder)) return 46;
        uint32_t part_end = 0, next_len = nres - ((1u << porder) - 1) * chunk, mode = 0, k = 0;
        for (uint32_t idx = 0; idx < nres; idx++) {
            if (idx == part_end) {
                k = br.get(hb);
                mode = 0;
                if (k == esc_code) {
                    k = br.get(5);
                    mode = k ? 1 : 2;
                }
                part_end += next_len;
                next_len = chunk;
                if (br.position() > br.endbit) return 1;
            }
            long long r;
            if (mode == 0) {
                const uint32_t msb = br.unary();
                if (br.err) return br.err;
                const uint32_t u = (msb << k) | br.get(k);
                r = (u & 1) ? -(long long)(u >> 1) - 1 : (long long)(u >> 1);
            } else if (mode == 1) r = br.get_signed(k);
            else r = 0;
            const uint32_t s = order + idx;
            long long sum = 0;
            for (uint32_t j = 0; j < order; j++) sum += at(s - 1 - j) * cf[j];
            put(s, r + (sum >> shift));
        }
    }
    if (wasted)
        for (uint32_t i = 0; i < n; i++) put(i, (long long)((unsigned long long)at(i) << wasted));
    return br.position() > br.endbit ? 1u : 0u;
}

__global__ void __launch_bounds__(DEC_THREADS) k_decode(DecCfg cfg, const uint8_t* __restrict__ bytes, const DecSeg* __restrict__ segs,
                                                       const FrameCand* __restrict__ cands, uint32_t ncand, int32_t* __restrict__ planes,
                                                       DecRec* __restrict__ recs, uint32_t only_wide)
{
    __shared__ int32_t s_hist[32 * DEC_THREADS];
    __shared__ int16_t s_coef[32 * DEC_THREADS];
    const uint32_t c = blockIdx.x * DEC_THREADS + threadIdx.x;
    if (c >= ncand) return;
    const FrameCand fc = cands[c];
    if (only_wide && !(fc.assignment >= 8 && cfg.bps == 32)) return;   // every other frame went through k_parse / k_restore
    const DecSeg sg = segs[fc.seg];
    BitReader br;
    br.bytes = bytes;
    br.nbytes = cfg.nbytes;
    br.endbit = sg.byte_end * 8;
    br.init((fc.off + fc.hdr_len) * 8);
    const uint32_t n = fc.block_size;
    int32_t* hist = s_hist + threadIdx.x;
    int16_t* coef = s_coef + threadIdx.x;
    uint32_t err = 0, wide = 0;
    if (n > cfg.bstride) err = 25;   // cannot happen when max_block_size was honoured
    const uint32_t ca = fc.assignment;
    if (!err) {
        if (ca <= 7) {
            for (uint32_t ch = 0; ch <= ca && !err; ch++) err = decode_subframe(br, cfg.bps, n, planes + plane_base(cfg, c, ch), hist, coef);
        } else {
            // 8: left, side   9: side, right   10: mid, side   (src/decode.rs:1512-1626)
            const uint32_t side_first = ca == 9;
            wide = cfg.bps == 32;
            for (uint32_t ch = 0; ch < 2 && !err; ch++) {
                const bool is_side = (ch == 0) == (side_first != 0);
                if (is_side && wide) err = decode_subframe_wide(br, cfg.bps + 1, n, planes + plane_base(cfg, c, ch), planes + plane_base(cfg, c, 2));
                else err = decode_subframe(br, is_side ? cfg.bps + 1 : cfg.bps, n, planes + plane_base(cfg, c, ch), hist, coef);
            }
        }
    }
    DecRec rec;
    rec.err = err;
    rec.wide = wide;
    rec.end = 0;
    if (!err) {
        const unsigned long long end = ((br.position() + 7) >> 3) + 2;   // byte_align; CRC-16  (:1629-1630)
        if (end > sg.byte_end) rec.err = 1;
        rec.end = end;
    }
    recs[c] = rec;
}

// ------------------------------------------------------------------------------------------------
// k_crc16f: CRC-16 residue of [off, end) must be 0 (src/decode.rs:1429); warp per candidate
// ------------------------------------------------------------------------------------------------
constexpr uint32_t CRCF_THREADS = 256;

__device__ Crc16Fold g_dec_crc16_tabs;   // built once per device, read through L1 (no per-CTA copy)

__global__ void k_dec_crc16_tables_init()
{
    crc16_fold_init(g_dec_crc16_tabs);
}

__global__ void __launch_bounds__(CRCF_THREADS) k_crc16f(const uint8_t* __restrict__ bytes, const FrameCand* __restrict__ cands, uint32_t ncand,
                                                        DecRec* __restrict__ recs)
{
    const uint32_t c = blockIdx.x * (CRCF_THREADS / 32) + (threadIdx.x >> 5);
    if (c >= ncand) return;
    if (recs[c].err) return;
    const unsigned long long off = cands[c].off, end = recs[c].end;
    const uint32_t crc = crc16_warp_fold(g_dec_crc16_tabs, bytes, off, end - off);
    if ((threadIdx.x & 31) == 0 && crc != 0) recs[c].err = 40;   // Crc16Mismatch
}

// ------------------------------------------------------------------------------------------------
// k_chain: the serial frame walk, re-established in parallel for one group of candidates
// ------------------------------------------------------------------------------------------------
constexpr uint32_t CHAIN_THREADS = 1024;
constexpr uint32_t CHAIN_MAX = 16384;   // candidates per slice (links are 16-bit indices)
constexpr uint32_t CHAIN_PER = CHAIN_MAX / CHAIN_THREADS;
constexpr uint32_t CHAIN_NONE = 0xFFFFu;

struct ChainSmem {
    uint16_t jump[2][CHAIN_MAX];
    uint32_t pref[CHAIN_MAX];      // inclusive prefix of marked block sizes
    uint16_t errpref[CHAIN_MAX];   // inclusive prefix of marked-and-failed
    uint8_t mark[CHAIN_MAX];
    uint8_t flag[CHAIN_MAX];       // bit0: successor lives in a later slice, bit1: link broken
    unsigned long long wsum[32];
    unsigned long long first_err;  // (candidate index << 32) | code, minimum wins
    unsigned long long miss;
    uint32_t carry;
    ChainState st;
};

__device__ inline uint32_t find_cand(const FrameCand* __restrict__ cands, uint32_t n, unsigned long long off)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cands[mid].off < off) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && cands[lo].off == off) ? lo : CHAIN_NONE;
}

// first candidate of the slice that belongs to segment `seg` (candidates are sorted by offset, hence by segment)
__device__ inline uint32_t first_of_segment(const FrameCand* __restrict__ cands, uint32_t upto, uint32_t seg)
{
    uint32_t lo = 0, hi = upto;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cands[mid].seg < seg) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// block-wide inclusive scan of one value per thread (1024 threads)
__device__ inline unsigned long long block_scan_incl(unsigned long long v, unsigned long long* wsum)
{
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    __syncthreads();
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const unsigned long long ws = wsum[lane];
        unsigned long long wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o) wi += t;
        }
        wsum[lane] = wi - ws;
    }
    __syncthreads();
    return incl + wsum[wid];
}

// One slice of n (<= CHAIN_MAX) candidates.  first_off/next_off: byte offsets of this slice's and of the next
// slice's first candidate (next_off = ~0 for the very last slice).  sm.st is the walk state carried between slices.
__device__ void chain_slice(ChainSmem& sm, const DecCfg& cfg, const uint8_t* __restrict__ bytes, const DecSeg* __restrict__ segs,
                            const FrameCand* __restrict__ cands, DecRec* __restrict__ recs, uint32_t n, unsigned long long first_off,
                            unsigned long long next_off, bool is_first, unsigned long long* __restrict__ pos_out)
{
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        sm.first_err = ~0ull;
        sm.miss = ~0ull;
        sm.carry = CHAIN_NONE;
    }
    __syncthreads();
    const ChainState st = sm.st;
    const bool active = st.active != 0;
    // ---- heads and links ----
    for (uint32_t c = tid; c < n; c += CHAIN_THREADS) {
        uint16_t nx = CHAIN_NONE;
        uint8_t mark = 0, flag = 0;
        const FrameCand fc = cands[c];
        const DecRec r = recs[c];
        const DecSeg sg = segs[fc.seg];
        if (fc.off == sg.byte_off) mark = 1;
        if (active && fc.seg == st.expect_seg && fc.off == st.expect_off) mark = 1;
        if (r.err == 0 && r.end < sg.byte_end) {
            if (r.end >= next_off) flag = 1;
            else {
                const uint32_t t = find_cand(cands, n, r.end);
                if (t != CHAIN_NONE && cands[t].seg == fc.seg) nx = (uint16_t)t;
                else flag = 2;
            }
        }
        sm.jump[0][c] = nx;
        sm.mark[c] = mark;
        sm.flag[c] = flag;
    }
    __syncthreads();
    // ---- pointer jumping: after round k every frame within 2^(k+1) - 1 links of a head is marked ----
    uint32_t cur = 0;
    for (uint32_t span = 1; span < n; span <<= 1) {
        for (uint32_t c = tid; c < n; c += CHAIN_THREADS) {
            const uint16_t j = sm.jump[cur][c];
            if (j != CHAIN_NONE && sm.mark[c]) sm.mark[j] = 1;
            sm.jump[cur ^ 1][c] = j == CHAIN_NONE ? (uint16_t)CHAIN_NONE : sm.jump[cur][j];
        }
        cur ^= 1;
        __syncthreads();
    }
    // ---- sample positions: prefix sums of block sizes over marked frames, rebased per segment ----
    {
        unsigned long long tv = 0;
        for (uint32_t k = 0; k < CHAIN_PER; k++) {
            const uint32_t c = tid * CHAIN_PER + k;
            if (c < n && sm.mark[c]) tv += cands[c].block_size;
        }
        unsigned long long run = block_scan_incl(tv, sm.wsum) - tv;
        for (uint32_t k = 0; k < CHAIN_PER; k++) {
            const uint32_t c = tid * CHAIN_PER + k;
            if (c < n && sm.mark[c]) run += cands[c].block_size;
            if (c < CHAIN_MAX) sm.pref[c] = (uint32_t)run;
        }
    }
    __syncthreads();
    // ---- per-frame rules that need the stream position (src/decode.rs:1402-1410) ----
    unsigned long long te = 0;
    uint32_t ebits = 0;
    for (uint32_t k = 0; k < CHAIN_PER; k++) {
        const uint32_t c = tid * CHAIN_PER + k;
        if (c >= n) continue;
        if (!sm.mark[c]) {
            pos_out[c] = ~0ull;
            continue;
        }
        const FrameCand fc = cands[c];
        const DecSeg sg = segs[fc.seg];
        const uint32_t lo = first_of_segment(cands, c, fc.seg);
        unsigned long long pos = (unsigned long long)(sm.pref[c] - fc.block_size - (lo ? sm.pref[lo - 1] : 0u));
        if (active && fc.seg == st.expect_seg) pos += st.seg_samples;
        DecRec r = recs[c];
        uint32_t err = r.err;
        if (sg.n_pcm && pos >= sg.n_pcm) {
            // Some(0) => Ok(None): the reader stops at the frame that would start exactly at the announced total, and at
            // everything behind it.  A frame beyond the total is reached only when an earlier one overshot the total (then no
            // frame ever starts exactly there): look for the total among the running sums of this segment in front of c.
            bool stopped = pos == sg.n_pcm;
            if (!stopped) {
                const unsigned long long off = (active && fc.seg == st.expect_seg) ? st.seg_samples : 0ull;
                if (sg.n_pcm >= off) {
                    const unsigned long long base = lo ? sm.pref[lo - 1] : 0u;
                    const unsigned long long V = sg.n_pcm - off + base;
                    if (V == base) stopped = true;
                    else {
                        uint32_t a = lo, b = c;
                        while (a < b) {
                            const uint32_t mid = (a + b) >> 1;
                            if (sm.pref[mid] < V) a = mid + 1;
                            else b = mid;
                        }
                        stopped = a < c && sm.pref[a] == V;
                    }
                }
            }
            if (stopped) {
                pos_out[c] = ~0ull;
                sm.mark[c] = 0;
                continue;
            }
        }
        if (sg.n_pcm) {
            // `total - current_sample` in u64 (src/decode.rs:1400): a frame that overshoots the announced total is accepted when it
            // has more than 14 samples, and from then on the difference has wrapped -- the reader never sees Some(0) again and
            // runs into the end of the bytes (Io) after delivering every frame (release builds; a debug build panics)
            const unsigned long long remaining = sg.n_pcm - pos;   // wraps like the reference's
            if (!(fc.block_size == remaining || fc.block_size > 14)) err = 21;   // ShortBlock
        }
        if (err == 0 && sg.pcm_off + pos + fc.block_size > cfg.out_samples) err = 0x80000000u;   // output too small
        recs[c].err = err;
        pos_out[c] = err ? ~0ull : sg.pcm_off + pos;
        if (err) {
            te++;
            ebits |= 1u << k;
            continue;
        }
        const bool done = sg.n_pcm != 0 && pos + fc.block_size == sg.n_pcm;
        if (sm.flag[c] & 2) {   // the bytes after this frame are not a valid frame header
            uint32_t bs, hl, ca;
            uint32_t he = parse_frame_header(bytes + r.end, sg.byte_end - r.end, cfg, &bs, &hl, &ca);
            const bool eof_ok = sg.n_pcm == 0 && he == 1 && sg.byte_end - r.end < 16;   // src/decode.rs:1416
            if (!eof_ok && !done) {
                if (he == 0) he = 23;
                atomicMin(&sm.first_err, ((unsigned long long)c << 32) | 0x40000000u | he);   // error of the NEXT frame
            }
        } else if (r.end == sg.byte_end) {
            if (sg.n_pcm && !done) atomicMin(&sm.first_err, ((unsigned long long)c << 32) | 0x40000000u | 1u);   // Io: ends early
        } else if ((sm.flag[c] & 1) && !done) {
            sm.carry = c;   // at most one chain leaves a slice (segments are disjoint and sorted)
        }
    }
    {
        unsigned long long erun = block_scan_incl(te, sm.wsum) - te;
        for (uint32_t k = 0; k < CHAIN_PER; k++) {
            const uint32_t c = tid * CHAIN_PER + k;
            if (ebits & (1u << k)) erun++;
            sm.errpref[c] = (uint16_t)erun;
        }
    }
    __syncthreads();
    // frames behind a failed frame of the same segment were never reached by the reference
    unsigned long long emitted = 0, samples = 0;
    for (uint32_t k = 0; k < CHAIN_PER; k++) {
        const uint32_t c = tid * CHAIN_PER + k;
        if (c >= n || !sm.mark[c]) continue;
        const FrameCand fc = cands[c];
        const uint32_t lo = first_of_segment(cands, c, fc.seg);
        const uint32_t errs_before = (uint32_t)(c ? sm.errpref[c - 1] : 0) - (uint32_t)(lo ? sm.errpref[lo - 1] : 0);
        const uint32_t err = recs[c].err;
        if (errs_before) {
            pos_out[c] = ~0ull;
            sm.mark[c] = 2;   // dropped
            if (sm.carry == c) sm.carry = CHAIN_NONE;
        } else if (err) {
            atomicMin(&sm.first_err, ((unsigned long long)c << 32) | (err & 0xBFFFFFFFu));
        } else {
            emitted++;
            samples += fc.block_size;
        }
    }
    emitted = block_scan_incl(emitted, sm.wsum);
    __syncthreads();
    samples = block_scan_incl(samples, sm.wsum);
    // segment heads that fall into this slice's byte range but are not candidates at all
    for (uint32_t s = tid; s < cfg.nseg; s += CHAIN_THREADS) {
        const DecSeg sg = segs[s];
        if (sg.byte_end <= sg.byte_off) continue;
        const bool mine = (is_first || sg.byte_off >= first_off) && sg.byte_off < next_off;
        if (mine && find_cand(cands, n, sg.byte_off) == CHAIN_NONE) atomicMin(&sm.miss, (unsigned long long)s);
    }
    __syncthreads();
    if (tid == CHAIN_THREADS - 1) {
        ChainState ns = st;
        uint32_t err = 0;
        unsigned long long err_frame = 0;
        const unsigned long long ferr = sm.first_err;
        if (ferr != ~0ull) {
            const uint32_t c = (uint32_t)(ferr >> 32);
            const uint32_t code = (uint32_t)ferr;
            if (sm.mark[c] == 1) {
                unsigned long long before = 0;   // emitted frames of this slice that precede the failing one
                for (uint32_t i = 0; i < c; i++) before += (sm.mark[i] == 1 && recs[i].err == 0) ? 1 : 0;
                if (code & 0x40000000u) before += 1;   // the failing frame is the one after c
                err = (code & 0x80000000u) ? 0x80000000u : (code & 0x3FFFFFFFu);
                err_frame = st.frames_total + before;
            }
        }
        // the chain that was expected to continue in this slice
        if (active && st.expect_off < next_off && find_cand(cands, n, st.expect_off) == CHAIN_NONE) {
            const DecSeg sg = segs[st.expect_seg];
            uint32_t bs, hl, ca;
            const uint32_t he = parse_frame_header(bytes + st.expect_off, sg.byte_end - st.expect_off, cfg, &bs, &hl, &ca);
            const bool eof_ok = sg.n_pcm == 0 && he == 1 && sg.byte_end - st.expect_off < 16;
            if (!eof_ok) {
                err = he ? he : 23;
                err_frame = st.frames_total;
            }
        }
        if (sm.miss != ~0ull && err == 0) {
            const DecSeg sg = segs[(uint32_t)sm.miss];
            uint32_t bs, hl, ca;
            const uint32_t he = parse_frame_header(bytes + sg.byte_off, sg.byte_end - sg.byte_off, cfg, &bs, &hl, &ca);
            const bool eof_ok = sg.n_pcm == 0 && he == 1 && sg.byte_end - sg.byte_off < 16;
            if (!eof_ok) {
                err = he ? he : 23;
                err_frame = st.frames_total;
            }
        }
        if (ns.err == 0 && err) {
            ns.err = err;
            ns.err_frame = err_frame;
        }
        ns.frames_total = st.frames_total + emitted;
        ns.samples_total = st.samples_total + samples;
        const bool keep_waiting = active && st.expect_off >= next_off;
        if (sm.carry != CHAIN_NONE) {
            const uint32_t c = sm.carry;
            const FrameCand fc = cands[c];
            ns.active = 1;
            ns.expect_seg = fc.seg;
            ns.expect_off = recs[c].end;
            ns.seg_samples = pos_out[c] - segs[fc.seg].pcm_off + fc.block_size;
        } else if (!keep_waiting) {
            ns.active = 0;
        }
        sm.st = ns;
    }
    __syncthreads();
}

// single CTA; dynamic shared memory = sizeof(ChainSmem).  Walks the candidates [0, ntotal) of one decode group in
// slices; `group_first`/`group_last` say whether earlier/later groups exist (for the byte-range ownership of heads).
__global__ void __launch_bounds__(CHAIN_THREADS) k_chain(DecCfg cfg, const uint8_t* __restrict__ bytes, const DecSeg* __restrict__ segs,
                                                        const FrameCand* __restrict__ cands, DecRec* __restrict__ recs, uint32_t ntotal,
                                                        const FrameCand* __restrict__ after, uint32_t group_first,
                                                        unsigned long long* __restrict__ pos_out, ChainState* __restrict__ state)
{
    extern __shared__ __align__(16) uint8_t chain_dyn[];
    ChainSmem& sm = *reinterpret_cast<ChainSmem*>(chain_dyn);
    if (threadIdx.x == 0) sm.st = *state;
    __syncthreads();
    const unsigned long long after_off = after ? after->off : ~0ull;   // first candidate of the next group
    uint32_t s0 = 0;
    do {
        const uint32_t n = min(CHAIN_MAX, ntotal - s0);
        const unsigned long long first_off = n ? cands[s0].off : 0;
        const unsigned long long next_off = s0 + n < ntotal ? cands[s0 + n].off : after_off;
        chain_slice(sm, cfg, bytes, segs, cands + s0, recs + s0, n, first_off, next_off, group_first && s0 == 0, pos_out + s0);
        s0 += n;
    } while (s0 < ntotal);
    if (threadIdx.x == 0) *state = sm.st;
}

// ------------------------------------------------------------------------------------------------
// k_chain_fast: the frame walk of a group in which nothing is wrong.  Candidates are of two kinds: frames, and false
// positives (a sync pattern with a valid CRC-8 inside a payload -- a few hundred in a 3.6 GB batch) that fail to decode.
// The group is REGULAR when
//   (1) every candidate that decoded and passed its CRC-16 starts a segment, or starts exactly where the nearest
//       error-free candidate before it ends (or where the previous group said the walk continues);
//   (2) it ends at the end of its segment, or exactly at a later candidate of its segment, and that candidate (if it
//       belongs to this group) is error-free;
//   (3) no failed candidate is a segment head or the expected continuation; every segment that starts in the group's
//       byte range has its head among the candidates; no stream-end rule (src/decode.rs:1402-1410) fires.
// Then the error-free candidates are exactly the frames the reference's serial reader visits (walk backwards from any of
// them: offsets strictly decrease, so the predecessors end at a head), in candidate order, and the output positions
// are a segmented prefix sum of their block sizes: no binary searches, no pointer jumping.  Anything else leaves
// *clean != 1 and the group to k_chain, which reproduces the reference's error semantics.  Single CTA; unlike k_chain
// it needs next to no shared memory, so it runs beside the kernels of the other stream.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t CHAINF_PER = 4;
constexpr uint32_t CHAINF_SCAN = 16;   // false candidates skipped when looking for a neighbour (more: not regular)
constexpr uint32_t CHAINF_THREADS = 512;
constexpr uint32_t CHAINF_CL = 8;      // CTAs of the cluster
constexpr uint32_t CHAINF_CHUNK = CHAINF_THREADS * CHAINF_PER;   // candidates per CTA and pass

// One thread-block CLUSTER of eight CTAs: a single CTA spent 1 ms per step pulling 130 bytes per candidate through one SM.
// A pass covers 8 x 2048 consecutive candidates, CTA r the r-th run of them; every CTA publishes the (sum, head seen) aggregate
// of its run in its shared memory, the cluster synchronises, and each CTA folds the aggregates of the CTAs before it -- read
// through distributed shared memory -- into its prefix and all eight into the carry of the next pass.  The counters of the
// verdict live in CTA 0's shared memory (remote atomics).
__global__ void __cluster_dims__(CHAINF_CL, 1, 1) __launch_bounds__(CHAINF_THREADS)
    k_chain_fast(DecCfg cfg, const DecSeg* __restrict__ segs, const FrameCand* __restrict__ cands, const DecRec* __restrict__ recs, uint32_t n,
                 uint32_t n_all, uint32_t group_first, unsigned long long* __restrict__ pos_out, ChainState* __restrict__ state,
                 uint32_t* __restrict__ clean)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    __shared__ unsigned long long w_sum[CHAINF_THREADS / 32];
    __shared__ uint32_t w_has[CHAINF_THREADS / 32];
    __shared__ unsigned long long a_sum[2];   // this CTA's aggregate of the pass (double-buffered by pass parity)
    __shared__ uint32_t a_has[2];
    __shared__ unsigned long long s_pre, s_carry, s_samples;
    __shared__ uint32_t s_bad, s_heads, s_mine, s_cont, s_last, s_frames;   // (CTA 0's are the cluster's)
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const ChainState st = *state;
    if (n == 0) {   // (every CTA returns: nobody waits at a cluster barrier)
        if (rank == 0 && tid == 0) *clean = 0x102u;
        return;
    }
    if (tid == 0) {
        s_bad = st.err != 0 ? 2u : 0u;
        s_heads = 0;
        s_mine = 0;
        s_cont = 0;
        s_frames = 0;
        s_last = 0xFFFFFFFFu;
        s_samples = 0;
        s_carry = st.active ? st.seg_samples : 0;
    }
    cluster.sync();   // CTA 0's counters are initialised before anybody adds to them
    uint32_t* const g_bad = cluster.map_shared_rank(&s_bad, 0);
    uint32_t* const g_heads = cluster.map_shared_rank(&s_heads, 0);
    uint32_t* const g_mine = cluster.map_shared_rank(&s_mine, 0);
    uint32_t* const g_cont = cluster.map_shared_rank(&s_cont, 0);
    uint32_t* const g_last = cluster.map_shared_rank(&s_last, 0);
    uint32_t* const g_frames = cluster.map_shared_rank(&s_frames, 0);
    unsigned long long* const g_samples = cluster.map_shared_rank(&s_samples, 0);
    const unsigned long long first_off = cands[0].off, next_off = n < n_all ? cands[n].off : ~0ull;
    {   // segments whose first byte lies in this group's byte range
        uint32_t mine = 0;
        for (uint32_t sgi = rank * CHAINF_THREADS + tid; sgi < cfg.nseg; sgi += CHAINF_CL * CHAINF_THREADS) {
            const DecSeg sg = segs[sgi];
            if (sg.byte_end <= sg.byte_off) continue;
            if ((group_first || sg.byte_off >= first_off) && sg.byte_off < next_off) mine++;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0 && mine) atomicAdd(g_mine, mine);
    }
    uint32_t bad = 0, heads = 0, frames = 0;
    unsigned long long samples = 0;
    uint32_t pass = 0;
    for (uint32_t base = 0; base < n; base += CHAINF_CL * CHAINF_CHUNK, pass++) {
        const uint32_t c0 = base + rank * CHAINF_CHUNK + tid * CHAINF_PER;   // (n < 2^32 - 2^15: no wrap)
        unsigned long long ex[CHAINF_PER];   // samples of the open segment before candidate k, counted from the thread's start
        uint32_t hb = 0, gb = 0;             // bit k: candidate k is a segment head / is a frame
        unsigned long long sum = 0;
        uint32_t has = 0;
        uint32_t bs[CHAINF_PER];
#pragma unroll
        for (uint32_t k = 0; k < CHAINF_PER; k++) {
            const uint32_t c = c0 + k;
            ex[k] = 0;
            bs[k] = 0;
            if (c >= n) continue;
            const FrameCand fc = cands[c];
            const DecRec r = recs[c];
            const DecSeg sg = segs[fc.seg];
            const bool head = fc.off == sg.byte_off;
            const bool expected = st.active && fc.seg == st.expect_seg && fc.off == st.expect_off;
            if (r.err) {   // must be a false positive
                if (head || expected) bad |= 4;
                continue;
            }
            gb |= 1u << k;
            frames++;
            if (!head) {   // (1)
                int j = (int)c - 1;
                uint32_t steps = 0;
                while (j >= 0 && recs[j].err != 0 && steps < CHAINF_SCAN) { j--; steps++; }
                if (j < 0) {
                    if (expected) atomicOr(g_cont, 1u);
                    else bad |= 8;
                } else if (recs[j].err != 0 || cands[j].seg != fc.seg || recs[j].end != fc.off) {
                    bad |= 8;
                }
            }
            if (r.end != sg.byte_end) {   // (2)
                uint32_t j = c + 1, steps = 0;
                while (j < n_all && cands[j].off < r.end && steps < CHAINF_SCAN) { j++; steps++; }
                if (!(j < n_all && cands[j].off == r.end && cands[j].seg == fc.seg)) bad |= 16;
                else if (j < n) { if (recs[j].err) bad |= 16; }
                else *g_last = c;   // the walk leaves the group here (only one frame can)
            }
            if (head) {
                hb |= 1u << k;
                heads++;
                sum = 0;
                has = 1;
            }
            ex[k] = sum;
            bs[k] = fc.block_size;
            sum += fc.block_size;
            samples += fc.block_size;
        }
        // CTA-wide inclusive segmented scan of (sum, has); exclusive value = what the threads before contribute
        unsigned long long isum = sum;
        uint32_t ihas = has;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long ts = __shfl_up_sync(0xffffffffu, isum, o);
            const uint32_t th = __shfl_up_sync(0xffffffffu, ihas, o);
            if (lane >= (uint32_t)o) {
                if (!ihas) isum += ts;
                ihas |= th;
            }
        }
        if (lane == 31) {
            w_sum[wid] = isum;
            w_has[wid] = ihas;
        }
        __syncthreads();
        // what precedes this thread inside the CTA's run: warps before it, then lanes before it
        unsigned long long psum = 0;
        uint32_t phas = 0;
        for (uint32_t w = 0; w < wid; w++) {
            if (w_has[w]) { psum = w_sum[w]; phas = 1; }
            else psum += w_sum[w];
        }
        {
            unsigned long long ls = __shfl_up_sync(0xffffffffu, isum, 1);
            uint32_t lh = __shfl_up_sync(0xffffffffu, ihas, 1);
            if (lane == 0) { ls = 0; lh = 0; }
            if (lh) { psum = ls; phas = 1; }
            else psum += ls;
        }
        if (tid == CHAINF_THREADS - 1) {   // the run's aggregate
            a_sum[pass & 1] = has ? sum : psum + sum;
            a_has[pass & 1] = has | phas;
        }
        cluster.sync();   // every CTA's aggregate of this pass is in place (and w_sum / w_has have been consumed)
        if (wid == 0) {
            // lane r holds CTA r's aggregate; fold the carry and the runs before this CTA's, and all of them for the next pass
            unsigned long long as = 0;
            uint32_t ah = 0;
            if (lane < CHAINF_CL) {
                as = *cluster.map_shared_rank(&a_sum[pass & 1], lane);
                ah = *cluster.map_shared_rank(&a_has[pass & 1], lane);
            }
            unsigned long long acc = s_carry, pre = 0;
#pragma unroll
            for (uint32_t r = 0; r < CHAINF_CL; r++) {
                if (r == rank) pre = acc;
                const unsigned long long rs = __shfl_sync(0xffffffffu, as, r);
                const uint32_t rh = __shfl_sync(0xffffffffu, ah, r);
                if (rh) acc = rs;
                else acc += rs;
            }
            if (lane == 0) {
                s_pre = pre;       // samples of the open segment before this CTA's run (the carry included unless a head came first)
                s_carry = acc;     // ... and at the end of the pass
            }
        }
        __syncthreads();
        const unsigned long long before = phas ? psum : s_pre + psum;   // samples of the open segment before this thread
        uint32_t seen = 0;
#pragma unroll
        for (uint32_t k = 0; k < CHAINF_PER; k++) {
            const uint32_t c = c0 + k;
            if (c >= n) continue;
            if (!(gb & (1u << k))) {
                pos_out[c] = ~0ull;
                continue;
            }
            if (hb & (1u << k)) seen = 1;
            const unsigned long long pos = seen ? ex[k] : before + ex[k];
            const FrameCand fc = cands[c];
            const DecSeg sg = segs[fc.seg];
            if (sg.n_pcm) {   // (3) the rules of src/decode.rs:1402-1410 must all be silent
                if (pos >= sg.n_pcm) bad |= 32;
                else {
                    const unsigned long long remaining = sg.n_pcm - pos;
                    if (!(bs[k] == remaining || bs[k] > 14) || bs[k] > remaining) bad |= 32;
                }
                if (recs[c].end == sg.byte_end && pos + bs[k] < sg.n_pcm) bad |= 32;   // the bytes end early
            }
            if (sg.pcm_off + pos + bs[k] > cfg.out_samples) bad |= 64;
            pos_out[c] = sg.pcm_off + pos;
        }
        __syncthreads();   // (s_pre is rewritten in the next pass)
    }
    heads = __reduce_add_sync(0xffffffffu, heads);
    frames = __reduce_add_sync(0xffffffffu, frames);
    bad = __reduce_or_sync(0xffffffffu, bad);
    samples = warp_sum_u64(samples);
    if (lane == 0) {
        if (heads) atomicAdd(g_heads, heads);
        if (frames) atomicAdd(g_frames, frames);
        if (bad) atomicOr(g_bad, bad);
        atomicAdd(g_samples, samples);
    }
    __threadfence();   // pos_out of every CTA is visible to CTA 0's last look
    cluster.sync();
    if (rank == 0 && tid == 0) {
        const bool ok = s_bad == 0 && s_heads == s_mine && (st.active != 0) == (s_cont != 0);
        *clean = ok ? 1u : (0x100u | s_bad | (s_heads != s_mine ? 128u : 0u) | ((st.active != 0) != (s_cont != 0) ? 0x200u : 0u));   // why not: FLACB200_DEBUG
        if (ok) {
            ChainState ns = st;
            ns.frames_total = st.frames_total + s_frames;
            ns.samples_total = st.samples_total + s_samples;
            ns.active = 0;
            if (s_last != 0xFFFFFFFFu) {   // the walk continues in the next group
                const FrameCand fl = cands[s_last];
                ns.active = 1;
                ns.expect_seg = fl.seg;
                ns.expect_off = recs[s_last].end;
                ns.seg_samples = *(volatile unsigned long long*)(pos_out + s_last) - segs[fl.seg].pcm_off + fl.block_size;
            }
            *state = ns;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_emit: stereo restoration (src/decode.rs:1524-1626) + Frame::to_buf (src/audio.rs:110-134)
// ------------------------------------------------------------------------------------------------
__device__ inline void store_sample(uint8_t* __restrict__ out, const DecCfg& cfg, unsigned long long idx, uint32_t ch, int32_t v)
{
    switch (cfg.pcm_kind) {
    case 2: reinterpret_cast<int32_t*>(out)[idx * cfg.channels + ch] = v; return;
    case 3: reinterpret_cast<int32_t*>(out)[(unsigned long long)ch * cfg.planar_stride + idx] = v; return;
    default: break;
    }
    uint8_t* p = out + (idx * cfg.channels + ch) * cfg.bytes_per_sample;
    const uint32_t u = (uint32_t)v;
    const bool be = cfg.pcm_kind == 1;
    switch (cfg.bytes_per_sample) {
    case 1: p[0] = (uint8_t)u; break;
    case 2:
        if (be) { p[0] = (uint8_t)(u >> 8); p[1] = (uint8_t)u; }
        else { p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); }
        break;
    case 3:
        if (be) { p[0] = (uint8_t)(u >> 16); p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)u; }
        else { p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); }
        break;
    default:
        if (be) { p[0] = (uint8_t)(u >> 24); p[1] = (uint8_t)(u >> 16); p[2] = (uint8_t)(u >> 8); p[3] = (uint8_t)u; }
        else { p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); p[3] = (uint8_t)(u >> 24); }
    }
}

// grid (ncand, ceil(bstride / 256))
__global__ void __launch_bounds__(256) k_emit(DecCfg cfg, const FrameCand* __restrict__ cands, const DecRec* __restrict__ recs,
                                              const unsigned long long* __restrict__ pos, const int32_t* __restrict__ planes,
                                              uint8_t* __restrict__ out)
{
    const uint32_t c = blockIdx.x;
    const unsigned long long p = pos[c];
    if (p == ~0ull) return;
    const FrameCand fc = cands[c];
    const uint32_t i = blockIdx.y * 256 + threadIdx.x;
    if (i >= fc.block_size) return;
    const uint32_t ca = fc.assignment;
    const uint32_t po = plane_off(i);
    if (ca <= 7) {
        for (uint32_t ch = 0; ch <= ca; ch++) store_sample(out, cfg, p + i, ch, planes[plane_base(cfg, c, ch) + po]);
        return;
    }
    const int32_t a = planes[plane_base(cfg, c, 0) + po], b = planes[plane_base(cfg, c, 1) + po];
    int32_t l, r;
    if (recs[c].wide) {
        const long long hi = planes[plane_base(cfg, c, 2) + po];
        if (ca == 8) {          // left, side(33 bit): right = left - side
            const long long side = (long long)(((unsigned long long)hi << 32) | (uint32_t)b);
            l = a;
            r = (int32_t)((long long)a - side);
        } else if (ca == 9) {   // side(33 bit), right: left = side + right
            const long long side = (long long)(((unsigned long long)hi << 32) | (uint32_t)a);
            l = (int32_t)(side + (long long)b);
            r = b;
        } else {                // mid, side(33 bit)   :1612-1622
            const long long side = (long long)(((unsigned long long)hi << 32) | (uint32_t)b);
            const long long sum = (long long)a * 2 + (side < 0 ? (-side) % 2 : side % 2);
            l = (int32_t)((sum + side) >> 1);
            r = (int32_t)((sum - side) >> 1);
        }
    } else if (ca == 8) {
        l = a;
        r = (int32_t)((uint32_t)a - (uint32_t)b);
    } else if (ca == 9) {
        l = (int32_t)((uint32_t)a + (uint32_t)b);
        r = b;
    } else {
        const int32_t sum = (int32_t)((uint32_t)a * 2u + (uint32_t)(b & 1));   // side.abs() % 2 == side & 1  (:1599)
        l = (int32_t)((uint32_t)sum + (uint32_t)b) >> 1;
        r = (int32_t)((uint32_t)sum - (uint32_t)b) >> 1;
    }
    store_sample(out, cfg, p + i, 0, l);
    store_sample(out, cfg, p + i, 1, r);
}

// ------------------------------------------------------------------------------------------------
// launch wrappers (called from engine.cu)
// ------------------------------------------------------------------------------------------------
uint32_t find_tiles(unsigned long long nbytes) { return (uint32_t)((nbytes + FIND_TILE - 1) / FIND_TILE); }

void launch_find_count(const DecCfg& cfg, const uint8_t* bytes, const DecSeg* segs, uint32_t* tile_counts, uint32_t* tile_base, uint32_t* max_bs,
                       cudaStream_t st)
{
    const uint32_t tiles = find_tiles(cfg.nbytes);
    count_launch(), k_find<0><<<tiles, FIND_THREADS, 0, st>>>(cfg, bytes, segs, tile_counts, nullptr, nullptr, max_bs);
    count_launch(), k_scan_u32<<<1, 1024, 0, st>>>(tiles, tile_counts, tile_base);
}

size_t find_slots_bytes(unsigned long long nbytes) { return (size_t)find_tiles(nbytes) * FIND_SLOTS * sizeof(FrameCand); }

// single pass: counts, parked candidates, scan of the counts
void launch_find_park(const DecCfg& cfg, const uint8_t* bytes, const DecSeg* segs, uint32_t* tile_counts, uint32_t* tile_base, FrameCand* slots,
                      uint32_t* max_bs, cudaStream_t st)
{
    const uint32_t tiles = find_tiles(cfg.nbytes);
    count_launch(), k_find<2><<<tiles, FIND_THREADS, 0, st>>>(cfg, bytes, segs, tile_counts, nullptr, slots, max_bs);
    count_launch(), k_scan_u32<<<1, 1024, 0, st>>>(tiles, tile_counts, tile_base);
}

void launch_find_compact(const DecCfg& cfg, const uint32_t* tile_counts, const uint32_t* tile_base, const FrameCand* slots, FrameCand* cands, cudaStream_t st)
{
    const uint32_t tiles = find_tiles(cfg.nbytes);
    const size_t n = (size_t)tiles * FIND_SLOTS;
    count_launch(), k_find_compact<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tiles, tile_counts, tile_base, slots, cands);
}

void launch_find_write(const DecCfg& cfg, const uint8_t* bytes, const DecSeg* segs, const uint32_t* tile_base, FrameCand* cands, cudaStream_t st)
{
    count_launch(), k_find<1><<<find_tiles(cfg.nbytes), FIND_THREADS, 0, st>>>(cfg, bytes, segs, nullptr, tile_base, cands, nullptr);
}

void launch_decode(const DecCfg& cfg, const uint8_t* bytes, const DecSeg* segs, const FrameCand* cands, uint32_t n, int32_t* planes, DecRec* recs,
                   bool only_wide, cudaStream_t st)
{
    count_launch(), k_decode<<<(n + DEC_THREADS - 1) / DEC_THREADS, DEC_THREADS, 0, st>>>(cfg, bytes, segs, cands, n, planes, recs, only_wide ? 1u : 0u);
}

void init_decode_tables(cudaStream_t st) { k_dec_crc16_tables_init<<<1, 256, 0, st>>>(); }

void launch_crc16f(const uint8_t* bytes, const FrameCand* cands, uint32_t n, DecRec* recs, cudaStream_t st)
{
    count_launch(), k_crc16f<<<(n + CRCF_THREADS / 32 - 1) / (CRCF_THREADS / 32), CRCF_THREADS, 0, st>>>(bytes, cands, n, recs);
}

cudaError_t launch_chain(const DecCfg& cfg, const uint8_t* bytes, const DecSeg* segs, const FrameCand* cands, DecRec* recs, uint32_t n,
                         const FrameCand* after, uint32_t group_first, unsigned long long* pos, ChainState* state, cudaStream_t st)
{
    {
        cudaError_t e = cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChainSmem));
        if (e != cudaSuccess) return e;
    }
    count_launch(), k_chain<<<1, CHAIN_THREADS, sizeof(ChainSmem), st>>>(cfg, bytes, segs, cands, recs, n, after, group_first, pos, state);
    return cudaGetLastError();
}

void launch_chain_fast(const DecCfg& cfg, const DecSeg* segs, const FrameCand* cands, const DecRec* recs, uint32_t n, uint32_t n_all,
                       uint32_t group_first, unsigned long long* pos, ChainState* state, uint32_t* clean, cudaStream_t st)
{
    count_launch(), k_chain_fast<<<CHAINF_CL, CHAINF_THREADS, 0, st>>>(cfg, segs, cands, recs, n, n_all, group_first, pos, state, clean);
}

// k_emit4: the common layouts (1 or 2 channels, 2 or 3 bytes per sample, packed bytes).  One CTA takes a bundle of 32
// frames over 32 groups of four samples.  Phase 1 reads the interleaved planes the way they lie (a warp = 32 frames at
// one group: 512 contiguous bytes per load), restores stereo (src/decode.rs:1524-1626), narrows and packs the group's
// 4 * C * B output bytes (Frame::to_buf, src/audio.rs:110-134) into a shared-memory tile [frame][group]; phase 2 writes
// every frame's run of the tile (up to 128 samples = 32 * C * B words) with whole-warp contiguous stores.
constexpr uint32_t EMIT_GROUPS = 32;   // groups of four samples per tile

template <int C, int B>
__global__ void __launch_bounds__(256) k_emit4(DecCfg cfg, const FrameCand* __restrict__ cands, uint32_t ncand, const DecRec* __restrict__ recs,
                                               const unsigned long long* __restrict__ pos, const int32_t* __restrict__ planes,
                                               uint8_t* __restrict__ out)
{
    constexpr uint32_t CB = C * B;                       // words per group of four samples
    constexpr uint32_t ROW = EMIT_GROUPS * CB + 1;       // words per frame row (+1: phase-1 stores of 32 frames hit 32 banks)
    __shared__ uint32_t tile[32 * ROW];
    __shared__ unsigned long long s_byte0[32];           // output byte of the frame's first sample in this tile (~0: nothing)
    __shared__ uint32_t s_bytes[32];                     // valid bytes of the frame in this tile
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c0 = blockIdx.x * 32;
    const uint32_t i0 = blockIdx.y * EMIT_GROUPS * 4;    // first sample of the tile
    const bool be = cfg.pcm_kind == 1;
    {
        const uint32_t c = c0 + lane;
        unsigned long long p = ~0ull;
        uint32_t bs = 0, ca = 0;
        if (c < ncand) {
            p = pos[c];
            const FrameCand fc = cands[c];
            bs = fc.block_size;
            ca = fc.assignment;
        }
        const bool live = p != ~0ull && i0 < bs;
        if (wid == 0) {
            s_byte0[lane] = live ? (p + i0) * (unsigned long long)(C * B) : ~0ull;
            s_bytes[lane] = live ? min(bs - i0, EMIT_GROUPS * 4u) * (C * B) : 0u;
        }
        if (live) {
            const int32_t* pa = planes + plane_base(cfg, c, 0);
            const int32_t* pb = planes + plane_base(cfg, c, C - 1);
#pragma unroll
            for (uint32_t r = 0; r < EMIT_GROUPS / 8; r++) {
                const uint32_t g = wid + 8 * r, i = i0 + 4 * g;
                if (i >= bs) continue;
                const int4 a = *reinterpret_cast<const int4*>(pa + plane_off(i));
                const int32_t av[4] = {a.x, a.y, a.z, a.w};
                int32_t v[4][C];
                if (C == 1) {
#pragma unroll
                    for (int s = 0; s < 4; s++) v[s][0] = av[s];
                } else {
                    const int4 b4 = *reinterpret_cast<const int4*>(pb + plane_off(i));
                    const int32_t bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        int32_t l, rr;
                        if (ca <= 7) { l = av[s]; rr = bv[s]; }
                        else if (ca == 8) { l = av[s]; rr = (int32_t)((uint32_t)av[s] - (uint32_t)bv[s]); }
                        else if (ca == 9) { l = (int32_t)((uint32_t)av[s] + (uint32_t)bv[s]); rr = bv[s]; }
                        else {
                            const int32_t sum = (int32_t)((uint32_t)av[s] * 2u + (uint32_t)(bv[s] & 1));
                            l = (int32_t)((uint32_t)sum + (uint32_t)bv[s]) >> 1;
                            rr = (int32_t)((uint32_t)sum - (uint32_t)bv[s]) >> 1;
                        }
                        v[s][0] = l;
                        v[s][C - 1] = rr;
                    }
                }
                uint32_t w[CB];
#pragma unroll
                for (uint32_t k = 0; k < CB; k++) w[k] = 0;
#pragma unroll
                for (int s = 0; s < 4; s++)
#pragma unroll
                    for (int ch = 0; ch < C; ch++)
#pragma unroll
                        for (int k = 0; k < B; k++) {
                            const int idx = (s * C + ch) * B + k;
                            const uint32_t byte = ((uint32_t)v[s][ch] >> (8 * (be ? B - 1 - k : k))) & 0xffu;
                            w[idx >> 2] |= byte << (8 * (idx & 3));
                        }
#pragma unroll
                for (uint32_t k = 0; k < CB; k++) tile[lane * ROW + g * CB + k] = w[k];
            }
        }
    }
    __syncthreads();
    // phase 2: warp w writes frames 4w .. 4w + 3
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        const uint32_t f = wid * 4 + j;
        const unsigned long long byte0 = s_byte0[f];
        const uint32_t nbytes = s_bytes[f];
        if (byte0 == ~0ull || nbytes == 0) continue;
        const uint32_t* row = tile + f * ROW;
        if (((reinterpret_cast<uintptr_t>(out) + byte0) & 3) == 0) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(out + byte0);
            const uint32_t nw = nbytes >> 2;
            for (uint32_t k = lane; k < nw; k += 32) dst[k] = row[k];
            const uint32_t tail = nbytes & 3;   // a block that ends inside a word
            if (lane < tail) out[byte0 + nw * 4 + lane] = (uint8_t)(row[nw] >> (8 * lane));
        } else {
            for (uint32_t k = lane; k < nbytes; k += 32) out[byte0 + k] = (uint8_t)(row[k >> 2] >> (8 * (k & 3)));
        }
    }
}

void launch_emit(const DecCfg& cfg, const FrameCand* cands, const DecRec* recs, const unsigned long long* pos, const int32_t* planes, uint32_t n,
                 uint8_t* out, cudaStream_t st)
{
    const bool packed = cfg.pcm_kind <= 1 && cfg.nslots == cfg.channels && (cfg.bytes_per_sample == 2 || cfg.bytes_per_sample == 3) &&
                        (reinterpret_cast<uintptr_t>(out) & 3) == 0 && (cfg.bstride & 3) == 0;
    if (packed && cfg.channels <= 2) {
        dim3 grid((n + 31) / 32, (cfg.bstride / 4 + EMIT_GROUPS - 1) / EMIT_GROUPS);
        if (cfg.channels == 1 && cfg.bytes_per_sample == 2) count_launch(), k_emit4<1, 2><<<grid, 256, 0, st>>>(cfg, cands, n, recs, pos, planes, out);
        else if (cfg.channels == 1) count_launch(), k_emit4<1, 3><<<grid, 256, 0, st>>>(cfg, cands, n, recs, pos, planes, out);
        else if (cfg.bytes_per_sample == 2) count_launch(), k_emit4<2, 2><<<grid, 256, 0, st>>>(cfg, cands, n, recs, pos, planes, out);
        else count_launch(), k_emit4<2, 3><<<grid, 256, 0, st>>>(cfg, cands, n, recs, pos, planes, out);
        return;
    }
    dim3 grid(n, (cfg.bstride + 255) / 256);
    count_launch(), k_emit<<<grid, 256, 0, st>>>(cfg, cands, recs, pos, planes, out);
}


}   // namespace flacb200
