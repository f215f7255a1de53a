// decode_parse.cu -- k_parse + k_restore: read_subframes (src/decode.rs:1494-1856) split into a bit-serial half and
// an arithmetic half, both with WARP-UNIFORM control flow.
//
// The serial dependency of a FLAC frame is the bit position: a Rice code can only be found once every code before it
// has been measured, and subframes carry no length.  A thread per candidate frame is therefore the natural grain of
// the bit walk -- but a general per-thread decoder (k_decode, decode_kernels.cu) diverges: the 32 lanes of a warp sit
// in different subframe kinds, predictor orders, partitions and refill states, so the warp executes them one after
// the other (measured: ~1750 cycles per sample and lane).  Here the walk is written as ONE loop over the sample index
// s that all lanes of a warp run in lockstep:
//
//   k_parse    lane per candidate frame.  Per subframe: header (:1635-1676), then for s = 0 .. n: one "token" per lane
//              -- a warm-up/verbatim sample (raw bits), a Rice code (:1825-1827), an escaped residual (raw bits), or a
//              constant -- with the rare events (LPC parameters + residual coding header at s == order, partition headers
//              at s == j * (n >> partition order); :1698-1733, :1800-1822) handled when a lane reaches them.  The bit
//              window is two 32-bit big-endian words and a bit offset (one funnel shift per look), the word after them
//              comes from a shared-memory ring that cp.async fills seven 16-byte chunks ahead.  Groups of four samples
//              in which every lane decodes plain Rice codes take a fast path without event or kind checks.  Output: the plane of each subframe holds [warm-up | residuals] (stored as
//              aligned 128-bit groups, every lane at the same s), plus a SubRec (kind, order, shift, wasted bits,
//              coefficients) per subframe and the DecRec (end offset, error) per frame.
//   k_restore  lane per subframe: predict (:1738-1752) in place over the plane, coefficients and a sliding window of
//              samples in registers, four samples per 128-bit load/store, then `<<= wasted` (:1671).  The predictor
//              length is the warp's largest order rounded up to a multiple of four (shorter predictors run with zero
//              coefficients), so the loop is uniform as well.  Fixed predictors are LPC with the FIXED_COEFFS
//              (src/stream.rs:1534) and shift 0; constant and verbatim subframes are order 0.
//
// Frames whose side channel is 33 bits wide (32-bit streams with stereo decorrelation, :1528-1546) stay with k_decode.
#include "common.cuh"
#include "decode.cuh"

namespace flacb200 {

constexpr uint32_t PARSE_THREADS = 64;
constexpr uint32_t RESTORE_THREADS = 64;
constexpr uint32_t RE_MAX_ORDER = 16;   // longest predictor of the fused restoration + emit kernel (k_restore_emit)

enum : uint32_t { TK_RAW = 0, TK_RICE = 1, TK_FILL = 2 };

__constant__ int16_t c_fixed_coeffs[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};   // src/stream.rs:1534

// two consecutive big-endian words of the stream (w0, w1) and a bit offset pos < 32 into w0; the two words after them
// are in flight as raw little-endian loads (r1, r2) and are byte-swapped only when they move into w1, so that a load
// has two word periods to land.  Words are addressed relative to the lane's first word (32-bit index, one compare
// against the number of whole words left in the buffer).
// Bit window of one lane: two consecutive big-endian words of the stream (w0, w1) and a bit offset pos < 32 into w0;
// r1 is the word after them (still little-endian), read from shared memory one word period before it is needed.
// The stream is staged through a per-lane ring of PARSE_CHUNKS 16-byte chunks in shared memory ([chunk][lane] layout),
// filled by cp.async seven chunks ahead of the window.  A lane walks its own frame, so every load of a warp touches 32
// different sectors and some lanes miss L1 in every word period: with the loads in registers the whole (lockstep) warp
// waited for DRAM once per word (ncu: 60 % of all stall samples).
//   * Requests are made at GROUP granularity (four samples): once per group every lane that has room requests one
//     chunk, then ALL lanes commit one cp.async group and wait until at most three groups are pending.  The commit must be
//     uniform: the hardware counts groups per warp, so with per-lane commits "all but the newest N" reached only a
//     token or two back and every word waited for a fresh DRAM request.  A group consumes at most one chunk (4 x 32
//     bits) in the common paths; the code that can consume more (subframe/LPC headers, unary runs longer than the
//     window) calls ensure_now(), which tops the ring up and waits for everything.
//   * Consuming bits is branch-free: pos += n; if that crossed a word, the window registers shift by selects and pos
//     wraps with a mask -- a divergent `if` would be taken by some lane at almost every token.
// Words are addressed relative to the chunk that holds the lane's first word; chunks at and beyond the end of the
// buffer are filled synchronously from the bytes that exist, so no access leaves the buffer, whatever its length.
// Ring depth: 8 chunks (128 bytes per lane) when the launch fills the GPU -- with 32 warps per SM a group of four samples takes
// far longer than a DRAM access; 16 chunks for small launches (one stream, a few hundred frames: one or two warps per SM, a
// group takes ~200 cycles and three groups in flight did not cover the latency of the chunk they wait for).
constexpr uint32_t PARSE_SLOT = PARSE_THREADS * 16;       // bytes between consecutive chunks of a lane

// the chunks at and beyond the end of the buffer: the bytes that exist, zeros after them, stored synchronously (a
// 16-byte cp.async with a short src-size would still be a 16-byte access that crosses the end of the caller's buffer)
static __device__ __noinline__ void fill_tail_chunk(uint32_t smem_addr, const uint8_t* bytes, unsigned long long nbytes, unsigned long long byte_off)
{
    uint32_t t[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < 16; i++)
        if (byte_off + i < nbytes) t[i >> 2] |= (uint32_t)bytes[byte_off + i] << (8 * (i & 3));
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]) : "memory");
}

template <uint32_t PARSE_CHUNKS>
struct LaneBits {
    static constexpr uint32_t PARSE_LEAD = PARSE_CHUNKS - 1;
    const uint4* base16;          // the 16-byte chunk that holds the lane's first word
    const uint8_t* bytes;         // the whole buffer (only the chunks at its end are read through this)
    unsigned long long nbytes;
    unsigned long long first;     // word index of base16's first word
    uint32_t nfull;               // whole chunks available from base16 on (clamped to 32 bits)
    uint32_t ring;                // shared-memory address of this lane's chunk slot 0
    uint32_t req;                 // chunks 0 .. req - 1 have been requested
    uint32_t roff;                // ring byte offset of the word after r1 (the next one to become r1)
    uint32_t bits;                // bits consumed since word `first`: w0 is word first + (bits >> 5), the window starts at bit bits & 31
    uint32_t w0, w1, r1;
    __device__ __forceinline__ uint32_t cur_chunk() const { return (bits + 96u) >> 7; }   // chunk of the word after r1

    __device__ __forceinline__ void request_chunk()   // chunk req -> slot req % PARSE_CHUNKS (no commit)
    {
        const uint32_t slot = ring + (req & (PARSE_CHUNKS - 1)) * PARSE_SLOT;
        if (req < nfull) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(base16 + req) : "memory");
        else fill_tail_chunk(slot, bytes, nbytes, ((first >> 2) + req) << 4);
        req++;
    }
    // start of a group of four samples, executed by all lanes together
    // `check_lag`: the previous group was not a run of plain Rice codes (a plain group consumes at most one chunk and requests
    // one, so it cannot fall behind; partition headers and one-sample partitions with escaped residuals consume a little more)
    __device__ __forceinline__ void ensure_group(bool check_lag)
    {
        if ((int32_t)(req - cur_chunk()) < (int32_t)PARSE_LEAD) request_chunk();
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(PARSE_CHUNKS - 5) : "memory");
        if (check_lag && (int32_t)(req - cur_chunk()) < (int32_t)PARSE_LEAD - 2) ensure_now();
    }
    // before code that may consume more than a chunk at once
    __device__ __forceinline__ void ensure_now()
    {
        while ((int32_t)(req - cur_chunk()) < (int32_t)PARSE_LEAD) request_chunk();
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    static __device__ __forceinline__ uint32_t lds(uint32_t addr)
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        return v;
    }
    static __device__ __forceinline__ uint32_t word_off(uint32_t w) { return ((w >> 2) & (PARSE_CHUNKS - 1)) * PARSE_SLOT + (w & 3) * 4; }
    __device__ __forceinline__ void init(const uint8_t* buf, unsigned long long buf_bytes, unsigned long long bitpos, uint32_t ring_addr)
    {
        const unsigned long long word = bitpos >> 5;
        first = word & ~3ull;
        const unsigned long long whole = buf_bytes >> 4, c0 = first >> 2;
        nfull = whole > c0 ? (uint32_t)min(whole - c0, 0xFFFFFF00ull) : 0u;
        bytes = buf;
        nbytes = buf_bytes;
        base16 = reinterpret_cast<const uint4*>(buf) + c0;   // (only dereferenced for chunks below nfull)
        ring = ring_addr;
        req = 0;
        const uint32_t rel = (uint32_t)(word & 3);
        bits = rel * 32u + (uint32_t)(bitpos & 31);
        ensure_now();   // the first PARSE_LEAD chunks
        w0 = __byte_perm(lds(ring + word_off(rel)), 0, 0x0123);
        w1 = __byte_perm(lds(ring + word_off(rel + 1)), 0, 0x0123);
        r1 = lds(ring + word_off(rel + 2));
        roff = word_off(rel + 3);
    }
    __device__ __forceinline__ unsigned long long position() const { return (first << 5) + bits; }
    __device__ __forceinline__ uint32_t window() const { return __funnelshift_l(w1, w0, bits); }   // the next 32 bits (the shift wraps at 32)
    __device__ __forceinline__ void skip(uint32_t n)   // n <= 32: at most one word boundary is crossed
    {
        const uint32_t nb = bits + n;
        const bool adv = ((bits ^ nb) & ~31u) != 0;
        bits = nb;
        uint32_t nw = r1;
        if (adv) nw = lds(ring + roff);
        const uint32_t t = roff + 4;
        const uint32_t nroff = (t & 12u) ? t : ((t + PARSE_SLOT - 16) & (PARSE_CHUNKS * PARSE_SLOT - 1));
        w0 = adv ? w1 : w0;
        w1 = adv ? __byte_perm(r1, 0, 0x0123) : w1;
        r1 = nw;
        roff = adv ? nroff : roff;
    }
    __device__ __forceinline__ uint32_t get(uint32_t n)   // n in 0..=32
    {
        const uint32_t v = __funnelshift_lc(window(), 0, n);   // window >> (32 - n), 0 for n == 0
        skip(n);
        return v;
    }
    __device__ __forceinline__ int32_t get_signed(uint32_t n)   // n in 1..=32
    {
        const uint32_t v = get(n);
        return (int32_t)(v << (32 - n)) >> (32 - n);
    }
};

// state of the subframe a lane is walking
struct LaneSub {
    uint32_t kind;      // 0 constant, 1 verbatim, 2 fixed, 3 lpc
    uint32_t order, n;
    uint32_t tk, k;     // current token kind and its bit count (raw: width, rice: parameter)
    int32_t fill;       // TK_FILL value
    uint32_t ev_s;      // sample index of the next event (0xFFFFFFFF: none)
    uint32_t started;   // the residual coding header has been read
    uint32_t hb, esc, chunk, ebps;
};

// the events of a predictive subframe: at s == order the LPC parameters (:1706-1729) and the residual coding header
// (:1806-1822), then (and at every later partition boundary) a ResidualPartitionHeader (src/stream.rs:1586-1600).
// Returns 0 or the Error ordinal.
template <class LaneBitsT>
__device__ __forceinline__ uint32_t parse_event(LaneBitsT& br, LaneSub& sf, unsigned long long endbit, SubRec* __restrict__ rec)
{
    if (!sf.started) {
        sf.started = 1;
        br.ensure_now();   // up to 4 + 5 + 32 * 15 bits of LPC parameters follow
        if (sf.kind == 3) {
            const uint32_t prec = br.get(4) + 1;
            if (prec > 15) return 49;   // InvalidQlpPrecision
            const int32_t sh = br.get_signed(5);
            if (sh < 0) return 50;      // NegativeLpcShift
            rec->shift = (uint8_t)sh;
            for (uint32_t j = 0; j < sf.order; j++) rec->coef[j] = (int16_t)br.get_signed(prec);
        } else {
            rec->shift = 0;
            for (uint32_t j = 0; j < sf.order; j++) rec->coef[j] = c_fixed_coeffs[sf.order][j];
        }
        if (br.position() > endbit) return 1;
        br.ensure_now();
        const uint32_t method = br.get(2);
        if (method > 1) return 45;   // InvalidCodingMethod
        sf.hb = method ? 5u : 4u;
        sf.esc = method ? 31u : 15u;
        const uint32_t porder = br.get(4);
        const uint32_t nres = sf.n - sf.order;
        sf.chunk = sf.n >> porder;
        if (sf.chunk == 0) return 46;   // InvalidPartitionOrder (rchunks_mut(0) panics in the reference)
        if ((nres + sf.chunk - 1) / sf.chunk != (1u << porder)) return 46;
        sf.ev_s = 0;   // the first partition ends at sample `chunk`
    }
    uint32_t k = br.get(sf.hb);
    sf.tk = TK_RICE;
    if (k == sf.esc) {
        k = br.get(5);
        sf.tk = k ? TK_RAW : TK_FILL;
        sf.fill = 0;
    }
    sf.k = k;
    if (br.position() > endbit) return 1;
    sf.ev_s += sf.chunk;
    if (sf.ev_s >= sf.n) sf.ev_s = 0xFFFFFFFFu;
    return 0;
}

// SubframeHeader (src/stream.rs:1382-1395, :1537-1553) and what precedes the sample loop.  Returns 0 or the ordinal.
template <class LaneBitsT>
__device__ __forceinline__ uint32_t parse_subframe_header(LaneBitsT& br, LaneSub& sf, uint32_t bps, uint32_t n, unsigned long long endbit,
                                                           SubRec* __restrict__ rec)
{
    br.ensure_now();
    const uint32_t h = br.get(8);
    if (h & 0x80) return 41;   // InvalidSubframeHeader
    const uint32_t type = (h >> 1) & 0x3f;
    uint32_t wasted = 0;
    if (h & 1) {               // unary(wasted - 1)
        uint32_t q = 0;
        for (;;) {
            const uint32_t win = br.window();
            if (win) {
                const uint32_t lz = (uint32_t)__clz((int)win);
                br.skip(lz + 1);
                q += lz;
                break;
            }
            q += 32;
            br.skip(32);
            br.ensure_now();
            if (br.position() > endbit) return 1;
        }
        wasted = q + 1;
        br.ensure_now();
    }
    uint32_t kind, order = 0;
    if (type == 0) kind = 0;
    else if (type == 1) kind = 1;
    else if (type >= 8 && type <= 12) { kind = 2; order = type - 8; }
    else if (type >= 32) { kind = 3; order = type - 31; }
    else return 42;            // InvalidSubframeHeaderType
    if (wasted >= bps) return 43;   // ExcessiveWastedBits (src/decode.rs:1644)
    sf.ebps = bps - wasted;
    sf.kind = kind;
    sf.order = order;
    sf.n = n;
    sf.started = kind < 2;   // no residual coding header to come
    sf.ev_s = 0xFFFFFFFFu;
    sf.fill = 0;
    rec->kind = (uint8_t)kind;
    rec->order = (uint8_t)order;
    rec->wasted = (uint8_t)wasted;
    rec->shift = 0;
    if (kind == 0) {
        sf.tk = TK_FILL;
        sf.k = 0;
        sf.fill = br.get_signed(sf.ebps);
        rec->order = 0;
        return 0;
    }
    sf.tk = TK_RAW;   // warm-up samples or the verbatim block: `ebps` bits each
    sf.k = sf.ebps;
    if (kind == 1) return 0;
    if (order > n) return kind == 2 ? 47u : 48u;   // InvalidFixedOrder / InvalidLpcOrder
    sf.ev_s = order;
    return 0;
}

template <uint32_t PARSE_CHUNKS>
__global__ void __launch_bounds__(PARSE_THREADS, PARSE_CHUNKS == 8 ? 16 : 8) k_parse(DecCfg cfg, const uint8_t* __restrict__ bytes, const DecSeg* __restrict__ segs,
                                                        const FrameCand* __restrict__ cands, uint32_t ncand, int32_t* __restrict__ planes,
                                                        SubRec* __restrict__ subs, DecRec* __restrict__ recs, uint32_t* __restrict__ high_order)
{
    const uint32_t c = blockIdx.x * PARSE_THREADS + threadIdx.x;
    const bool exists = c < ncand;
    FrameCand fc;
    fc.off = 0; fc.block_size = 0; fc.seg = 0; fc.hdr_len = 0; fc.assignment = 0;
    if (exists) fc = cands[c];
    const uint32_t ca = fc.assignment;
    const bool wide = exists && ca >= 8 && cfg.bps == 32;   // 33-bit side channel: k_decode handles the frame
    unsigned long long endbit = 0, byte_end = 0;
    __shared__ uint4 s_ring[PARSE_CHUNKS * PARSE_THREADS];
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(s_ring + threadIdx.x);
    LaneBits<PARSE_CHUNKS> br;
    br.base16 = reinterpret_cast<const uint4*>(bytes);
    br.bytes = bytes; br.nbytes = cfg.nbytes; br.first = 0; br.nfull = 0; br.ring = ring_addr; br.req = 0; br.roff = 0; br.bits = 0; br.w0 = br.w1 = br.r1 = 0;
    const uint32_t n = fc.block_size;
    uint32_t err = 0;
    bool live = exists && !wide;
    if (live) {
        const DecSeg sg = segs[fc.seg];
        byte_end = sg.byte_end;
        endbit = byte_end * 8;
        br.init(bytes, cfg.nbytes, (fc.off + fc.hdr_len) * 8, ring_addr);
        if (n > cfg.bstride) { err = 25; live = false; }   // cannot happen when max_block_size was honoured
    }
    const uint32_t nch = ca <= 7 ? ca + 1 : 2;
    SubRec* const myrec = subs + (size_t)(exists ? c : 0) * cfg.channels;
    if (exists && wide)
        for (uint32_t ch = 0; ch < cfg.channels; ch++) myrec[ch].kind = 0xFF;
    const uint32_t nch_max = __reduce_max_sync(0xffffffffu, live ? nch : 0u);
    for (uint32_t ch = 0; ch < nch_max; ch++) {
        bool act = live && ch < nch;
        LaneSub sf;
        sf.kind = 0; sf.order = 0; sf.n = 0; sf.tk = TK_FILL; sf.k = 0; sf.fill = 0; sf.ev_s = 0xFFFFFFFFu; sf.started = 1;
        sf.hb = 4; sf.esc = 15; sf.chunk = 1; sf.ebps = 1;
        SubRec* rec = myrec + (act ? ch : 0);
        if (act) {
            // 8: left, side   9: side, right   10: mid, side   (src/decode.rs:1512-1626): the side channel has one bit more
            const bool is_side = ca >= 8 && ((ch == 0) == (ca == 9));
            const uint32_t e = parse_subframe_header(br, sf, is_side ? cfg.bps + 1 : cfg.bps, n, endbit, rec);
            if (e) { err = e; live = act = false; }
            else if (sf.order > RE_MAX_ORDER) atomicOr(high_order, 1u);   // k_restore_emit keeps two predictors of at most this length in registers
        }
        int32_t* const plane = planes + plane_base(cfg, exists ? c : 0, ch);   // groups of four samples are 128 words apart
        const uint32_t n4 = __reduce_max_sync(0xffffffffu, act ? (n + 3u) & ~3u : 0u);
        uint32_t nlane = act ? n : 0u;   // 0 once the lane has failed: it keeps walking the loop without reading
        // one Rice code (src/decode.rs:1825-1827) of this lane; a code longer than the 32-bit window takes the slow branch
        auto rice_token = [&]() -> int32_t {
            const uint32_t win = br.window();
            const uint32_t lz = (uint32_t)__clz((int)win);
            const uint32_t total = lz + 1 + sf.k;
            uint32_t u;
            if (total <= 32) {   // the whole code sits in the window (win != 0)
                const uint32_t lsb = __funnelshift_lc((win << lz) << 1, 0, sf.k);
                u = (lz << sf.k) | lsb;
                br.skip(total);
            } else {
                uint32_t msb = 0;
                uint32_t w2 = win;
                while (w2 == 0) {
                    msb += 32;
                    br.skip(32);
                    br.ensure_now();
                    if (br.position() > endbit) { err = 1; nlane = 0; break; }
                    w2 = br.window();
                }
                const uint32_t lz2 = w2 ? (uint32_t)__clz((int)w2) : 0u;
                br.skip(w2 ? lz2 + 1 : 0u);
                msb += lz2;
                br.ensure_now();
                u = (msb << sf.k) | br.get(sf.k);
            }
            return (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
        };
        bool check_lag = true;
        for (uint32_t s0 = 0; s0 < n4; s0 += 4) {
            int32_t o[4];
            br.ensure_group(check_lag);
            // the common group: every lane of the warp is alive, inside a Rice partition, and has no event before s0 + 4
            const bool plain = s0 + 4 <= nlane && sf.tk == TK_RICE && sf.ev_s - s0 >= 4;
            check_lag = !__all_sync(0xffffffffu, plain);
            if (!check_lag) {
#pragma unroll
                for (int e4 = 0; e4 < 4; e4++) o[e4] = rice_token();
                *reinterpret_cast<int4*>(plane + plane_off(s0)) = make_int4(o[0], o[1], o[2], o[3]);
                continue;
            }
#pragma unroll
            for (int e4 = 0; e4 < 4; e4++) {
                const uint32_t s = s0 + e4;
                int32_t v = 0;
                if (s < nlane) {
                    if (s == sf.ev_s) {
                        const uint32_t e = parse_event(br, sf, endbit, rec);
                        if (e) { err = e; nlane = 0; sf.tk = TK_FILL; sf.fill = 0; sf.ev_s = 0xFFFFFFFFu; }
                    }
                    if (sf.tk == TK_RICE) v = rice_token();
                    else if (sf.tk == TK_RAW) v = br.get_signed(sf.k);
                    else v = sf.fill;
                }
                o[e4] = v;
            }
            if (s0 < nlane) *reinterpret_cast<int4*>(plane + plane_off(s0)) = make_int4(o[0], o[1], o[2], o[3]);
        }
        if (err) live = act = false;
        if (act) {
            // a predictive subframe with order == n never reaches s == order inside the loop: its headers still have to be read
            if (!sf.started) {
                const uint32_t e = parse_event(br, sf, endbit, rec);
                if (e) { err = e; live = act = false; }
            }
            if (act && br.position() > endbit) { err = 1; live = act = false; }
        }
    }
    if (exists && !wide) {
        DecRec r;
        r.err = err;
        r.wide = 0;
        r.end = 0;
        if (!err) {
            const unsigned long long end = ((br.position() + 7) >> 3) + 2;   // byte_align; CRC-16  (:1629-1630)
            if (end > byte_end) r.err = 1;
            r.end = end;
        }
        recs[c] = r;
    }
}

// predict (src/decode.rs:1738-1752) over one plane, HB = predictor length (multiple of 4).  The plane is read through
// a ring of RESTORE_RING 128-bit groups per lane in shared memory, filled by cp.async RESTORE_RING groups ahead: every
// lane reads its own plane, so a load in registers exposed the whole warp to DRAM latency once per group.
constexpr uint32_t RESTORE_RING = 8;

template <int HB>
__device__ __forceinline__ void restore_plane(int32_t* __restrict__ plane, uint32_t n4, uint32_t nmax4, uint32_t order, uint32_t shift,
                                              uint32_t wasted, const SubRec* __restrict__ rec, int4* ring)
{
    constexpr int W = HB > 0 ? HB : 1;
    int32_t q[W], w[W + 4];
#pragma unroll
    for (int j = 0; j < W; j++) q[j] = (HB > 0 && (uint32_t)j < order) ? (int32_t)rec->coef[j] : 0;
#pragma unroll
    for (int j = 0; j < W + 4; j++) w[j] = 0;
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring);
    auto request = [&](uint32_t s0) {   // group at sample s0 -> slot (s0 / 4) % RESTORE_RING; one commit per call
        if (s0 < n4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_addr + ((s0 >> 2) & (RESTORE_RING - 1)) * (RESTORE_THREADS * 16)),
                         "l"(plane + plane_off(s0))
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (uint32_t g = 0; g < RESTORE_RING; g++) request(g * 4);
    for (uint32_t s0 = 0; s0 < nmax4; s0 += 4) {
        const bool on = s0 < n4;
        asm volatile("cp.async.wait_group %0;" ::"n"(RESTORE_RING - 1) : "memory");
        int4 a = make_int4(0, 0, 0, 0);
        if (on) a = ring[((s0 >> 2) & (RESTORE_RING - 1)) * RESTORE_THREADS];
        const int32_t v[4] = {a.x, a.y, a.z, a.w};
        int32_t x[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int32_t xe = v[e];
            if (HB > 0) {
                long long sum = 0;
#pragma unroll
                for (int j = 0; j < W; j++) sum = mad_wide_s32(w[W + e - 1 - j], q[j], sum);
                const uint32_t pred = __funnelshift_r((uint32_t)(unsigned long long)sum, (uint32_t)((unsigned long long)sum >> 32), shift);
                if (s0 + e >= order) xe = (int32_t)((uint32_t)xe + pred);   // warm-up samples are stored as they are
                w[W + e] = xe;
            }
            x[e] = (int32_t)((uint32_t)xe << wasted);   // `<<= wasted_bps`  src/decode.rs:1671
        }
        if (on && (HB > 0 || wasted)) *reinterpret_cast<int4*>(plane + plane_off(s0)) = make_int4(x[0], x[1], x[2], x[3]);
        request(s0 + 4 * RESTORE_RING);   // into the slot just read (its value is in registers: x depends on it)
        if (HB > 0) {
#pragma unroll
            for (int j = 0; j < W; j++) w[j] = w[j + 4];
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(RESTORE_THREADS) k_restore(DecCfg cfg, const FrameCand* __restrict__ cands, uint32_t ncand,
                                                            const SubRec* __restrict__ subs, const DecRec* __restrict__ recs,
                                                            int32_t* __restrict__ planes)
{
    __shared__ int4 s_ring[RESTORE_RING * RESTORE_THREADS];
    int4* const ring = s_ring + threadIdx.x;
    // a warp = one channel of the 32 frames of a bundle: its plane accesses are 512 contiguous bytes
    const uint32_t t = blockIdx.x * RESTORE_THREADS + threadIdx.x;
    const uint32_t wg = t >> 5;
    const uint32_t c = (wg / cfg.channels) * 32 + (t & 31), ch = wg % cfg.channels;
    uint32_t n = 0, order = 0, shift = 0, wasted = 0;
    const SubRec* rec = subs;
    if (c < ncand) {
        const FrameCand fc = cands[c];
        const uint32_t nch = fc.assignment <= 7 ? fc.assignment + 1u : 2u;
        rec = subs + (size_t)c * cfg.channels + ch;
        if (ch < nch && recs[c].err == 0 && rec->kind != 0xFF) {
            n = fc.block_size;
            order = rec->order;
            shift = rec->shift;
            wasted = rec->wasted;
            if (order == 0 && wasted == 0) n = 0;   // constant / verbatim / fixed order 0: the plane already holds the samples
        }
    }
    int32_t* plane = planes + plane_base(cfg, c < ncand ? c : 0, ch);
    const uint32_t n4 = (n + 3u) & ~3u;
    const uint32_t nmax4 = __reduce_max_sync(0xffffffffu, n4);
    const uint32_t cls = __reduce_max_sync(0xffffffffu, n ? (order + 3u) >> 2 : 0u);
    if (nmax4 == 0) return;
    switch (cls) {
    case 0: restore_plane<0>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 1: restore_plane<4>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 2: restore_plane<8>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 3: restore_plane<12>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 4: restore_plane<16>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 5: restore_plane<20>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 6: restore_plane<24>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    case 7: restore_plane<28>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    default: restore_plane<32>(plane, n4, nmax4, order, shift, wasted, rec, ring); break;
    }
}


// ---- k_restore_emit: predictor restoration, stereo restoration and the packed PCM write in ONE pass over the planes --------
// k_restore rewrote every plane in place (one read + one write of 4 bytes per sample) and k_emit4 read it again.  Here a warp
// owns a bundle of 32 frames: lane = frame, BOTH channels of it -- two independent predictor chains per lane (the recurrence is
// latency-bound: a second chain fills its bubbles) and the mid/side arithmetic of src/decode.rs:1524-1626 needs no exchange
// between lanes.  Per step a lane restores one group of four samples per channel (coefficients and sliding windows in
// registers, planes staged by cp.async RE_RING groups ahead), restores stereo, narrows and packs the 4 * C * B bytes
// (Frame::to_buf, src/audio.rs:110-134) into the warp's shared-memory tile [frame][group]; every RE_TG groups the warp writes
// each frame's run of the tile (32 samples = 8 * C * B words) with contiguous stores.  The planes are read once and nothing is
// written back: per sample 4 bytes in + B bytes out instead of 4 + 4 + 4 + B.
// It needs every frame's output position, so it runs after the frame walk (k_chain_fast / k_chain), not beside it.
// Shapes: 1-2 channels, 2-3 bytes per sample, packed output; predictors longer than RE_MAX_ORDER (k_parse raises *high_order)
// leave the launch group to k_restore + k_emit.
constexpr uint32_t RE_RING = 4;     // groups of four samples in flight per lane and channel
constexpr uint32_t RE_TG = 8;       // groups per output tile
constexpr uint32_t RE_WARPS = 2;    // independent bundles per CTA
#ifndef FLACB200_RE_OCC
#define FLACB200_RE_OCC 8            // CTAs per SM the register budget is cut for
#endif

template <int C, int B, int HB>
__device__ __forceinline__ void restore_emit_run(const int32_t* const (&plane)[C], uint32_t n, uint32_t nmax4, const uint32_t (&order)[C],
                                                 const uint32_t (&shift)[C], const uint32_t (&wasted)[C], const SubRec* const (&rec)[C], uint32_t ca,
                                                 bool be, int4* ring, uint32_t* tile, const unsigned long long* s_base, const uint32_t* s_n,
                                                 uint8_t* __restrict__ out)
{
    constexpr int W = HB > 0 ? HB : 1;
    // the sample window is a circular buffer of W + 4 registers with STATIC indices: U steps of four samples bring it back to
    // where it started, so the step loop is unrolled U times and a tile holds a multiple of U steps (no register moves)
    constexpr int WN = W + 4;
    constexpr uint32_t U = HB > 0 ? WN / 4 : 1;
    constexpr uint32_t TG = (RE_TG / U) * U;    // groups per output tile
    constexpr uint32_t CB = C * B;              // bytes per inter-channel sample = words per group of four
    constexpr uint32_t RW = TG * CB, ROW = RW + 1;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n4 = (n + 3u) & ~3u;
    int32_t q[C][W], w[C][W + 4];
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
#pragma unroll
        for (int j = 0; j < W; j++) q[ch][j] = (HB > 0 && (uint32_t)j < order[ch]) ? (int32_t)rec[ch]->coef[j] : 0;
#pragma unroll
        for (int j = 0; j < WN; j++) w[ch][j] = 0;
    }
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16;
    auto request = [&](uint32_t s0) {   // the group at sample s0 of both planes -> slot (s0 / 4) % RE_RING; one commit per call
        if (s0 < n4) {
#pragma unroll
            for (int ch = 0; ch < C; ch++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_addr + (((s0 >> 2) & (RE_RING - 1)) * C + ch) * 512u),
                             "l"(plane[ch] + plane_off(s0))
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (uint32_t g = 0; g < RE_RING; g++) request(g * 4);
    const bool ms = ca == 10;
    const uint32_t add_b = ca == 9 ? 0xFFFFFFFFu : 0u, sub_b = ca == 8 ? 0xFFFFFFFFu : 0u;
    // this lane's frame can take the fast copy-out of a tile: a 4-byte aligned run of whole words
    const bool row_ok = s_base[lane] != ~0ull && ((reinterpret_cast<uintptr_t>(out) + s_base[lane]) & 3) == 0;
    for (uint32_t t0 = 0; t0 < nmax4; t0 += 4 * TG) {
        const uint32_t gmax = min(TG, (nmax4 - t0) >> 2);
#pragma unroll 1
        for (uint32_t g0 = 0; g0 < gmax; g0 += U) {
#pragma unroll
          for (uint32_t u = 0; u < U; u++) {
            const uint32_t g = g0 + u;
            if (g >= gmax) break;
            const uint32_t s0 = t0 + 4 * g;
            const bool on = s0 < n4;
            asm volatile("cp.async.wait_group %0;" ::"n"(RE_RING - 1) : "memory");
            int32_t x[C][4];
#pragma unroll
            for (int ch = 0; ch < C; ch++) {
                int4 a = make_int4(0, 0, 0, 0);
                if (on) a = ring[(((s0 >> 2) & (RE_RING - 1)) * C + ch) * 32 + lane];
                const int32_t v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    int32_t xe = v[e];
                    if (HB > 0) {
                        long long sum = 0;
#pragma unroll
                        for (int j = 0; j < W; j++) sum = mad_wide_s32(w[ch][(W + 4 * u + e - 1 - j + WN) % WN], q[ch][j], sum);
                        const uint32_t pred = __funnelshift_r((uint32_t)(unsigned long long)sum, (uint32_t)((unsigned long long)sum >> 32), shift[ch]);
                        if (s0 + e >= order[ch]) xe = (int32_t)((uint32_t)xe + pred);   // warm-up samples are stored as they are
                        w[ch][(W + 4 * u + e) % WN] = xe;
                    }
                    x[ch][e] = (int32_t)((uint32_t)xe << wasted[ch]);   // `<<= wasted_bps`  src/decode.rs:1671
                }
            }
            request(s0 + 4 * RE_RING);   // into the slot just read
            // stereo restoration (src/decode.rs:1524-1626) without branches -- the lanes of a warp are different frames with
            // different channel assignments -- then the sample's B low bytes in the caller's byte order
            uint32_t val[4 * C];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (C == 1) val[e] = (uint32_t)x[0][e];
                else {
                    const uint32_t av = (uint32_t)x[0][e], bv = (uint32_t)x[C - 1][e];
                    const uint32_t sum = av * 2u + (bv & 1u);                       // mid/side: `mid` with the bit the encoder dropped
                    const uint32_t lm = (uint32_t)((int32_t)(sum + bv) >> 1), rm = (uint32_t)((int32_t)(sum - bv) >> 1);
                    const uint32_t l = av + (bv & add_b);                           // side/right: left = side + right
                    const uint32_t r = bv + ((av - bv - bv) & sub_b);               // left/side: right = left - side
                    val[e * C] = ms ? lm : l;
                    val[e * C + C - 1] = ms ? rm : r;
                }
            }
            if (be) {
#pragma unroll
                for (int i = 0; i < 4 * C; i++) val[i] = __byte_perm(val[i], 0, B == 3 ? 0x3012 : 0x3201);
            }
            uint32_t* const dst = tile + lane * ROW + g * CB;
            if (B == 3) {
#pragma unroll
                for (int i = 0; i < C; i++) {   // four samples -> three words
                    dst[3 * i + 0] = __byte_perm(val[4 * i + 0], val[4 * i + 1], 0x4210);
                    dst[3 * i + 1] = __byte_perm(val[4 * i + 1], val[4 * i + 2], 0x5421);
                    dst[3 * i + 2] = __byte_perm(val[4 * i + 2], val[4 * i + 3], 0x6542);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 2 * C; i++) dst[i] = __byte_perm(val[2 * i], val[2 * i + 1], 0x5410);
            }
          }
        }
        __syncwarp();
        // the warp writes every frame's run of the tile
        if (__all_sync(0xffffffffu, row_ok && n >= t0 + 4 * TG)) {   // the usual tile: 32 whole, aligned runs
            uint8_t* const lane_out = out + (size_t)t0 * CB + lane * 4;
#pragma unroll 8
            for (uint32_t f = 0; f < 32; f++) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(lane_out + s_base[f]);
                const uint32_t* row = tile + f * ROW + lane;
                dst[0] = row[0];
                if (RW > 32 && lane < RW - 32) dst[32] = row[32];
            }
        } else {
#pragma unroll 1
            for (uint32_t f = 0; f < 32; f++) {
                const unsigned long long base = s_base[f];
                const uint32_t nf = s_n[f];
                if (base == ~0ull || nf <= t0) continue;
                const uint32_t nbytes = min(nf - t0, 4 * TG) * CB;
                const unsigned long long byte0 = base + (unsigned long long)t0 * CB;
                const uint32_t* row = tile + f * ROW;
                if (((reinterpret_cast<uintptr_t>(out) + byte0) & 3) == 0) {
                    uint32_t* dst = reinterpret_cast<uint32_t*>(out + byte0);
                    const uint32_t nw = nbytes >> 2;
                    if (lane < nw) dst[lane] = row[lane];
                    if (RW > 32 && lane + 32 < nw) dst[lane + 32] = row[lane + 32];
                    const uint32_t tail = nbytes & 3;   // a block that ends inside a word
                    if (lane < tail) out[byte0 + nw * 4 + lane] = (uint8_t)(row[nw] >> (8 * lane));
                } else {
                    for (uint32_t k = lane; k < nbytes; k += 32) out[byte0 + k] = (uint8_t)(row[k >> 2] >> (8 * (k & 3)));
                }
            }
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int C, int B>
__global__ void __launch_bounds__(32 * RE_WARPS, FLACB200_RE_OCC) k_restore_emit(DecCfg cfg, const FrameCand* __restrict__ cands, uint32_t ncand,
                                                                   const SubRec* __restrict__ subs, const DecRec* __restrict__ recs,
                                                                   const unsigned long long* __restrict__ pos, const int32_t* __restrict__ planes,
                                                                   uint8_t* __restrict__ out, const uint32_t* __restrict__ flags, uint32_t need_clean)
{
    static_assert(RE_TG * C * B <= 64, "a frame's run of the tile is written with at most two stores per lane");
    // flags[0]: k_chain_fast's verdict (1: pos[] is final); flags[1]: a predictor longer than RE_MAX_ORDER was seen
    if (flags[1] != 0 || (need_clean && flags[0] != 1)) return;
    constexpr uint32_t ROW = RE_TG * C * B + 1;
    __shared__ int4 s_ring[RE_WARPS][RE_RING * C * 32];
    __shared__ uint32_t s_tile[RE_WARPS][32 * ROW];
    __shared__ unsigned long long s_base[RE_WARPS][32];
    __shared__ uint32_t s_n[RE_WARPS][32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c = (blockIdx.x * RE_WARPS + wid) * 32 + lane;
    uint32_t n = 0, ca = 0, order[C], shift[C], wasted[C];
    const SubRec* rec[C];
    const int32_t* plane[C];
    unsigned long long base = ~0ull;
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
        order[ch] = shift[ch] = wasted[ch] = 0;
        rec[ch] = subs;
        plane[ch] = planes + plane_base(cfg, c < ncand ? c : 0, ch);
    }
    if (c < ncand) {
        const FrameCand fc = cands[c];
        const unsigned long long p = pos[c];
        const uint32_t nch = fc.assignment <= 7 ? fc.assignment + 1u : 2u;
        bool ok = p != ~0ull && nch == (uint32_t)C && recs[c].err == 0;
#pragma unroll
        for (int ch = 0; ch < C; ch++) {
            rec[ch] = subs + (size_t)c * cfg.channels + ch;
            ok = ok && rec[ch]->kind != 0xFF;
        }
        if (ok) {
            n = fc.block_size;
            ca = fc.assignment;
            base = p * (unsigned long long)(C * B);
#pragma unroll
            for (int ch = 0; ch < C; ch++) {
                order[ch] = rec[ch]->order;
                shift[ch] = rec[ch]->shift;
                wasted[ch] = rec[ch]->wasted;
            }
        }
    }
    s_base[wid][lane] = base;
    s_n[wid][lane] = n;
    __syncwarp();
    uint32_t omax = 0;
#pragma unroll
    for (int ch = 0; ch < C; ch++) omax = max(omax, order[ch]);
    const uint32_t nmax4 = __reduce_max_sync(0xffffffffu, (n + 3u) & ~3u);
    const uint32_t cls = __reduce_max_sync(0xffffffffu, n ? (omax + 3u) >> 2 : 0u);
    if (nmax4 == 0) return;
    const bool be = cfg.pcm_kind == 1;
#define FLACB200_RE(HBV) restore_emit_run<C, B, HBV>(plane, n, nmax4, order, shift, wasted, rec, ca, be, s_ring[wid], s_tile[wid], s_base[wid], s_n[wid], out)
    switch (cls) {
    case 0: FLACB200_RE(0); break;
    case 1: FLACB200_RE(4); break;
    case 2: FLACB200_RE(8); break;
    case 3: FLACB200_RE(12); break;
    default: FLACB200_RE(16); break;   // (cls <= 4: flags[1] is raised otherwise)
    }
#undef FLACB200_RE
}

bool restore_emit_ok(const DecCfg& cfg, const uint8_t* out)
{
    return cfg.pcm_kind <= 1 && cfg.nslots == cfg.channels && cfg.channels <= 2 && (cfg.bytes_per_sample == 2 || cfg.bytes_per_sample == 3) &&
           (reinterpret_cast<uintptr_t>(out) & 3) == 0 && (cfg.bstride & 3) == 0;
}

void launch_restore_emit(const DecCfg& cfg, const FrameCand* cands, uint32_t n, const SubRec* subs, const DecRec* recs, const unsigned long long* pos,
                         const int32_t* planes, uint8_t* out, const uint32_t* flags, bool need_clean, cudaStream_t st)
{
    const uint32_t bundles = (n + 31) / 32, grid = (bundles + RE_WARPS - 1) / RE_WARPS;
    const uint32_t nc = need_clean ? 1u : 0u;
    if (cfg.channels == 1 && cfg.bytes_per_sample == 2) count_launch(), k_restore_emit<1, 2><<<grid, 32 * RE_WARPS, 0, st>>>(cfg, cands, n, subs, recs, pos, planes, out, flags, nc);
    else if (cfg.channels == 1) count_launch(), k_restore_emit<1, 3><<<grid, 32 * RE_WARPS, 0, st>>>(cfg, cands, n, subs, recs, pos, planes, out, flags, nc);
    else if (cfg.bytes_per_sample == 2) count_launch(), k_restore_emit<2, 2><<<grid, 32 * RE_WARPS, 0, st>>>(cfg, cands, n, subs, recs, pos, planes, out, flags, nc);
    else count_launch(), k_restore_emit<2, 3><<<grid, 32 * RE_WARPS, 0, st>>>(cfg, cands, n, subs, recs, pos, planes, out, flags, nc);
}

void launch_parse(const DecCfg& cfg, const uint8_t* bytes, const DecSeg* segs, const FrameCand* cands, uint32_t n, int32_t* planes, SubRec* subs,
                  DecRec* recs, uint32_t* high_order, cudaStream_t st)
{
    // (148 SMs x 16 CTAs of 64 lanes fill the GPU; below a quarter of that the walk is latency-bound)
    if (n >= 37888u) count_launch(), k_parse<8><<<(n + PARSE_THREADS - 1) / PARSE_THREADS, PARSE_THREADS, 0, st>>>(cfg, bytes, segs, cands, n, planes, subs, recs, high_order);
    else count_launch(), k_parse<16><<<(n + PARSE_THREADS - 1) / PARSE_THREADS, PARSE_THREADS, 0, st>>>(cfg, bytes, segs, cands, n, planes, subs, recs, high_order);
}

void launch_restore(const DecCfg& cfg, const FrameCand* cands, uint32_t n, const SubRec* subs, const DecRec* recs, int32_t* planes, cudaStream_t st)
{
    const uint32_t threads = ((n + 31) / 32) * 32 * cfg.channels;
    count_launch(), k_restore<<<(threads + RESTORE_THREADS - 1) / RESTORE_THREADS, RESTORE_THREADS, 0, st>>>(cfg, cands, n, subs, recs, planes);
}

}   // namespace flacb200
