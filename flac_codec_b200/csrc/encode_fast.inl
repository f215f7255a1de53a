// encode_fast.inl -- register-tiled encode kernels for the common case (block <= 4096 samples, samples <= 28 bits,
// max LPC order <= 16): included by encode_kernels.cu inside namespace flacb200.
//
// One CTA of 256 threads owns one unit (a stereo frame with its L/R/M/S candidates, or one channel of a frame);
// every thread owns 16 consecutive samples in registers for the whole analysis, so the fixed differences, the
// LPC FIR (INT32 x INT16 -> INT64 MACs over a register window), the partition sums and the exact Rice bit counts
// never re-read memory.  PCM is unpacked straight from the caller's packed bytes with 128-bit loads: the int32
// candidate planes of the generic path (k_planes) never reach HBM.

constexpr int AN_THREADS = 256;
constexpr int AN_SPT = 16;                      // samples per thread
constexpr int AN_TILE = AN_THREADS * AN_SPT;    // 4096: largest block of the fast path
constexpr int AN_PAD = 32;                      // zero samples in front of every plane (history of the first thread)
constexpr int AN_STRIDE = AN_TILE + AN_PAD;

struct AnSmem {
    unsigned long long chunk_sum[2 * MAX_PARTS];   // two residual sets (fixed, LPC) are searched together
    unsigned long long acc[8];
    unsigned long long absum[4];
    uint32_t part_est[256];
    uint8_t part_code[256];
    uint32_t ord_est[16], ord_cnt[16], ord_ok[16];
    uint32_t orm[4];
    uint32_t flag, flag2[2];
    RiceChoice fixed, lpc;
    int16_t q[MAX_LPC];
};

// ---- packed PCM -> 16 consecutive inter-channel samples of C channels (Frame::fill_from_buf, src/audio.rs:149-187) ----
template <int C, int B>
__device__ inline void load16(const uint8_t* __restrict__ p, bool big_endian, int32_t* __restrict__ v)
{
    constexpr int NB = 16 * C * B, NW = NB / 4;
    uint32_t w[NW + 1];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < NW / 4; k++) {
            const uint4 t = reinterpret_cast<const uint4*>(p)[k];
            w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
        }
    } else if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = reinterpret_cast<const uint32_t*>(p)[k];
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++)
            w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
    }
    w[NW] = 0;
#pragma unroll
    for (int k = 0; k < 16 * C; k++) {
        const int bo = k * B;
        const uint32_t raw = __funnelshift_r(w[bo >> 2], w[(bo >> 2) + 1], (bo & 3) * 8);   // sample bytes in memory order
        if (big_endian) v[k] = (int32_t)__byte_perm(raw, 0, 0x0123) >> (32 - 8 * B);
        else v[k] = (int32_t)(raw << (32 - 8 * B)) >> (32 - 8 * B);
    }
}

template <int C>
__device__ inline void load16_any(const uint8_t* __restrict__ p, uint32_t bytes_per_sample, bool big_endian, int32_t* __restrict__ v)
{
    switch (bytes_per_sample) {
    case 1: load16<C, 1>(p, big_endian, v); break;
    case 2: load16<C, 2>(p, big_endian, v); break;
    case 3: load16<C, 3>(p, big_endian, v); break;
    default: load16<C, 4>(p, big_endian, v); break;
    }
}

// Fills the thread's 16 samples of up to two channels (ch0, ch0 + 1 when C == 2) of the block; samples past n are 0.
template <int C>
__device__ inline void load_thread_samples(const EncCfg& cfg, const FrameDesc& d, const uint8_t* __restrict__ pcm, uint32_t i0, uint32_t ch0,
                                           int32_t* __restrict__ a, int32_t* __restrict__ b)
{
    const bool full = i0 + AN_SPT <= d.n;
    if (full && cfg.pcm_kind != 3 && cfg.channels == (uint32_t)C) {
        int32_t v[16 * C];
        load16_any<C>(pcm + (d.pcm_off + i0) * (unsigned long long)(C * cfg.bytes_per_sample), cfg.bytes_per_sample, cfg.pcm_kind == 1, v);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            a[e] = v[e * C];
            if (C == 2) b[e] = v[e * C + 1];
        }
        return;
    }
    if (full && cfg.pcm_kind == 3) {
        load16<1, 4>(pcm + ((unsigned long long)ch0 * cfg.planar_stride + d.pcm_off + i0) * 4ull, false, a);
        if (C == 2) load16<1, 4>(pcm + ((unsigned long long)(ch0 + 1) * cfg.planar_stride + d.pcm_off + i0) * 4ull, false, b);
        return;
    }
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const bool ok = i0 + e < d.n;
        a[e] = ok ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch0) : 0;
        if (C == 2) b[e] = ok ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch0 + 1) : 0;
    }
}

__device__ inline void store16(int32_t* __restrict__ dst, const int32_t* __restrict__ v)
{
#pragma unroll
    for (int k = 0; k < 4; k++) reinterpret_cast<int4*>(dst)[k] = make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

// block accumulate: every thread contributes v to *slot (zeroed and synchronised by the caller)
__device__ inline void block_add(unsigned long long* slot, unsigned long long v)
{
    v = warp_sum_u64(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(slot, v);
}

// ---- best_partitions + try_reduce_rice + exact size, on residuals held in registers ----
// Searches NS (1 or 2) residual sets of the same block at once -- the fixed and the LPC residuals of a candidate --
// so that both share every barrier.  r[k][e] is the residual of sample i0 + e in set k; only samples in [o[k], n) count.

// rare: a thread's 16 samples straddle a boundary of the finest partition (block length not a multiple of 16 << p_max)
__device__ __noinline__ void chunk_sums_slow(const int32_t* r, uint32_t i0, uint32_t lo, uint32_t hi, uint32_t cf, unsigned long long* chunk_sum)
{
    for (uint32_t i = lo; i < hi; i++)
        if (r[i - i0]) atomicAdd(&chunk_sum[i / cf], (unsigned long long)uabs32(r[i - i0]));
}

// rare: partial thread, escaped partition, or more than one partition inside the thread's samples
__device__ __noinline__ void exact_bits_slow(const int32_t* r, uint32_t i0, uint32_t lo, uint32_t hi, uint32_t cp, uint32_t j0, const uint8_t* rice,
                                             unsigned long long* bits_out, uint32_t* bad_out)
{
    unsigned long long bits = 0;
    uint32_t bad = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = rice[i / cp - j0];
        const int32_t s = r[i - i0];
        if (c < 0x40) bits += (zigzag32(s) >> c) + 1u + c;
        else if (c & 0x40) {
            const uint32_t w = c & 31u;
            bits += w;
            if (s < -(1 << (w - 1)) || s > (1 << (w - 1)) - 1) bad = 1;   // write_signed_counted fails
        }
    }
    *bits_out = bits;
    *bad_out = bad;
}

template <int NS>
__device__ void rice_search_regs(const int32_t (*r)[AN_SPT], uint32_t i0, const uint32_t* o, uint32_t n, const EncCfg& cfg, AnSmem& sm,
                                 RiceChoice* const* outs)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t rice_max = cfg.use_rice2 ? 31u : 15u;
    uint32_t p_max = (uint32_t)__ffs((int)n) - 1u;
    if (p_max > cfg.max_porder) p_max = cfg.max_porder;
    if (p_max > MAX_PORDER) p_max = MAX_PORDER;
    const uint32_t cf = n >> p_max;   // finest chunk
    const bool cf_pow2 = (cf & (cf - 1)) == 0;
    const uint32_t cf_shift = 31u - (uint32_t)__clz((int)cf);
    if (tid < NS * MAX_PARTS) sm.chunk_sum[tid] = 0;
    if (tid < 2) { sm.acc[tid] = 0; sm.flag2[tid] = 0; }
    __syncthreads();
    uint32_t lo[NS], hi[NS];
    bool full[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        lo[k] = max(i0, o[k]);
        hi[k] = min(i0 + AN_SPT, n);
        full[k] = lo[k] == i0 && hi[k] == i0 + AN_SPT;   // all 16 samples are residuals
        if (lo[k] < hi[k]) {
            const uint32_t m_lo = cf_pow2 ? lo[k] >> cf_shift : lo[k] / cf, m_hi = cf_pow2 ? (hi[k] - 1) >> cf_shift : (hi[k] - 1) / cf;
            if (m_lo == m_hi) {
                unsigned long long sum = 0;
                if (full[k]) {
#pragma unroll
                    for (int e = 0; e < AN_SPT; e++) sum = acc_u32(sum, uabs32(r[k][e]));
                } else {
#pragma unroll
                    for (int e = 0; e < AN_SPT; e++)
                        if (i0 + e >= lo[k] && i0 + e < hi[k]) sum += uabs32(r[k][e]);
                }
                if (sum) atomicAdd(&sm.chunk_sum[k * MAX_PARTS + m_lo], sum);
            } else {
                int32_t tmp[AN_SPT];   // a copy: taking the address of r itself would push the register tile into local memory
#pragma unroll
                for (int e = 0; e < AN_SPT; e++) tmp[e] = r[k][e];
                chunk_sums_slow(tmp, i0, lo[k], hi[k], cf, sm.chunk_sum + k * MAX_PARTS);
            }
        }
    }
    __syncthreads();
    {   // threads 0..126 -> (order p, partition j) of set 0, threads 128..254 -> set 1
        const uint32_t k = tid >> 7, t = tid & 127;
        if (k < (uint32_t)NS && t < 127) {
            const uint32_t p = 31u - (uint32_t)__clz((int)(t + 1));
            const uint32_t j = t + 1 - (1u << p);
            uint8_t code = 0xFE;
            uint32_t est = 0;
            if (p <= p_max) {
                const uint32_t cp = n >> p;
                const uint32_t a = j * cp, b = a + cp;
                if (b > o[k]) {
                    const uint32_t span = 1u << (p_max - p);
                    const unsigned long long* cs = sm.chunk_sum + k * MAX_PARTS;
                    unsigned long long sum = 0;
                    for (uint32_t m = j * span; m < (j + 1) * span; m++) sum += cs[m];
                    code = partition_code(sum, b - max(a, o[k]), rice_max, &est);
                }
            }
            sm.part_code[tid] = code;
            sm.part_est[tid] = est;
        }
    }
    __syncthreads();
    {   // warp p sums the estimates of partition order p (both sets)
        const uint32_t p = tid >> 5, lane = tid & 31;
        if (p <= p_max) {
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const uint32_t base = k * 128 + (1u << p) - 1, cnt_all = 1u << p;
                uint32_t est = 0, cnt = 0, bad = 0;
                for (uint32_t j = lane; j < cnt_all; j += 32) {
                    const uint8_t c = sm.part_code[base + j];
                    if (c == 0xFE) continue;
                    if (c == 0xFF) bad = 1;
                    cnt++;
                    est += sm.part_est[base + j];
                }
                est = __reduce_add_sync(0xffffffffu, est);
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                bad = __reduce_or_sync(0xffffffffu, bad);
                if (lane == 0) {
                    sm.ord_est[k * 8 + p] = est;
                    sm.ord_cnt[k * 8 + p] = cnt;
                    sm.ord_ok[k * 8 + p] = (!bad && cnt != 0 && (cnt & (cnt - 1)) == 0) ? 1u : 0u;   // :3880-3881
                }
            }
        }
    }
    __syncthreads();
    uint32_t porder_g[NS], nparts[NS], method[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        RiceChoice& out = *outs[k];
        bool have = false;
        uint32_t best_est = 0, best_p = 0, best_count = 0;
        for (uint32_t p = 0; p <= p_max; p++) {
            if (!sm.ord_ok[k * 8 + p]) continue;
            const uint32_t est = sm.ord_est[k * 8 + p];
            if (!have || est < best_est) { have = true; best_est = est; best_p = p; best_count = sm.ord_cnt[k * 8 + p]; }   // first minimum :3885
        }
        if (!have) {   // unwrap_or_else (:3887): one partition escaped at 31 bits
            porder_g[k] = 0; nparts[k] = 1;
            method[k] = cfg.use_rice2 ? 1 : 0;
            if (tid == 0) { out.porder_g = 0; out.porder_w = 0; out.nparts = 1; out.rice[0] = 0x40 | 31; }
        } else {
            porder_g[k] = best_p; nparts[k] = best_count;
            const uint32_t base = k * 128 + (1u << best_p) - 1, j0 = (1u << best_p) - best_count;
            uint32_t big = 0;
            if (tid < best_count) {
                const uint8_t c = sm.part_code[base + j0 + tid];
                out.rice[tid] = c;
                big = (c < 0x40 && c >= 15) ? 1u : 0u;
            }
            if (__any_sync(0xffffffffu, big) && (tid & 31) == 0) atomicOr(&sm.flag2[k], 1u);
            if (tid == 0) {
                out.porder_g = (uint8_t)best_p;
                out.nparts = (uint8_t)best_count;
                out.porder_w = (uint8_t)(31u - (uint32_t)__clz((int)best_count));   // partitions.len().ilog2() :3902
            }
            method[k] = 2;   // resolved after the barrier
        }
    }
    __syncthreads();
    // exact size: what Partition::to_writer will emit (:3834-3863), plus the partition headers
#pragma unroll
    for (int k = 0; k < NS; k++) {
        RiceChoice& out = *outs[k];
        if (method[k] == 2) method[k] = (cfg.use_rice2 && (sm.flag2[k] & 1u)) ? 1 : 0;   // try_reduce_rice :3929-3942
        const uint32_t cp = n >> porder_g[k];
        const uint32_t j0 = (1u << porder_g[k]) - nparts[k];
        const bool cp_pow2 = (cp & (cp - 1)) == 0;
        const uint32_t cp_shift = 31u - (uint32_t)__clz((int)cp);
        unsigned long long bits = 0;
        uint32_t bad = 0;
        if (lo[k] < hi[k]) {
            const uint32_t ja = cp_pow2 ? lo[k] >> cp_shift : lo[k] / cp, jb = cp_pow2 ? (hi[k] - 1) >> cp_shift : (hi[k] - 1) / cp;
            const uint32_t c = out.rice[ja - j0];
            if (full[k] && ja == jb && c < 0x40) {   // the common case: 16 Rice codes with one parameter
#pragma unroll
                for (int e = 0; e < AN_SPT; e++) bits = acc_u32(bits, zigzag32(r[k][e]) >> c);
                bits += (unsigned long long)(AN_SPT * (1u + c));
            } else {
                int32_t tmp[AN_SPT];
#pragma unroll
                for (int e = 0; e < AN_SPT; e++) tmp[e] = r[k][e];
                exact_bits_slow(tmp, i0, lo[k], hi[k], cp, j0, out.rice, &bits, &bad);
            }
        }
        if (tid < nparts[k]) bits += (out.rice[tid] < 0x40) ? (method[k] ? 5u : 4u) : (method[k] ? 10u : 9u);   // partition headers
        block_add(&sm.acc[k], bits);
        if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) atomicOr(&sm.flag2[k], 2u);
    }
    __syncthreads();
    if (tid < (uint32_t)NS) {
        RiceChoice& out = *outs[tid];
        out.resid_bits = (uint32_t)sm.acc[tid] + 2 + 4;   // + coding method + partition order
        out.fail = (sm.flag2[tid] & 2u) ? 1u : 0u;
        out.method = (uint8_t)(tid == 0 ? method[0] : method[NS - 1]);
    }
    __syncthreads();
}

// One candidate channel: encode_subframe (src/encode.rs:2849-2980) without emitting bits.
// plane: the candidate's samples in shared memory (AN_PAD zeros in front); HB: LPC order rounded up to 4/8/12/16.
template <int HB>
__device__ void analyze_candidate(const EncCfg& cfg, AnSmem& sm, const int32_t* __restrict__ plane, uint32_t n, uint32_t full_bps, uint32_t mask,
                                  const LpcRec& lp, CandRec* __restrict__ rec)
{
    const uint32_t tid = threadIdx.x, i0 = tid * AN_SPT;
    const uint32_t wasted = (mask & 1u) ? 0u : (uint32_t)__ffs((int)mask) - 1u;   // :2878-2898
    const uint32_t bps = full_bps - wasted;
    const bool interior = i0 >= 4 && i0 + AN_SPT <= n;   // every sample of the thread counts everywhere
    // register window: w[HB + e] = x[i0 + e] >> wasted, w[HB - 1 - j] = x[i0 - 1 - j]
    int32_t w[HB + AN_SPT];
#pragma unroll
    for (int k = 0; k < (HB + AN_SPT) / 4; k++) {
        const int4 t = reinterpret_cast<const int4*>(plane + (int)i0 - HB)[k];
        w[4 * k] = t.x >> wasted; w[4 * k + 1] = t.y >> wasted; w[4 * k + 2] = t.z >> wasted; w[4 * k + 3] = t.w >> wasted;
    }
    // ---- LPC residuals first (encode_lpc_subframe :3090-3136, :3174-3203); lp.ok is uniform across the CTA ----
    int32_t r[2][AN_SPT];   // [0] fixed, [1] LPC
    bool lpc_ok = lp.ok != 0;
    if (tid < 8) sm.acc[tid] = 0;
    if (tid == 0) sm.flag = 0;
    if (lpc_ok && tid < MAX_LPC) sm.q[tid] = tid < lp.order ? lp.q[tid] : (int16_t)0;
    __syncthreads();
    // ---- encode_fixed_subframe (:3020-3088).  Samples are <= 28 bits wide here, so no difference up to order 4
    // can leave i32 (checked_sub never fails) and plain 32-bit arithmetic is exact. ----
    const uint32_t kmax = min(4u, n - 1);
    unsigned long long s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    const int32_t h1 = w[HB - 1] - w[HB - 2], h1b = w[HB - 2] - w[HB - 3], h1c = w[HB - 3] - w[HB - 4];
    const int32_t h2 = h1 - h1b, h2b = h1b - h1c, h3 = h2 - h2b;   // differences of the samples just before i0
    if (interior) {
        int32_t p1 = h1, p2 = h2, p3 = h3;
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) {
            const int32_t e1 = w[HB + e] - w[HB + e - 1], e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
            p1 = e1; p2 = e2; p3 = e3;
            s0 = acc_u32(s0, uabs32(w[HB + e])); s1 = acc_u32(s1, uabs32(e1)); s2 = acc_u32(s2, uabs32(e2));
            s3 = acc_u32(s3, uabs32(e3)); s4 = acc_u32(s4, uabs32(e4));
        }
    } else {
        int32_t p1 = h1, p2 = h2, p3 = h3;
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) {
            const int32_t e1 = w[HB + e] - w[HB + e - 1], e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
            p1 = e1; p2 = e2; p3 = e3;
            const uint32_t i = i0 + e;
            if (i >= kmax && i < n) {
                s0 += uabs32(w[HB + e]); s1 += uabs32(e1); s2 += uabs32(e2); s3 += uabs32(e3); s4 += uabs32(e4);
            }
        }
    }
    block_add(&sm.acc[2], s0); block_add(&sm.acc[3], s1); block_add(&sm.acc[4], s2); block_add(&sm.acc[5], s3); block_add(&sm.acc[6], s4);
    if (lpc_ok) {
        const uint32_t order = lp.order, shift = lp.shift;
        int32_t q[HB];
#pragma unroll
        for (int j = 0; j < HB; j++) q[j] = sm.q[j];
        uint32_t ovf = 0;
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) {
            long long sum = 0;
#pragma unroll
            for (int j = 0; j < HB; j++) sum = mad_wide_s32(w[HB + e - 1 - j], q[j], sum);   // :3187-3192
            const int32_t pred = (int32_t)(uint32_t)(unsigned long long)(sum >> shift);     // `as i32`
            const int32_t x = w[HB + e];
            const int32_t rr = (int32_t)((uint32_t)x - (uint32_t)pred);
            const uint32_t o1 = (uint32_t)((x ^ pred) & (x ^ rr));                          // sign bit: checked_sub overflowed
            if (interior && i0 >= order) ovf |= o1;
            else if (i0 + e >= order && i0 + e < n) ovf |= o1;
            r[1][e] = rr;
        }
        if (__any_sync(0xffffffffu, ovf >> 31) && (tid & 31) == 0) atomicOr(&sm.flag, 4u);   // ResidualOverflow
    }
    __syncthreads();
    s0 = sm.acc[2]; s1 = sm.acc[3]; s2 = sm.acc[4]; s3 = sm.acc[5]; s4 = sm.acc[6];
    if (sm.flag & 4u) lpc_ok = false;
    uint32_t fo = 0;   // first minimum among the orders that exist (:3065-3075)
    {
        unsigned long long best = s0;
        if (kmax >= 1 && s1 < best) { best = s1; fo = 1; }
        if (kmax >= 2 && s2 < best) { best = s2; fo = 2; }
        if (kmax >= 3 && s3 < best) { best = s3; fo = 3; }
        if (kmax >= 4 && s4 < best) { best = s4; fo = 4; }
    }
    // residuals of the chosen fixed order (fo is uniform across the CTA)
    if (fo == 0) {
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) r[0][e] = w[HB + e];
    } else if (fo == 1) {
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) r[0][e] = w[HB + e] - w[HB + e - 1];
    } else {
        int32_t p1 = h1, p2 = h2, p3 = h3;
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) {
            const int32_t e1 = w[HB + e] - w[HB + e - 1], e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
            p1 = e1; p2 = e2; p3 = e3;
            r[0][e] = fo == 2 ? e2 : fo == 3 ? e3 : e4;
        }
    }
    const uint32_t orders[2] = {fo, lp.order};
    RiceChoice* outs[2] = {&sm.fixed, &sm.lpc};
    if (lpc_ok) rice_search_regs<2>(r, i0, orders, n, cfg, sm, outs);
    else rice_search_regs<1>(r, i0, orders, n, cfg, sm, outs);
    const uint32_t hdr_bits = 8 + wasted;   // pad + type + wasted flag (+ unary(wasted - 1)) (src/stream.rs:1397)
    const bool fixed_ok = sm.fixed.fail == 0;
    const uint32_t fixed_bits = hdr_bits + fo * bps + sm.fixed.resid_bits;
    uint32_t lpc_bits = 0;
    if (lpc_ok) {
        if (sm.lpc.fail) lpc_ok = false;
        lpc_bits = hdr_bits + lp.order * bps + 4 + 5 + lp.order * lp.precision + sm.lpc.resid_bits;
    }
    // ---- choose (:2929-2979): fixed wins ties; VERBATIM unless strictly smaller ----
    const uint32_t verbatim_len = n * bps;
    int pick = -1;   // 0 fixed, 1 lpc
    if (fixed_ok && lpc_ok) pick = lpc_bits < fixed_bits ? 1 : 0;
    else if (fixed_ok) pick = 0;
    else if (lpc_ok) pick = 1;
    const uint32_t best_bits = pick == 1 ? lpc_bits : fixed_bits;
    if (pick >= 0 && !(best_bits < verbatim_len)) pick = -1;
    const RiceChoice& ch = pick == 1 ? sm.lpc : sm.fixed;
    if (tid == 0) {
        rec->wasted = (uint8_t)wasted;
        rec->bps = (uint8_t)bps;
        if (pick < 0) {
            rec->type = 1; rec->order = 0;
            rec->bits = hdr_bits + verbatim_len;
        } else {
            rec->type = pick == 1 ? 3 : 2;
            rec->order = pick == 1 ? lp.order : (uint8_t)fo;
            rec->precision = lp.precision; rec->shift = lp.shift;
            rec->method = ch.method; rec->porder_w = ch.porder_w; rec->porder_g = ch.porder_g; rec->nparts = ch.nparts;
            rec->bits = best_bits;
        }
    }
    if (pick >= 0) {
        if (tid < MAX_PARTS) rec->rice[tid] = ch.rice[tid];
        if (tid < MAX_LPC) rec->q[tid] = lp.q[tid];
    }
    __syncthreads();
}

// grid: frames (STEREO: L/R/M/S candidates of a two-channel frame) or frames * channels (one channel per CTA)
// dynamic smem: nplanes * AN_STRIDE int32
template <bool STEREO>
__global__ void __launch_bounds__(AN_THREADS, 2) k_analyze(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
                                                          const LpcRec* __restrict__ lpcs, CandRec* __restrict__ out,
                                                          unsigned long long* __restrict__ abssum)
{
    extern __shared__ __align__(16) int32_t an_planes[];
    __shared__ AnSmem sm;
    const uint32_t tid = threadIdx.x, i0 = tid * AN_SPT;
    const uint32_t f = STEREO ? blockIdx.x : blockIdx.x / cfg.channels;
    const uint32_t ch = STEREO ? 0 : blockIdx.x % cfg.channels;
    const FrameDesc d = descs[f];
    const uint32_t n = d.n;
    constexpr int NPL = STEREO ? 4 : 1;
    if (tid < AN_PAD) {
#pragma unroll
        for (int p = 0; p < NPL; p++) an_planes[p * AN_STRIDE + tid] = 0;
    }
    if (tid < 4) { sm.orm[tid] = 0; sm.absum[tid] = 0; }
    __syncthreads();
    int32_t a[AN_SPT], b[AN_SPT];
    if (STEREO) load_thread_samples<2>(cfg, d, pcm, i0, 0, a, b);
    else if (cfg.channels == 1) load_thread_samples<1>(cfg, d, pcm, i0, 0, a, b);
    else {
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) a[e] = i0 + e < n ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch) : 0;
    }
    if (i0 + AN_SPT > n) {
#pragma unroll
        for (int e = 0; e < AN_SPT; e++)
            if (i0 + e >= n) { a[e] = 0; if (STEREO) b[e] = 0; }
    }
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    store16(an_planes + AN_PAD + i0, a);
#pragma unroll
    for (int e = 0; e < AN_SPT; e++) o0 |= (uint32_t)a[e];
    if (STEREO) {
        store16(an_planes + AN_STRIDE + AN_PAD + i0, b);
        int32_t m[AN_SPT], s[AN_SPT];
        unsigned long long sl = 0, sr = 0, smid = 0, sside = 0;
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) {
            m[e] = (a[e] + b[e]) >> 1;   // :2721
            s[e] = a[e] - b[e];          // :2734
            o1 |= (uint32_t)b[e]; o2 |= (uint32_t)m[e]; o3 |= (uint32_t)s[e];
            sl += uabs32(a[e]); sr += uabs32(b[e]); smid += uabs32(m[e]); sside += uabs32(s[e]);
        }
        store16(an_planes + 2 * AN_STRIDE + AN_PAD + i0, m);
        store16(an_planes + 3 * AN_STRIDE + AN_PAD + i0, s);
        if (cfg.mode == MODE_FAST_MID_SIDE || cfg.mode == MODE_FAST_SIDE) {   // correlate_channels abs sums (:2475-2503)
            block_add(&sm.absum[0], sl); block_add(&sm.absum[1], sr); block_add(&sm.absum[2], smid); block_add(&sm.absum[3], sside);
        }
    }
    o0 = __reduce_or_sync(0xffffffffu, o0);
    if (STEREO) { o1 = __reduce_or_sync(0xffffffffu, o1); o2 = __reduce_or_sync(0xffffffffu, o2); o3 = __reduce_or_sync(0xffffffffu, o3); }
    if ((tid & 31) == 0) {
        if (o0) atomicOr(&sm.orm[0], o0);
        if (STEREO) { if (o1) atomicOr(&sm.orm[1], o1); if (o2) atomicOr(&sm.orm[2], o2); if (o3) atomicOr(&sm.orm[3], o3); }
    }
    __syncthreads();
    if (STEREO && tid < 4 && abssum) abssum[(size_t)f * 4 + tid] = sm.absum[tid];
    for (uint32_t slot = 0; slot < (uint32_t)NPL; slot++) {
        const uint32_t cand = STEREO ? f * 4 + slot : blockIdx.x;
        CandRec* rec = out + cand;
        if (STEREO && !slot_active(cfg, sm.absum, slot)) {
            if (tid == 0) { rec->type = 0xFF; rec->bits = 0; }
            continue;
        }
        const uint32_t full_bps = STEREO ? cand_bps(cfg, slot) : cfg.bps;
        const uint32_t mask = sm.orm[slot];
        if (mask == 0) {   // all samples zero -> CONSTANT (:2870, :2883)
            if (tid == 0) {
                rec->type = 0; rec->order = 0; rec->wasted = 0; rec->bps = (uint8_t)full_bps;
                rec->bits = 8 + full_bps;
            }
            continue;
        }
        const LpcRec lp = lpcs[cand];
        const int32_t* plane = an_planes + slot * AN_STRIDE + AN_PAD;
        const uint32_t hb = lp.ok ? (lp.order + 3u) >> 2 : 1u;
        switch (hb) {
        case 1: analyze_candidate<4>(cfg, sm, plane, n, full_bps, mask, lp, rec); break;
        case 2: analyze_candidate<8>(cfg, sm, plane, n, full_bps, mask, lp, rec); break;
        case 3: analyze_candidate<12>(cfg, sm, plane, n, full_bps, mask, lp, rec); break;
        default: analyze_candidate<16>(cfg, sm, plane, n, full_bps, mask, lp, rec); break;
        }
    }
}

bool analyze_fast_ok(const EncCfg& cfg)
{
    const uint32_t widest = cfg.bps + (cfg.mode != MODE_INDEPENDENT ? 1u : 0u);
    return cfg.block_size <= (uint32_t)AN_TILE && widest <= 28 && cfg.max_lpc_order <= 16;
}

cudaError_t launch_analyze(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const LpcRec* lpcs, CandRec* cands,
                           unsigned long long* abssum, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_analyze<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * AN_STRIDE * 4);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (cfg.mode != MODE_INDEPENDENT)
        k_analyze<true><<<cfg.nframes, AN_THREADS, 4 * AN_STRIDE * 4, st>>>(cfg, descs, pcm, lpcs, cands, abssum);
    else
        k_analyze<false><<<cfg.nframes * cfg.channels, AN_THREADS, AN_STRIDE * 4, st>>>(cfg, descs, pcm, lpcs, cands, abssum);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_lpc2: LpcParameters::best (src/encode.rs:3292-3332) straight from the packed PCM.
//
// The reference's autocorrelation is a strict left-to-right f64 sum per lag (:3491-3497); to stay bit-identical
// each lag is one sequential chain of separately rounded multiplies and adds.  A lane owns FOUR consecutive lags
// of one candidate: the four samples x[i + 4m .. i + 4m + 3] it needs slide through registers, so a step costs one
// 64-bit shared load for the new sample, one broadcast load for x[i], and four independent DMUL + DADD chains.
// ceil((M + 1) / 4) lanes make a candidate and a warp runs 32 / that many candidates (8 for M <= 15, i.e. two
// stereo frames) from per-candidate rings of windowed samples (four 32-sample tiles + two mirror tiles, so that
// inside a tile every shared address is `lane base + immediate`; the rings are skewed by one double so that the
// 32 lanes of a load spread over all banks).  Wasted bits are assumed 0 while the OR mask is gathered on the fly;
// the rare candidate with wasted bits is run again with the shift applied.
// ------------------------------------------------------------------------------------------------
constexpr int L2_WARPS = 4;
constexpr int L2_RING = 192;      // 4 tiles + mirrors of tiles 0 and 1
constexpr int L2_MAXC = 10;       // candidates per warp, at most

struct Lpc2Cand {
    double* ring;   // the ring is dead once R[] is known: R, err, bits and the coefficient sets reuse its space
    double* R;      // M + 1
    double* err;    // M
    double* bits;   // M
    double* sets;   // M (M + 1) / 2: coefficient set of every order, triangular
};
__host__ __device__ inline uint32_t lpc2_cand_doubles(uint32_t M)
{
    const uint32_t scratch = (M + 1) + M + M + M * (M + 1) / 2;
    return (scratch > (uint32_t)L2_RING ? scratch : (uint32_t)L2_RING) + 1;   // + 1: bank skew between candidates
}
__device__ inline Lpc2Cand lpc2_cand(double* base, uint32_t M)
{
    Lpc2Cand c;
    c.ring = base;
    c.R = base;
    c.err = c.R + (M + 1);
    c.bits = c.err + M;
    c.sets = c.bits + M;
    return c;
}

__host__ __device__ inline uint32_t lpc2_lanes_per_cand(uint32_t M) { return (M + 1 + 3) / 4; }
__host__ __device__ inline uint32_t lpc2_cpw(uint32_t M, uint32_t nslots, bool stereo4)
{
    uint32_t cpw = 32u / lpc2_lanes_per_cand(M);
    if (cpw > (uint32_t)L2_MAXC) cpw = L2_MAXC;
    if (stereo4 && cpw >= 4) cpw &= ~3u;   // whole stereo frames (L, R, M, S) per warp: the loader reads L/R once
    (void)nslots;
    return cpw ? cpw : 1u;
}

// dynamic smem: L2_WARPS * cpw * lpc2_cand_doubles(M) doubles
__global__ void __launch_bounds__(32 * L2_WARPS) k_lpc2(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
                                                       const double* __restrict__ winpool, LpcRec* __restrict__ out, uint32_t cpw, uint32_t ncand)
{
    extern __shared__ __align__(16) uint8_t l2_dyn[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t M = cfg.max_lpc_order;
    const uint32_t c0 = (blockIdx.x * L2_WARPS + wid) * cpw;   // first candidate of this warp
    if (c0 >= ncand) return;
    const uint32_t ncs = min(cpw, ncand - c0);
    const uint32_t cdoubles = lpc2_cand_doubles(M);
    double* wbase = reinterpret_cast<double*>(l2_dyn) + (size_t)wid * cpw * cdoubles;
    const uint32_t LPCL = lpc2_lanes_per_cand(M);
    uint32_t g = lane / LPCL, m = lane % LPCL;
    const bool live = g < ncs;
    if (!live) { g = 0; m = 0; }
    const bool stereo4 = cfg.mode != MODE_INDEPENDENT && (cpw & 3u) == 0;   // whole L/R/M/S frames per warp
    if (lane < ncs) out[c0 + lane].ok = 0;
    // per-candidate block lengths; the warp iterates over the longest
    uint32_t nmax = 0;
    for (uint32_t c = 0; c < ncs; c++) nmax = max(nmax, descs[(c0 + c) / cfg.nslots].n);
    const uint32_t ntiles = (nmax + 31) / 32;
    uint32_t shift_mask = 0;   // bit c: candidate c must be redone with its wasted bits shifted out
    uint32_t wasted_of[L2_MAXC], masks[L2_MAXC];
#pragma unroll
    for (int c = 0; c < L2_MAXC; c++) { wasted_of[c] = 0; masks[c] = 0; }
    double acc0 = -0.0, acc1 = -0.0, acc2 = -0.0, acc3 = -0.0;   // Iterator::sum::<f64>() folds from -0.0
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && shift_mask == 0) break;
        int32_t px[L2_MAXC];   // samples of the tile being prefetched, one per candidate
        double pw[L2_MAXC];    // and their window values
        auto fetch_tile = [&](uint32_t tile) {   // global loads only: nothing here waits for them
            const uint32_t idx = tile * 32 + lane;
            if (stereo4) {
#pragma unroll
                for (int fi = 0; fi < L2_MAXC / 4; fi++) {
                    const uint32_t c = fi * 4;
                    int32_t l = 0, r = 0;
                    double wv = 0.0;
                    if (c < ncs) {
                        const FrameDesc d = descs[(c0 + c) >> 2];
                        if (idx < d.n && d.n > M) {
                            l = load_pcm_sample(pcm, cfg, d.pcm_off + idx, 0);
                            r = load_pcm_sample(pcm, cfg, d.pcm_off + idx, 1);
                            wv = winpool[d.win_off + idx];
                        }
                    }
                    px[c] = l; px[c + 1] = r; px[c + 2] = (l + r) >> 1; px[c + 3] = l - r;
                    pw[c] = pw[c + 1] = pw[c + 2] = pw[c + 3] = wv;
                }
#pragma unroll
                for (int c = (L2_MAXC / 4) * 4; c < L2_MAXC; c++) { px[c] = 0; pw[c] = 0.0; }
            } else {
#pragma unroll
                for (int c = 0; c < L2_MAXC; c++) {
                    px[c] = 0;
                    pw[c] = 0.0;
                    if ((uint32_t)c < ncs) {
                        const uint32_t cand = c0 + c;
                        const FrameDesc d = descs[cand / cfg.nslots];
                        if (idx < d.n && d.n > M) {
                            const uint32_t slot = cand % cfg.nslots;
                            if (cfg.mode == MODE_INDEPENDENT) px[c] = load_pcm_sample(pcm, cfg, d.pcm_off + idx, slot);
                            else {
                                const int32_t l = load_pcm_sample(pcm, cfg, d.pcm_off + idx, 0), r = load_pcm_sample(pcm, cfg, d.pcm_off + idx, 1);
                                px[c] = slot == 0 ? l : slot == 1 ? r : slot == 2 ? ((l + r) >> 1) : (l - r);
                            }
                            pw[c] = winpool[d.win_off + idx];
                        }
                    }
                }
            }
        };
        auto store_tile = [&](uint32_t tile) {
            const uint32_t pos = (tile & 3) * 32 + lane;
#pragma unroll
            for (int c = 0; c < L2_MAXC; c++) {
                if ((uint32_t)c >= ncs || (pass == 1 && !((shift_mask >> c) & 1u))) continue;
                if (pass == 0) masks[c] |= (uint32_t)px[c];
                const double v = __dmul_rn((double)(px[c] >> wasted_of[c]), pw[c]);   // Window::apply :1799
                double* ring = wbase + (size_t)c * cdoubles;
                ring[pos] = v;
                if ((tile & 3) < 2) ring[128 + pos] = v;   // mirrors of tiles 0 and 1 (mod 4)
            }
        };
        for (uint32_t t0 = 0; t0 < 3; t0++) {
            fetch_tile(t0);
            store_tile(t0);
        }
        __syncwarp();
        acc0 = acc1 = acc2 = acc3 = -0.0;
        const double* ring = wbase + (size_t)g * cdoubles;
        for (uint32_t t = 0; t < ntiles; t++) {
            fetch_tile(t + 3);   // in flight while the 32 steps below run
            const double* pa = ring + (t & 3) * 32;
            const double* pb = pa + 4 * m;
            double b0 = pb[0], b1 = pb[1], b2 = pb[2];
#pragma unroll
            for (int s = 0; s < 32; s++) {   // autocorrelate :3491-3497, four lags per lane
                const double a = pa[s], b3 = pb[s + 3];
                acc0 = __dadd_rn(acc0, __dmul_rn(a, b0));
                acc1 = __dadd_rn(acc1, __dmul_rn(a, b1));
                acc2 = __dadd_rn(acc2, __dmul_rn(a, b2));
                acc3 = __dadd_rn(acc3, __dmul_rn(a, b3));
                b0 = b1; b1 = b2; b2 = b3;
            }
            __syncwarp();
            store_tile(t + 3);   // replaces tile t - 1; tiles t + 1 and t + 2 stay resident
            __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < L2_MAXC; c++) masks[c] = __reduce_or_sync(0xffffffffu, masks[c]);
        // the rings are dead now: R[] goes on top of them
        const bool mine_redo = (shift_mask >> g) & 1u;
        if (live && (pass == 0 || mine_redo)) {
            double* Rg = wbase + (size_t)g * cdoubles;
            if (4 * m + 0 <= M) Rg[4 * m + 0] = acc0;
            if (4 * m + 1 <= M) Rg[4 * m + 1] = acc1;
            if (4 * m + 2 <= M) Rg[4 * m + 2] = acc2;
            if (4 * m + 3 <= M) Rg[4 * m + 3] = acc3;
        }
        if (pass == 0) {
#pragma unroll
            for (int c = 0; c < L2_MAXC; c++) {
                const uint32_t mk = masks[c];
                wasted_of[c] = (mk == 0 || (mk & 1u)) ? 0u : (uint32_t)__ffs((int)mk) - 1u;   // :2878-2898
                if (wasted_of[c]) shift_mask |= 1u << c;
            }
        }
        __syncwarp();
    }
    // a second pass rebuilds only the rings of the shifted candidates (store_tile skips the others), so the R[] that the
    // untouched candidates parked on top of their own rings is still intact here
    __syncwarp();
    // ---- per candidate: Levinson-Durbin on the group's first lane, order estimate on all its lanes ----
    const Lpc2Cand my = lpc2_cand(wbase + (size_t)g * cdoubles, M);
    uint32_t mask = 0, wasted = 0;
#pragma unroll
    for (int c = 0; c < L2_MAXC; c++)
        if ((uint32_t)c == g) { mask = masks[c]; wasted = wasted_of[c]; }
    const uint32_t cand = c0 + g;
    const uint32_t n = descs[cand / cfg.nslots].n;
    const bool run = live && mask != 0 && n > M;   // all-zero candidates become CONSTANT (:2883); n <= M: InsufficientLpcSamples (:3300)
    const uint32_t bps = cand_bps(cfg, cand % cfg.nslots) - wasted;
    const uint32_t precision = lpc_precision_for(n);
    if (run && m == 0) {   // lp_coefficients (:3536-3580), every order's set kept
        const double* R = my.R;
        double* a = my.sets;   // order 1
        double k = __ddiv_rn(R[1], R[0]);
        a[0] = k;
        my.err[0] = __dmul_rn(R[0], __dsub_rn(1.0, __dmul_rn(k, k)));
        for (uint32_t i = 1; i < M; i++) {
            double* b = a + i;   // the set of order i + 1 follows the i entries of order i
            double s = -0.0;
            for (uint32_t j = 0; j < i; j++) s = __dadd_rn(s, __dmul_rn(R[i - j], a[j]));
            const double q = __dsub_rn(R[i + 1], s);
            k = __ddiv_rn(q, my.err[i - 1]);
            for (uint32_t j = 0; j < i; j++) b[j] = __dsub_rn(a[j], __dmul_rn(k, a[i - 1 - j]));
            b[i] = k;
            my.err[i] = __dmul_rn(my.err[i - 1], __dsub_rn(1.0, __dmul_rn(k, k)));
            a = b;
        }
    }
    __syncwarp();
    if (run) {   // subframe_bits_by_order (:3656-3686): this lane's orders 4m + 1 .. 4m + 4
        const double error_scale = __ddiv_rn(0.5, (double)n);
        const double divisor = 2.0 * 0.693147180559945309417232121458176568;
        for (uint32_t o = 4 * m + 1; o <= min(4 * m + 4, M); o++) {
            const double bpr = __ddiv_rn(log(__dmul_rn(my.err[o - 1], error_scale)), divisor);
            my.bits[o - 1] = fma(bpr, (double)(n - o), (double)(o * (bps + precision)));
        }
    }
    __syncwarp();
    if (run && m == 0) {
        int best = 0;   // compute_best_order (:3688-3702): take_while(err > 0), first minimum under total_cmp
        double best_bits = 0.0;
        for (uint32_t o = 1; o <= M; o++) {
            if (!(my.err[o - 1] > 0.0)) break;
            const double b = my.bits[o - 1];
            if (best == 0 || total_key(b) < total_key(best_bits)) { best = (int)o; best_bits = b; }
        }
        if (best != 0) {
            const double* cur = my.sets + (size_t)best * (best - 1) / 2;
            // quantize (:3334-3401)
            double l = fabs(cur[0]);
            for (int j = 1; j < best; j++) {
                const double a = fabs(cur[j]);
                if (total_key(a) >= total_key(l)) l = a;
            }
            if (l > 0.0) {
                const int32_t max_coeff = (1 << (precision - 1)) - 1, min_coeff = -(1 << (precision - 1));
                const int32_t lg = f64_as_i32_sat(floor(log2(l)));
                long long sh = (long long)((int32_t)precision - 1) - (long long)lg - 1;   // :3360
                if (sh > 15) sh = 15;
                if (sh >= -16) {
                    LpcRec rec;
                    double error = 0.0;
                    if (sh >= 0) {
                        const double scale = (double)(1 << sh);
                        for (int j = 0; j < best; j++) {
                            const double sum = fma(cur[j], scale, error);   // mul_add :3372
                            int32_t q = f64_as_i32_sat(round(sum));
                            q = q < min_coeff ? min_coeff : (q > max_coeff ? max_coeff : q);
                            error = __dsub_rn(sum, (double)q);
                            rec.q[j] = (int16_t)q;
                        }
                        rec.shift = (uint8_t)sh;
                    } else {
                        const double scale = (double)(1 << (-sh));
                        for (int j = 0; j < best; j++) {
                            const double sum = __dadd_rn(__ddiv_rn(cur[j], scale), error);   // :3391
                            int32_t q = f64_as_i32_sat(round(sum));
                            q = q < min_coeff ? min_coeff : (q > max_coeff ? max_coeff : q);
                            error = __dsub_rn(sum, (double)q);
                            rec.q[j] = (int16_t)q;
                        }
                        rec.shift = 0;
                    }
                    for (int j = best; j < MAX_LPC; j++) rec.q[j] = 0;
                    rec.ok = 1;
                    rec.order = (uint8_t)best;
                    rec.precision = (uint8_t)precision;
                    rec.pad = 0;
                    out[cand] = rec;
                }
            }
        }
    }
}

cudaError_t launch_lpc2(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const double* winpool, LpcRec* lpcs, cudaStream_t st)
{
    const uint32_t cpw = lpc2_cpw(cfg.max_lpc_order, cfg.nslots, cfg.mode != MODE_INDEPENDENT);
    const uint32_t ncand = cfg.nframes * cfg.nslots;
    const uint32_t nwarps = (ncand + cpw - 1) / cpw;
    const size_t smem = (size_t)L2_WARPS * cpw * lpc2_cand_doubles(cfg.max_lpc_order) * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_lpc2, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k_lpc2<<<(nwarps + L2_WARPS - 1) / L2_WARPS, 32 * L2_WARPS, smem, st>>>(cfg, descs, pcm, winpool, lpcs, cpw, ncand);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_pack2: bit emission (src/encode.rs:2982-3136, :3834-3863) for blocks of the fast path.
//
// CTA per emitted subframe, 16 consecutive samples per thread.  A thread's 16 codes are contiguous in the
// bitstream: it assembles them in a register word by word, stores the words it owns outright and ORs only the
// two words it shares with its neighbours.  The subframe is built in shared memory and copied out with
// coalesced stores (byte-swapped), the two boundary words of the subframe with atomicOr.
// ------------------------------------------------------------------------------------------------
struct BitSink {
    uint32_t* words;
    uint32_t first_word, last_word;   // words shared with the neighbouring threads
    uint32_t cur_idx, cur;
    __device__ inline void flush()
    {
        if (cur == 0) return;
        if (cur_idx == first_word || cur_idx == last_word) atomicOr(words + cur_idx, cur);
        else words[cur_idx] = cur;
    }
    // OR the low nbits (1..32) of v at bit position q (relative to words[0]); positions never decrease
    __device__ inline void put(uint32_t q, uint32_t nbits, uint32_t v)
    {
        if (nbits < 32) v &= (1u << nbits) - 1u;
        const uint32_t wi = q >> 5, off = q & 31;
        const unsigned long long wide = ((unsigned long long)v) << (64 - nbits - off);
        const uint32_t hi = (uint32_t)(wide >> 32), lo = (uint32_t)wide;
        if (wi != cur_idx) {
            flush();
            cur_idx = wi;
            cur = 0;
        }
        cur |= hi;
        if (off + nbits > 32) {
            flush();
            cur_idx = wi + 1;
            cur = lo;
        }
    }
};

// grid F * nsub_max, block AN_THREADS; dynamic smem: cap_words uint32
template <bool STEREO>
__global__ void __launch_bounds__(AN_THREADS) k_pack2(EncCfg cfg, uint32_t nsub_max, uint32_t cap_words, const FrameDesc* __restrict__ descs,
                                                     const uint8_t* __restrict__ pcm, const CandRec* __restrict__ cands,
                                                     const FrameRec* __restrict__ frecs, uint8_t* __restrict__ out)
{
    extern __shared__ __align__(16) uint32_t p2_words[];
    __shared__ uint32_t warp_tot[AN_THREADS / 32];
    __shared__ CandRec cr;
    const uint32_t f = blockIdx.x / nsub_max, c = blockIdx.x % nsub_max, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const FrameRec& fr = frecs[f];
    if (c >= fr.nsub) return;
    const uint32_t slot = fr.slot[c];
    const uint32_t cand = f * cfg.nslots + slot;
    if (tid < sizeof(CandRec) / 4) reinterpret_cast<uint32_t*>(&cr)[tid] = reinterpret_cast<const uint32_t*>(cands + cand)[tid];
    const FrameDesc d = descs[f];
    const uint32_t n = d.n, i0 = tid * AN_SPT;
    // bit range of this CTA inside the output: subframe c (plus the frame header for c == 0)
    const unsigned long long frame_bit0 = fr.out_off * 8ull;
    const unsigned long long g0 = frame_bit0 + (c == 0 ? 0u : fr.sub_bit[c]);
    __syncthreads();
    const unsigned long long g1 = frame_bit0 + fr.sub_bit[c] + cr.bits;
    const unsigned long long w0 = g0 >> 5, w1 = (g1 + 31) >> 5;
    const uint32_t nwords = (uint32_t)(w1 - w0);
    const uint32_t origin_sub = (uint32_t)(frame_bit0 + fr.sub_bit[c] - (w0 << 5));   // subframe start relative to words[0]
    for (uint32_t i = tid; i < nwords + 1 && i < cap_words; i += AN_THREADS) p2_words[i] = 0;
    const uint32_t wasted = cr.wasted, bps = cr.bps, type = cr.type, order = (type >= 2) ? cr.order : 0;
    // ---- this thread's samples and the 16 before them ----
    int32_t own[AN_SPT], prev[AN_SPT];
    {
        int32_t a[AN_SPT], b[AN_SPT], pa[AN_SPT], pb[AN_SPT];
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) { a[e] = b[e] = pa[e] = pb[e] = 0; }
        if (STEREO) {
            if (i0 < n) load_thread_samples<2>(cfg, d, pcm, i0, 0, a, b);
            if (tid > 0 && i0 - AN_SPT < n) load_thread_samples<2>(cfg, d, pcm, i0 - AN_SPT, 0, pa, pb);
#pragma unroll
            for (int e = 0; e < AN_SPT; e++) {
                own[e] = slot == 0 ? a[e] : slot == 1 ? b[e] : slot == 2 ? ((a[e] + b[e]) >> 1) : (a[e] - b[e]);
                prev[e] = slot == 0 ? pa[e] : slot == 1 ? pb[e] : slot == 2 ? ((pa[e] + pb[e]) >> 1) : (pa[e] - pb[e]);
            }
        } else if (cfg.channels == 1) {
            if (i0 < n) load_thread_samples<1>(cfg, d, pcm, i0, 0, own, b);
            if (tid > 0 && i0 - AN_SPT < n) load_thread_samples<1>(cfg, d, pcm, i0 - AN_SPT, 0, prev, pb);
            else {
#pragma unroll
                for (int e = 0; e < AN_SPT; e++) prev[e] = 0;
            }
        } else {
#pragma unroll
            for (int e = 0; e < AN_SPT; e++) {
                own[e] = i0 + e < n ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, slot) : 0;
                prev[e] = (tid > 0 && i0 - AN_SPT + e < n) ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 - AN_SPT + e, slot) : 0;
            }
        }
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) { own[e] >>= wasted; prev[e] >>= wasted; }   // :2891
    }
    __syncthreads();
    auto put0 = [&](uint32_t q, uint32_t nbits, uint32_t v) { put_bits<false>(p2_words, q, nbits, v); };   // shared-memory atomics
    uint32_t pos = origin_sub;
    if (tid == 0) {
        if (c == 0)
            for (uint32_t i = 0; i < fr.hdr_len; i++) put0((uint32_t)(frame_bit0 - (w0 << 5)) + 8 * i, 8, fr.hdr[i]);
        // SubframeHeader (src/stream.rs:1397-1413): pad, 6-bit type, wasted flag, unary(wasted - 1)
        const uint32_t code = type == 0 ? 0u : type == 1 ? 1u : type == 2 ? 8u + order : 31u + order;
        put0(pos, 8, (code << 1) | (wasted ? 1u : 0u));
        if (wasted) put0(pos + 8 + (wasted - 1), 1, 1);
    }
    pos += 8 + wasted;
    BitSink sink;
    sink.words = p2_words;
    sink.cur = 0;
    sink.cur_idx = 0xFFFFFFFFu;
    if (type == 0) {   // CONSTANT: the sample is zero by construction (:2870-2887)
        if (tid == 0) put0(pos, bps, (uint32_t)own[0]);
    } else if (type == 1) {   // VERBATIM (:3000-3018)
        const uint32_t lo = min(i0, n), hi = min(i0 + AN_SPT, n);
        if (lo < hi) {
            const uint32_t p = pos + lo * bps, t = (hi - lo) * bps;
            sink.first_word = p >> 5;
            sink.last_word = (p + t - 1) >> 5;
#pragma unroll
            for (int e = 0; e < AN_SPT; e++)
                if (i0 + e < n) sink.put(pos + (i0 + e) * bps, bps, (uint32_t)own[e]);
            sink.flush();
        }
    } else {
        if (tid == 0) {
            for (uint32_t i = 0; i < order; i++) put0(pos + i * bps, bps, (uint32_t)own[i]);   // warm-up (:3083, :3118); order <= 16
        }
        pos += order * bps;
        if (type == 3) {
            const uint32_t prec = cr.precision;
            if (tid == 0) {
                put0(pos, 4, prec - 1);       // :3122
                put0(pos + 4, 5, cr.shift);   // :3129
                for (uint32_t j = 0; j < order; j++) put0(pos + 9 + j * prec, prec, (uint32_t)(int32_t)cr.q[j]);   // :3131
            }
            pos += 9 + order * prec;
        }
        if (tid == 0) {   // residual block header (:3944-3961)
            put0(pos, 2, cr.method);
            put0(pos + 2, 4, cr.porder_w);
        }
        pos += 6;
        // ---- residuals of this thread's samples ----
        int32_t r[AN_SPT];
        if (type == 2) {
            int32_t x1 = prev[15], x2 = prev[14], x3 = prev[13], x4 = prev[12];
#pragma unroll
            for (int e = 0; e < AN_SPT; e++) {
                const int32_t x0 = own[e];
                r[e] = order == 0 ? x0 : order == 1 ? x0 - x1 : order == 2 ? x0 - 2 * x1 + x2 : order == 3 ? x0 - 3 * x1 + 3 * x2 - x3
                                                                                                           : x0 - 4 * x1 + 6 * x2 - 4 * x3 + x4;
                x4 = x3; x3 = x2; x2 = x1; x1 = x0;
            }
        } else {
            const uint32_t shift = cr.shift;
            int32_t w[2 * AN_SPT];
#pragma unroll
            for (int e = 0; e < AN_SPT; e++) { w[e] = prev[e]; w[AN_SPT + e] = own[e]; }
            auto fir = [&](auto hb_tag) {
                constexpr int HB = decltype(hb_tag)::value;
                int32_t q[HB];
#pragma unroll
                for (int j = 0; j < HB; j++) q[j] = (uint32_t)j < order ? (int32_t)cr.q[j] : 0;
#pragma unroll
                for (int e = 0; e < AN_SPT; e++) {
                    long long sum = 0;
#pragma unroll
                    for (int j = 0; j < HB; j++) sum = mad_wide_s32(w[AN_SPT + e - 1 - j], q[j], sum);
                    r[e] = (int32_t)((uint32_t)w[AN_SPT + e] - (uint32_t)(unsigned long long)(sum >> shift));
                }
            };
            switch ((order + 3) >> 2) {
            case 1: fir(std::integral_constant<int, 4>{}); break;
            case 2: fir(std::integral_constant<int, 8>{}); break;
            case 3: fir(std::integral_constant<int, 12>{}); break;
            default: fir(std::integral_constant<int, 16>{}); break;
            }
        }
        // ---- code lengths (partition headers ride on the first residual of each partition) ----
        const uint32_t cp = n >> cr.porder_g;
        const uint32_t j0 = (1u << cr.porder_g) - cr.nparts;
        const bool cp_pow2 = (cp & (cp - 1)) == 0;
        const uint32_t cp_shift = 31u - (uint32_t)__clz((int)cp);
        const uint32_t hb = cr.method ? 5u : 4u;
        const uint32_t escape_code = cr.method ? 31u : 15u;
        const uint32_t lo = max(i0, order), hi = min(i0 + AN_SPT, n);
        uint32_t len[AN_SPT], code[AN_SPT];
        uint32_t firsts = 0, tsum = 0;
#pragma unroll
        for (int e = 0; e < AN_SPT; e++) {
            const uint32_t i = i0 + e;
            len[e] = 0;
            code[e] = 0;
            if (i >= lo && i < hi) {
                const uint32_t pj = cp_pow2 ? (i >> cp_shift) : (i / cp);
                const uint32_t cc = cr.rice[pj - j0];
                code[e] = cc;
                const uint32_t pstart = pj * cp;
                const bool first = i == (pstart > order ? pstart : order);
                uint32_t l = 0;
                if (cc < 0x40) l = (zigzag32(r[e]) >> cc) + 1u + cc;
                else if (cc & 0x40) l = cc & 31u;
                if (first) { l += (cc < 0x40) ? hb : hb + 5; firsts |= 1u << e; }
                len[e] = l;
                tsum += l;
            }
        }
        // block exclusive scan of tsum
        uint32_t incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t k = 0; k < wid; k++) before += warp_tot[k];
        uint32_t p = pos + before + incl - tsum;
        if (tsum) {
            sink.first_word = p >> 5;
            sink.last_word = (p + tsum - 1) >> 5;
#pragma unroll
            for (int e = 0; e < AN_SPT; e++) {
                if (len[e] == 0) continue;
                const uint32_t cc = code[e];
                uint32_t q = p;
                if (firsts & (1u << e)) {   // ResidualPartitionHeader::to_writer (src/stream.rs:1603-1619)
                    if (cc < 0x40) { sink.put(q, hb, cc); q += hb; }
                    else { sink.put(q, hb, escape_code); sink.put(q + hb, 5, (cc & 0x40) ? (cc & 31u) : 0u); q += hb + 5; }
                }
                if (cc < 0x40) {
                    const uint32_t u = zigzag32(r[e]);
                    sink.put(q + (u >> cc), cc + 1, (1u << cc) | (u & ((1u << cc) - 1u)));   // unary stop bit + cc LSBs (:3850-3851)
                } else if (cc & 0x40) {
                    sink.put(q, cc & 31u, (uint32_t)r[e]);   // escaped: raw two's complement (:3857)
                }
                p += len[e];
            }
            sink.flush();
        }
    }
    __syncthreads();
    // interior words belong to this CTA alone: plain stores; the two boundary words are shared with the neighbours
    uint32_t* gwords = reinterpret_cast<uint32_t*>(out);
    for (uint32_t i = tid; i < nwords; i += AN_THREADS) {
        const uint32_t v = __byte_perm(p2_words[i], 0, 0x0123);
        if (i == 0 || i == nwords - 1) { if (v) atomicOr(gwords + w0 + i, v); }
        else gwords[w0 + i] = v;
    }
}

// CRC-16 of every frame, warp per frame, coalesced word loads (src/crc.rs:144-188, src/encode.rs:2408-2409)
__global__ void __launch_bounds__(256) k_crc16w(const FrameRec* __restrict__ frecs, uint32_t nframes, uint8_t* __restrict__ out)
{
    __shared__ Crc16Tables tabs;
    crc16_tables_init(tabs);
    __syncthreads();
    const uint32_t f = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (f >= nframes) return;
    const FrameRec& fr = frecs[f];
    const uint32_t total = fr.frame_bytes - 2;
    const uint32_t crc = crc16_warp(tabs, out, fr.out_off, total);
    if ((threadIdx.x & 31) == 0) {
        out[fr.out_off + total] = (uint8_t)(crc >> 8);
        out[fr.out_off + total + 1] = (uint8_t)crc;
    }
}

cudaError_t launch_pack2_crc(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const CandRec* cands, const FrameRec* frecs,
                             uint8_t* out, cudaStream_t st)
{
    const uint32_t nsub_max = cfg.mode == MODE_INDEPENDENT ? cfg.channels : 2;
    const uint32_t cap_words = pack_cap_words(cfg);
    const size_t smem = (size_t)cap_words * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_pack2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pack2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (cfg.mode != MODE_INDEPENDENT)
        k_pack2<true><<<cfg.nframes * nsub_max, AN_THREADS, smem, st>>>(cfg, nsub_max, cap_words, descs, pcm, cands, frecs, out);
    else
        k_pack2<false><<<cfg.nframes * nsub_max, AN_THREADS, smem, st>>>(cfg, nsub_max, cap_words, descs, pcm, cands, frecs, out);
    k_crc16w<<<(cfg.nframes + 7) / 8, 256, 0, st>>>(frecs, cfg.nframes, out);
    return cudaGetLastError();
}
