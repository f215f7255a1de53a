// encode_fast.inl -- register-tiled encode kernels for the common case (block <= 4096 samples, samples <= 28 bits,
// max LPC order <= 16): included by encode_kernels.cu inside namespace flacb200.
//
// One CTA of 256 threads owns one unit (a stereo frame with its L/R/M/S candidates, or one channel of a frame);
// every thread owns 16 consecutive samples in registers for the whole analysis, so the fixed differences, the
// LPC FIR (INT32 x INT16 -> INT64 MACs over a register window), the partition sums and the exact Rice bit counts
// never re-read memory.  PCM is unpacked straight from the caller's packed bytes with 128-bit loads: the int32
// candidate planes of the generic path (k_planes) never reach HBM.

// block accumulate: every thread contributes v to *slot (zeroed and synchronised by the caller)
__device__ inline void block_add(unsigned long long* slot, unsigned long long v)
{
    v = warp_sum_u64(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(slot, v);
}

// ------------------------------------------------------------------------------------------------
// k_analyze: encode_subframe (src/encode.rs:2849-2980) without emitting bits -- ONE WARP PER CANDIDATE.
//
// The block is cut into 16-sample tiles; in round r lane l owns tile 32 r + l, so a round's PCM loads are one
// contiguous 128-bit-per-lane read.  A tile's history (the fixed differences and the LPC FIR look back <= 16 samples)
// is the left neighbour's tile, fetched with warp shuffles -- the samples never touch shared memory.  Everything that
// used to be a CTA-wide barrier is a warp shuffle; warps never wait for each other.
//   pass 1: |residual| sums of fixed orders 0..4 and of the LPC residual, per finest Rice partition
//           (24-bit limbs in warp-private shared memory, native 32-bit atomics) + the totals that pick the fixed order
//   then  : partition tree, Partition::new for every (order, partition), first-minimum partition order
//   pass 2: residuals of the chosen fixed order and of the LPC predictor again, exact Rice bit counts
// Wasted bits are assumed 0 while the OR mask is gathered in pass 1; a candidate that has some restarts once.
// ------------------------------------------------------------------------------------------------
constexpr int AW_WARPS = 4;
constexpr int AW_SETS = 6;   // fixed orders 0..4, LPC

struct AwSmem {   // per warp
    uint32_t limb_lo[AW_SETS][MAX_PARTS], limb_hi[AW_SETS][MAX_PARTS];
    unsigned long long tree[2][128];   // [set][ (1 << p) - 1 + j ]: sum |r| of partition j at order p
    uint32_t part_est[2][128];
    uint8_t part_code[2][128];
    unsigned long long u[5];   // per fixed order k: sum |r| of the samples in [k, kmax)
    RiceChoice choice[2];
};

// HB: the launch's max LPC order rounded up to 4/8/12/16 (one instantiation per launch keeps the instruction
// footprint small; predictors of lower order run with zero coefficients)
template <int HB, bool STEREO>
__device__ void aw_candidate(const EncCfg& cfg, const FrameDesc& d, const uint8_t* __restrict__ pcm, uint32_t slot, uint32_t full_bps,
                             const LpcRec& lp, AwSmem& sm, CandRec* __restrict__ rec)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = d.n;
    const uint32_t rounds = ((n + 15) / 16 + 31) / 32;   // in round r lane l owns the 16-sample tile r * 32 + l: coalesced PCM loads
    uint32_t p_max = (uint32_t)__ffs((int)n) - 1u;
    if (p_max > cfg.max_porder) p_max = cfg.max_porder;
    if (p_max > MAX_PORDER) p_max = MAX_PORDER;
    const uint32_t cf = n >> p_max;   // finest partition
    const bool cf16 = (cf & 15u) == 0;
    const uint32_t kmax = min(4u, n - 1);
    const bool have_lpc = lp.ok != 0;
    const uint32_t order = have_lpc ? lp.order : 0, shift = lp.shift;
    int32_t q[HB];
#pragma unroll
    for (int j = 0; j < HB; j++) q[j] = (have_lpc && (uint32_t)j < order) ? (int32_t)lp.q[j] : 0;

    uint32_t wasted = 0, mask = 0, ovf = 0, fo = 0, bps = full_bps;
    bool lpc_ok = have_lpc;
    unsigned long long s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    unsigned long long bits_f = 0, bits_l = 0;
    uint32_t bad_f = 0, bad_l = 0;
    uint32_t cpf = n, j0f = 0, cpl = n, j0l = 0;
    bool cpf16 = false, cpl16 = false;
    // stage 0: pass 1 assuming no wasted bits; stage 1: pass 1 again with the wasted bits shifted out (rare);
    // stage 2: pass 2 (exact sizes).  One loop body serves all stages so that the FIR code exists once.
    for (int stage = 0; stage < 3; stage++) {
        if (stage == 1 && wasted == 0) continue;
        if (stage < 2) {
            for (uint32_t t = lane; t < AW_SETS * MAX_PARTS; t += 32) {
                (&sm.limb_lo[0][0])[t] = 0;
                (&sm.limb_hi[0][0])[t] = 0;
            }
            if (lane < 5) sm.u[lane] = 0;
            mask = 0; ovf = 0;
            __syncwarp();
        } else {
            // ---- between the passes: fixed order, partition trees, Rice parameters ----
            if (mask == 0) {   // all samples zero -> CONSTANT (:2870, :2883)
                if (lane == 0) {
                    rec->type = 0; rec->order = 0; rec->wasted = 0; rec->bps = (uint8_t)full_bps;
                    rec->bits = 8 + full_bps;
                }
                return;
            }
            bps = full_bps - wasted;
            __syncwarp();
            {   // sums over the common tail = everything set k counted, minus its samples before kmax
                unsigned long long tk[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    unsigned long long v = 0;
                    for (uint32_t j = lane; j < MAX_PARTS; j += 32) v += (unsigned long long)sm.limb_lo[k][j] + ((unsigned long long)sm.limb_hi[k][j] << 24);
                    tk[k] = warp_sum_u64(v) - sm.u[k];
                }
                s0 = tk[0]; s1 = tk[1]; s2 = tk[2]; s3 = tk[3]; s4 = tk[4];
            }
            {   // first minimum among the orders that exist (:3065-3075)
                unsigned long long best = s0;
                if (kmax >= 1 && s1 < best) { best = s1; fo = 1; }
                if (kmax >= 2 && s2 < best) { best = s2; fo = 2; }
                if (kmax >= 3 && s3 < best) { best = s3; fo = 3; }
                if (kmax >= 4 && s4 < best) { best = s4; fo = 4; }
            }
            lpc_ok = have_lpc && !__any_sync(0xffffffffu, ovf >> 31);   // ResidualOverflow
            __syncwarp();
            const uint32_t nleaf = 1u << p_max;
            for (uint32_t k = 0; k < 2; k++) {
                const uint32_t set = k == 0 ? fo : 5;
                for (uint32_t j = lane; j < nleaf; j += 32)
                    sm.tree[k][nleaf - 1 + j] = (unsigned long long)sm.limb_lo[set][j] + ((unsigned long long)sm.limb_hi[set][j] << 24);
            }
            __syncwarp();
            for (int p = (int)p_max - 1; p >= 0; p--) {
                const uint32_t base = (1u << p) - 1, child = (2u << p) - 1;
                for (uint32_t j = lane; j < (1u << p); j += 32) {
                    sm.tree[0][base + j] = sm.tree[0][child + 2 * j] + sm.tree[0][child + 2 * j + 1];
                    sm.tree[1][base + j] = sm.tree[1][child + 2 * j] + sm.tree[1][child + 2 * j + 1];
                }
                __syncwarp();
            }
            aw_choose_partitions(cfg, n, fo, p_max, sm.tree[0], sm.part_est[0], sm.part_code[0], sm.choice[0]);
            if (lpc_ok) aw_choose_partitions(cfg, n, order, p_max, sm.tree[1], sm.part_est[1], sm.part_code[1], sm.choice[1]);
            cpf = n >> sm.choice[0].porder_g;
            j0f = (1u << sm.choice[0].porder_g) - sm.choice[0].nparts;
            cpf16 = (cpf & 15u) == 0;
            if (lpc_ok) {
                cpl = n >> sm.choice[1].porder_g;
                j0l = (1u << sm.choice[1].porder_g) - sm.choice[1].nparts;
                cpl16 = (cpl & 15u) == 0;
            }
        }
        int32_t carry[16];   // lane 31's tile of the previous round: the history of lane 0
#pragma unroll
        for (int e = 0; e < 16; e++) carry[e] = 0;
        for (uint32_t rd = 0; rd < rounds; rd++) {
            const uint32_t i0 = (rd * 32 + lane) * 16;
            const bool live = i0 < n;
            int32_t x[16], h[16];   // h: the 16 samples before the tile (after the wasted-bit shift) = the left neighbour's tile
#pragma unroll
            for (int e = 0; e < 16; e++) x[e] = 0;
            if (live) aw_load_tile<STEREO>(cfg, d, pcm, slot, i0, x);
#pragma unroll
            for (int e = 0; e < 16; e++) { mask |= (uint32_t)x[e]; x[e] >>= wasted; }   // :2878-2898 (mask only meaningful in stage 0)
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int32_t up = __shfl_up_sync(0xffffffffu, x[e], 1);
                h[e] = lane == 0 ? carry[e] : up;
                carry[e] = __shfl_sync(0xffffffffu, x[e], 31);
            }
            if (!live) continue;
            const bool tail = i0 + 16 > n;   // tile cut by the block end: rare, sample-by-sample path
            // ---- LPC residuals (:3174-3203) ----
            int32_t rl[16];
            if (lpc_ok) {
                uint32_t om = 0;
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    long long sum = 0;
#pragma unroll
                    for (int j = 0; j < HB; j++) sum = mad_wide_s32(e - 1 - j >= 0 ? x[e - 1 - j >= 0 ? e - 1 - j : 0] : h[16 + e - 1 - j >= 0 ? 16 + e - 1 - j : 0], q[j], sum);
                    const int32_t pred = (int32_t)(uint32_t)(unsigned long long)(sum >> shift);   // `as i32`
                    const int32_t rr = (int32_t)((uint32_t)x[e] - (uint32_t)pred);
                    rl[e] = rr;
                    om |= (((uint32_t)((x[e] ^ pred) & (x[e] ^ rr))) >> 31) << e;   // checked_sub: sign bit set when it overflowed
                }
                if (stage < 2 && om) {   // only samples in [order, n) count
                    const uint32_t first = order > i0 ? min(order - i0, 16u) : 0u, last = min(16u, n - i0);
                    const uint32_t valid = (last >= 16 ? 0xFFFFu : (1u << last) - 1u) & ~((1u << first) - 1u);
                    if (om & valid) ovf = 0x80000000u;
                }
            }
            // ---- fixed differences (:3039-3060); <= 28-bit samples cannot overflow i32 up to order 4 ----
            int32_t p1 = h[15] - h[14], p2 = p1 - (h[14] - h[13]), p3 = p2 - ((h[14] - h[13]) - (h[13] - h[12]));
            if (stage < 2) {
                const uint32_t chunk = i0 / cf;
                if (!tail && (cf16 || (i0 + 15) / cf == chunk)) {
                    unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, al = 0;
                    int32_t prev = h[15];
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const int32_t e1 = x[e] - prev, e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
                        prev = x[e]; p1 = e1; p2 = e2; p3 = e3;
                        a0 = acc_u32(a0, uabs32(x[e])); a1 = acc_u32(a1, uabs32(e1)); a2 = acc_u32(a2, uabs32(e2));
                        a3 = acc_u32(a3, uabs32(e3)); a4 = acc_u32(a4, uabs32(e4));
                        if (lpc_ok) al = acc_u32(al, uabs32(rl[e]));
                    }
                    if (i0 == 0) {
                        // the block's first tile (n >= 16 here, so kmax == 4): with zero history the differences of the first
                        // samples are these closed forms.  Set k does not count samples before k; the order comparison
                        // (:3062-3073) does not count samples before kmax either -- remembered in sm.u[] for later.
                        const int32_t y0 = x[0], y1 = x[1], y2 = x[2], y3 = x[3];
                        const uint32_t f11 = uabs32(y1 - y0), f12 = uabs32(y2 - y1), f13 = uabs32(y3 - y2);
                        const uint32_t f21 = uabs32(y1 - 2 * y0), f22 = uabs32(y2 - 2 * y1 + y0), f23 = uabs32(y3 - 2 * y2 + y1);
                        const uint32_t f31 = uabs32(y1 - 3 * y0), f32 = uabs32(y2 - 3 * y1 + 3 * y0), f33 = uabs32(y3 - 3 * y2 + 3 * y1 - y0);
                        const uint32_t f41 = uabs32(y1 - 4 * y0), f42 = uabs32(y2 - 4 * y1 + 6 * y0), f43 = uabs32(y3 - 4 * y2 + 6 * y1 - 4 * y0);
                        const unsigned long long f0 = uabs32(y0);
                        a1 -= f0;
                        a2 -= f0 + f21;
                        a3 -= f0 + f31 + f32;
                        a4 -= f0 + f41 + f42 + f43;
                        sm.u[0] = f0 + uabs32(y1) + uabs32(y2) + uabs32(y3);
                        sm.u[1] = (unsigned long long)f11 + f12 + f13;
                        sm.u[2] = (unsigned long long)f22 + f23;
                        sm.u[3] = f33;
                        sm.u[4] = 0;
                        if (lpc_ok) {
#pragma unroll
                            for (int e = 0; e < 16; e++)
                                if ((uint32_t)e < order) al -= uabs32(rl[e]);
                        }
                    }
                    aw_add_limbs(sm.limb_lo[0], sm.limb_hi[0], chunk, a0);
                    aw_add_limbs(sm.limb_lo[1], sm.limb_hi[1], chunk, a1);
                    aw_add_limbs(sm.limb_lo[2], sm.limb_hi[2], chunk, a2);
                    aw_add_limbs(sm.limb_lo[3], sm.limb_hi[3], chunk, a3);
                    aw_add_limbs(sm.limb_lo[4], sm.limb_hi[4], chunk, a4);
                    if (lpc_ok) aw_add_limbs(sm.limb_lo[5], sm.limb_hi[5], chunk, al);
                } else {
                    // copies: taking the address of the register tiles themselves would push them into local memory for good
                    int32_t tx[16], th[16], tl[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) { tx[e] = x[e]; th[e] = h[e]; tl[e] = lpc_ok ? rl[e] : 0; }
                    AwSmemLimbs limbs = {&sm.limb_lo[0][0], &sm.limb_hi[0][0]};
                    aw_pass1_tile_slow(tx, th, tl, lpc_ok, order, i0, n, kmax, cf, limbs, sm.u);
                }
            } else {
                // ---- pass 2: exact size of both residual blocks (what Partition::to_writer will emit, :3834-3863) ----
                int32_t rf[16];
                {
                    int32_t prev = h[15];
                    switch (fo) {   // uniform across the warp
                    case 0:
#pragma unroll
                        for (int e = 0; e < 16; e++) rf[e] = x[e];
                        break;
                    case 1:
#pragma unroll
                        for (int e = 0; e < 16; e++) { rf[e] = x[e] - prev; prev = x[e]; }
                        break;
                    default:
#pragma unroll
                        for (int e = 0; e < 16; e++) {
                            const int32_t e1 = x[e] - prev, e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
                            prev = x[e]; p1 = e1; p2 = e2; p3 = e3;
                            rf[e] = fo == 2 ? e2 : fo == 3 ? e3 : e4;
                        }
                        break;
                    }
                }
                const RiceChoice& chf = sm.choice[0];
                const uint32_t pf = i0 / cpf;
                const uint32_t codef = chf.rice[pf - j0f >= chf.nparts ? 0 : pf - j0f];
                if (!tail && i0 != 0 && (cpf16 || (i0 + 15) / cpf == pf) && codef < 0x40) {
#pragma unroll
                    for (int e = 0; e < 16; e++) bits_f = acc_u32(bits_f, zigzag32(rf[e]) >> codef);
                    bits_f += 16u * (1u + codef);
                } else if (!tail && i0 == 0 && cpf >= 16 && codef < 0x40) {   // the block's first tile: `fo` warm-up samples, then partition 0
#pragma unroll
                    for (int e = 0; e < 16; e++)
                        if ((uint32_t)e >= fo) bits_f = acc_u32(bits_f, zigzag32(rf[e]) >> codef);
                    bits_f += (16u - fo) * (1u + codef);
                } else {
                    int32_t tmp[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) tmp[e] = rf[e];
                    unsigned long long tb = 0;
                    uint32_t tbad = 0;
                    aw_tile_bits_slow(tmp, i0, fo, n, cpf, j0f, chf.rice, &tb, &tbad);
                    bits_f += tb;
                    bad_f |= tbad;
                }
                if (lpc_ok) {
                    const RiceChoice& chl = sm.choice[1];
                    const uint32_t pl = i0 / cpl;
                    const uint32_t codel = chl.rice[pl - j0l >= chl.nparts ? 0 : pl - j0l];
                    if (!tail && i0 != 0 && (cpl16 || (i0 + 15) / cpl == pl) && codel < 0x40) {
#pragma unroll
                        for (int e = 0; e < 16; e++) bits_l = acc_u32(bits_l, zigzag32(rl[e]) >> codel);
                        bits_l += 16u * (1u + codel);
                    } else if (!tail && i0 == 0 && cpl >= 16 && order <= 16 && codel < 0x40) {
#pragma unroll
                        for (int e = 0; e < 16; e++)
                            if ((uint32_t)e >= order) bits_l = acc_u32(bits_l, zigzag32(rl[e]) >> codel);
                        bits_l += (16u - order) * (1u + codel);
                    } else {
                        int32_t tmp[16];
#pragma unroll
                        for (int e = 0; e < 16; e++) tmp[e] = rl[e];
                        unsigned long long tb = 0;
                        uint32_t tbad = 0;
                        aw_tile_bits_slow(tmp, i0, order, n, cpl, j0l, chl.rice, &tb, &tbad);
                        bits_l += tb;
                        bad_l |= tbad;
                    }
                }
            }
        }
        if (stage < 2) {
            mask = __reduce_or_sync(0xffffffffu, mask);
            if (stage == 0) wasted = (mask == 0 || (mask & 1u)) ? 0u : (uint32_t)__ffs((int)mask) - 1u;
        }
    }
    const RiceChoice& cf_ = sm.choice[0];
    const RiceChoice& cl_ = sm.choice[1];
    // partition headers: 4/5-bit parameter (+ 5-bit escape width)
    for (uint32_t j = lane; j < cf_.nparts; j += 32) bits_f += (cf_.rice[j] < 0x40) ? (cf_.method ? 5u : 4u) : (cf_.method ? 10u : 9u);
    if (lpc_ok)
        for (uint32_t j = lane; j < cl_.nparts; j += 32) bits_l += (cl_.rice[j] < 0x40) ? (cl_.method ? 5u : 4u) : (cl_.method ? 10u : 9u);
    bits_f = warp_sum_u64(bits_f);
    bits_l = warp_sum_u64(bits_l);
    const bool fixed_ok = !__any_sync(0xffffffffu, bad_f);
    if (__any_sync(0xffffffffu, bad_l)) lpc_ok = false;
    const uint32_t hdr_bits = 8 + wasted;   // pad + type + wasted flag (+ unary(wasted - 1)) (src/stream.rs:1397)
    const uint32_t fixed_bits = hdr_bits + fo * bps + (uint32_t)bits_f + 6;
    const uint32_t lpc_bits = hdr_bits + order * bps + 4 + 5 + order * lp.precision + (uint32_t)bits_l + 6;
    // ---- choose (:2929-2979): fixed wins ties; VERBATIM unless strictly smaller ----
    const uint32_t verbatim_len = n * bps;
    int pick = -1;   // 0 fixed, 1 lpc
    if (fixed_ok && lpc_ok) pick = lpc_bits < fixed_bits ? 1 : 0;
    else if (fixed_ok) pick = 0;
    else if (lpc_ok) pick = 1;
    const uint32_t best_bits = pick == 1 ? lpc_bits : fixed_bits;
    if (pick >= 0 && !(best_bits < verbatim_len)) pick = -1;
    const RiceChoice& ch = pick == 1 ? sm.choice[1] : sm.choice[0];
    if (lane == 0) {
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
        for (uint32_t j = lane; j < MAX_PARTS; j += 32) rec->rice[j] = ch.rice[j];
        if (lane < MAX_LPC) rec->q[lane] = lp.q[lane];
    }
    __syncwarp();
}

// one warp per candidate: grid = ceil(ncand / AW_WARPS)
template <bool STEREO>
__global__ void __launch_bounds__(32 * AW_WARPS, 4) k_analyze(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
                                                          const LpcRec* __restrict__ lpcs, CandRec* __restrict__ out,
                                                          unsigned long long* __restrict__ abssum, uint32_t ncand)
{
    __shared__ AwSmem sm_all[AW_WARPS];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cand = blockIdx.x * AW_WARPS + wid;
    if (cand >= ncand) return;
    AwSmem& sm = sm_all[wid];
    const uint32_t f = cand / cfg.nslots, slot = cand % cfg.nslots;
    const FrameDesc d = descs[f];
    CandRec* rec = out + cand;
    if (STEREO) {
        if (cfg.mode == MODE_FAST_MID_SIDE || cfg.mode == MODE_FAST_SIDE) {   // correlate_channels abs sums (:2475-2503)
            unsigned long long sl = 0, sr = 0, smid = 0, sside = 0;
            for (uint32_t i0 = lane * 16; i0 < d.n; i0 += 32 * 16) {
                int32_t a[16], b[16];
                load_thread_samples<2>(cfg, d, pcm, i0, 0, a, b);
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    sl += uabs32(a[e]); sr += uabs32(b[e]); smid += uabs32((a[e] + b[e]) >> 1); sside += uabs32(a[e] - b[e]);
                }
            }
            unsigned long long sums[4] = {warp_sum_u64(sl), warp_sum_u64(sr), warp_sum_u64(smid), warp_sum_u64(sside)};
            if (slot == 0 && lane < 4) abssum[(size_t)f * 4 + lane] = sums[lane];
            if (!slot_active(cfg, sums, slot)) {
                if (lane == 0) { rec->type = 0xFF; rec->bits = 0; }
                return;
            }
        } else if (cfg.mode == MODE_EXH_SIDE && slot == 2) {
            if (lane == 0) { rec->type = 0xFF; rec->bits = 0; }
            return;
        }
    }
    const uint32_t full_bps = STEREO ? cand_bps(cfg, slot) : cfg.bps;
    const LpcRec lp = lpcs[cand];
    const uint32_t hb = cfg.max_lpc_order ? (cfg.max_lpc_order + 3u) >> 2 : 1u;   // uniform for the whole launch
    switch (hb) {
    case 1: aw_candidate<4, STEREO>(cfg, d, pcm, slot, full_bps, lp, sm, rec); break;
    case 2: aw_candidate<8, STEREO>(cfg, d, pcm, slot, full_bps, lp, sm, rec); break;
    case 3: aw_candidate<12, STEREO>(cfg, d, pcm, slot, full_bps, lp, sm, rec); break;
    default: aw_candidate<16, STEREO>(cfg, d, pcm, slot, full_bps, lp, sm, rec); break;
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
    const uint32_t ncand = cfg.nframes * cfg.nslots;
    const uint32_t grid = (ncand + AW_WARPS - 1) / AW_WARPS;
    if (cfg.mode != MODE_INDEPENDENT) count_launch(), k_analyze<true><<<grid, 32 * AW_WARPS, 0, st>>>(cfg, descs, pcm, lpcs, cands, abssum, ncand);
    else count_launch(), k_analyze<false><<<grid, 32 * AW_WARPS, 0, st>>>(cfg, descs, pcm, lpcs, cands, abssum, ncand);
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
    const bool rawmode = stereo4 && cfg.pcm_kind <= 1 && (reinterpret_cast<uintptr_t>(pcm) & 1) == 0;   // packed bytes, 16-bit loads
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
        uint32_t raw[L2_MAXC / 4][4];   // rawmode: the frame's 2 * B PCM bytes as 16-bit halves, decoded only in store_tile
        auto fetch_tile = [&](uint32_t tile) {   // global loads only: nothing here waits for them
            const uint32_t idx = tile * 32 + lane;
            if (rawmode) {
#pragma unroll
                for (int fi = 0; fi < L2_MAXC / 4; fi++) {
                    const uint32_t c = fi * 4;
                    double wv = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; k++) raw[fi][k] = 0;
                    if (c < ncs) {
                        const FrameDesc d = descs[(c0 + c) >> 2];
                        if (idx < d.n && d.n > M) {
                            const uint16_t* p16 = reinterpret_cast<const uint16_t*>(pcm + (d.pcm_off + idx) * (unsigned long long)(2 * cfg.bytes_per_sample));
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                if ((uint32_t)k < cfg.bytes_per_sample) raw[fi][k] = p16[k];
                            wv = winpool[d.win_off + idx];
                        }
                    }
                    pw[c] = pw[c + 1] = pw[c + 2] = pw[c + 3] = wv;
                }
            } else if (stereo4) {
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
            if (rawmode) {   // first use of the prefetched bytes (Frame::fill_from_buf, src/audio.rs:149-187)
                const uint32_t B = cfg.bytes_per_sample, sh = 32 - 8 * B;
#pragma unroll
                for (int fi = 0; fi < L2_MAXC / 4; fi++) {
                    const unsigned long long wide = (unsigned long long)(raw[fi][0] | (raw[fi][1] << 16)) |
                                                    ((unsigned long long)(raw[fi][2] | (raw[fi][3] << 16)) << 32);
                    uint32_t lw = (uint32_t)wide, rw = (uint32_t)(wide >> (8 * B));   // low B bytes: the sample in memory order
                    int32_t l, r;
                    if (cfg.pcm_kind == 1) {
                        l = (int32_t)__byte_perm(lw, 0, 0x0123) >> sh;
                        r = (int32_t)__byte_perm(rw, 0, 0x0123) >> sh;
                    } else {
                        l = (int32_t)(lw << sh) >> sh;
                        r = (int32_t)(rw << sh) >> sh;
                    }
                    px[fi * 4] = l; px[fi * 4 + 1] = r; px[fi * 4 + 2] = (l + r) >> 1; px[fi * 4 + 3] = l - r;
                }
#pragma unroll
                for (int c = (L2_MAXC / 4) * 4; c < L2_MAXC; c++) { px[c] = 0; pw[c] = 0.0; }
            }
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
            const double bpr = __ddiv_rn(glibc_log(__dmul_rn(my.err[o - 1], error_scale)), divisor);
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
                const int32_t lg = f64_as_i32_sat(floor(glibc_log2(l)));
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
    {   // per device, so on every launch (a process may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(k_lpc2, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        if (e != cudaSuccess) return e;
    }
    count_launch(), k_lpc2<<<(nwarps + L2_WARPS - 1) / L2_WARPS, 32 * L2_WARPS, smem, st>>>(cfg, descs, pcm, winpool, lpcs, cpw, ncand);
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
            switch ((cfg.max_lpc_order + 3) >> 2) {   // uniform for the launch: one FIR body in the instruction cache
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
    {
        cudaError_t e = cudaFuncSetAttribute(k_pack2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pack2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
    }
    if (cfg.mode != MODE_INDEPENDENT)
        count_launch(), k_pack2<true><<<cfg.nframes * nsub_max, AN_THREADS, smem, st>>>(cfg, nsub_max, cap_words, descs, pcm, cands, frecs, out);
    else
        count_launch(), k_pack2<false><<<cfg.nframes * nsub_max, AN_THREADS, smem, st>>>(cfg, nsub_max, cap_words, descs, pcm, cands, frecs, out);
    count_launch(), k_crc16w<<<(cfg.nframes + 7) / 8, 256, 0, st>>>(frecs, cfg.nframes, out);
    return cudaGetLastError();
}
