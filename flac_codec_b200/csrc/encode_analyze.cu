// encode_analyze.cu -- k_analyze3: encode_subframe (src/encode.rs:2849-2980) without emitting bits, ONE CTA PER FRAME.
//
// The candidates of a frame (stereo: L, R, M = (L + R) >> 1, S = L - R; otherwise two channels per CTA) share one
// unpacking of the packed PCM: the CTA writes the two source channels once into shared memory as int32 planes in a
// tile-transposed layout (tile = 16 samples; chunk c of tile t sits at ((t / 32) * 4 + c) * 128 + (t % 32) * 4 words,
// so that the 128-bit loads of 32 lanes that own 32 consecutive tiles are conflict free) and every candidate reads
// its tiles AND their history from there -- no per-candidate unpacking, no shuffles.
//
// Two warps work on one candidate (alternating rounds of 32 tiles) and meet on a named barrier:
//   pass 1: |residual| sums of the fixed orders 0..4 and of the LPC residual per finest Rice partition (24-bit limbs,
//           shared-memory atomics), the OR mask (wasted bits), checked_sub overflow; the LPC residuals are parked in
//           shared memory as int16 (flag if one does not fit)
//   then  : fixed order (:3062-3075), partition trees, Partition::new for every (order, partition), first-minimum
//           partition order (:3865-3942) -- by the candidate's first warp
//   pass 2: exact Rice bit counts (what Partition::to_writer will emit, :3834-3863): the fixed residual is rebuilt from
//           the planes (<= 4 subtractions), the LPC residual is read back from the int16 copy -- the FIR runs once per
//           candidate instead of twice (it is recomputed only for the rare candidate whose residuals overflow int16)
// Wasted bits are assumed 0 in pass 1; a candidate that has some repeats pass 1 with the shift applied.
#include <cstdlib>

#include "common.cuh"
#include "frame_decide.cuh"
#include "pack_bits.cuh"
#include "tiles.cuh"
#include "rice.cuh"

namespace flacb200 {

bool analyze_fast_ok(const EncCfg& cfg);   // encode_kernels.cu

#ifndef FLACB200_A3_WPC
#define FLACB200_A3_WPC 2
#endif
constexpr int A3_WPC = FLACB200_A3_WPC;   // warps per candidate
// LPC residual FIR of k_analyze3: 0 = IMAD.WIDE chains, 1 = DFMA chains, outputs in two halves of 8, 2 = DFMA, 16 outputs at once
#ifndef FLACB200_A3_FIR
#define FLACB200_A3_FIR 1
#endif
constexpr int A3_FIR_OUT = FLACB200_A3_FIR == 2 ? 16 : 8;
#ifndef FLACB200_A3_PF
#define FLACB200_A3_PF 1
#endif
#ifndef FLACB200_A3_PF_DIST
#define FLACB200_A3_PF_DIST 296
#endif
constexpr uint32_t A3_PF_DIST = FLACB200_A3_PF_DIST;   // frames ahead: the CTAs resident at once (2 x 148 SMs)
// int32 -> double without the conversion unit: 2^52 + 2^31 + v is exact bit-pasting, the subtraction is one DADD
#ifndef FLACB200_A3_I2F
#define FLACB200_A3_I2F 0
#endif
__device__ inline double a3_i2d(int32_t v)
{
#if FLACB200_A3_I2F
    return (double)v;
#else
    return __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)), 4503601774854144.0);
#endif
}
constexpr int A3_SETS = 6;       // fixed orders 0..4, LPC
constexpr int A3_PLANE = 4096;   // samples per plane (largest block of the register-tiled kernels)

struct A3Cand {   // per candidate, shared by its warps
    uint32_t limb_lo[A3_SETS][MAX_PARTS], limb_hi[A3_SETS][MAX_PARTS];
    unsigned long long tree[2][128];   // [set][(1 << p) - 1 + j]: sum |r| of partition j at order p
    uint32_t part_est[2][128];
    uint8_t part_code[2][128];
    unsigned long long u[5];           // per fixed order k: sum |r| of the samples in [k, kmax)
    RiceChoice choice[2];
    unsigned long long bits_f, bits_l;
    uint32_t mask, ovf, bad16, bad_f, bad_l;
    uint32_t fo, lpc_ok;               // decisions of the first warp, read by the second
    uint32_t lb_f;                     // lower bound of the fixed predictor's residual code bits (aw_choose_partitions_flat)
    uint32_t round_bits[2][8];         // pass 2: code bits per round of 32 tiles, [fixed | LPC] (k_frame4 places its warps with them)
};

__device__ inline void a3_pair_sync(uint32_t cand)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1u + cand), "r"(32u * A3_WPC) : "memory");
}

__device__ inline uint32_t a3_tile_base(uint32_t t) { return ((t >> 5) * 128u + (t & 31u)) * 4u; }   // word index of chunk 0

// chunks [c0, 4) of tile t of the candidate into v[4 c0 .. 16); slot selects the stereo combination (uniform across the
// warp): one branch per tile, the chunk loads inside are straight-line 128-bit shared loads
template <bool STEREO>
__device__ inline void a3_tile(const int32_t* __restrict__ planes, uint32_t slot, uint32_t t, int c0, int32_t* v)
{
    const uint32_t w = a3_tile_base(t);
    if (!STEREO || slot < 2) {
        const int32_t* p = planes + slot * A3_PLANE + w;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c < c0) continue;
            const int4 a = *reinterpret_cast<const int4*>(p + c * 128);
            v[4 * c] = a.x; v[4 * c + 1] = a.y; v[4 * c + 2] = a.z; v[4 * c + 3] = a.w;
        }
    } else if (slot == 2) {   // mid (:2721)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c < c0) continue;
            const int4 a = *reinterpret_cast<const int4*>(planes + w + c * 128);
            const int4 b = *reinterpret_cast<const int4*>(planes + A3_PLANE + w + c * 128);
            v[4 * c] = (a.x + b.x) >> 1; v[4 * c + 1] = (a.y + b.y) >> 1; v[4 * c + 2] = (a.z + b.z) >> 1; v[4 * c + 3] = (a.w + b.w) >> 1;
        }
    } else {                  // side (:2734)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c < c0) continue;
            const int4 a = *reinterpret_cast<const int4*>(planes + w + c * 128);
            const int4 b = *reinterpret_cast<const int4*>(planes + A3_PLANE + w + c * 128);
            v[4 * c] = a.x - b.x; v[4 * c + 1] = a.y - b.y; v[4 * c + 2] = a.z - b.z; v[4 * c + 3] = a.w - b.w;
        }
    }
}

// HB: the launch's max LPC order rounded up to 4/8/12/16
template <int HB, bool STEREO>
__device__ void a3_candidate(const EncCfg& cfg, const FrameDesc& d, const int32_t* __restrict__ planes, uint4* __restrict__ res16, uint32_t slot,
                             uint32_t pslot, uint32_t cand, uint32_t wsub, uint32_t full_bps, const LpcRec& lp, A3Cand& sm, CandRec* __restrict__ rec,
                             uint4* __restrict__ gres16 = nullptr)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = d.n;
    const uint32_t rounds = ((n + 15) / 16 + 31) / 32;
    uint32_t p_max = (uint32_t)__ffs((int)n) - 1u;
    if (p_max > cfg.max_porder) p_max = cfg.max_porder;
    if (p_max > MAX_PORDER) p_max = MAX_PORDER;
    const uint32_t cf = n >> p_max;   // finest partition
    const bool cf16 = (cf & 15u) == 0;
    const UDiv dcf = udiv_make(cf);
    UDiv dcpf = dcf, dcpl = dcf;
    const uint32_t kmax = min(4u, n - 1);
    const bool have_lpc = lp.ok != 0;
    const uint32_t order = have_lpc ? lp.order : 0, shift = lp.shift;
#if FLACB200_A3_FIR == 0
    int32_t q[HB];
#pragma unroll
    for (int j = 0; j < HB; j++) q[j] = (have_lpc && (uint32_t)j < order) ? (int32_t)lp.q[j] : 0;
#else
    double qd[HB];   // the coefficients as doubles: the FIR runs on the FP64 pipe (see a3_fir_f64)
#pragma unroll
    for (int j = 0; j < HB; j++) qd[j] = (have_lpc && (uint32_t)j < order) ? (double)lp.q[j] : 0.0;
#endif

    uint32_t wasted = 0, fo = 0, bps = full_bps;
    bool lpc_ok = have_lpc, use16 = false;
    unsigned long long bits_f = 0, bits_l = 0;
    uint32_t bad_f = 0, bad_l = 0;
    uint32_t cpf = n, j0f = 0, cpl = n, j0l = 0;
    bool cpf16 = false, cpl16 = false;
    // stage 0: pass 1 assuming no wasted bits; stage 1: pass 1 again with the wasted bits shifted out (rare);
    // stages 2 and 3: pass 2 (exact sizes).  One loop body serves all stages so that the FIR code exists once.
    // Pass 2 runs as two sweeps: first the LPC residuals (the int16 copy: no plane loads), then the fixed residuals -- unless
    // the exact LPC size is already below a lower bound of the fixed size (sm.lb_f), in which case LPC wins whatever the exact
    // fixed size is (:2929-2979) and the second sweep is not needed.  A candidate without a 16-bit LPC copy does both in stage 3.
    uint32_t hf = 0, hl = 0;   // partition headers: 4/5-bit parameter (+ 5-bit escape width)
    bool skip_f = false, swept_l = false;
    for (int stage = 0; stage < 4; stage++) {
        if (stage == 1 && wasted == 0) continue;
        uint32_t mask = 0, ovf = 0, b16 = 0;
        if (stage == 3) {
            if (swept_l) {   // the LPC sweep's total, then the decision
                bits_l = warp_sum_u64(bits_l);
                bad_l = __any_sync(0xffffffffu, bad_l) ? 1u : 0u;
                if (lane == 0) {
                    if (bits_l) atomicAdd(&sm.bits_l, bits_l);
                    if (bad_l) atomicOr(&sm.bad_l, 1u);
                }
                bits_l = 0; bad_l = 0;
                a3_pair_sync(cand);
                const uint32_t hdr_bits = 8 + wasted;
                const unsigned long long lpc_bits = (unsigned long long)hdr_bits + order * bps + 4 + 5 + order * lp.precision + (sm.bits_l + hl) + 6;
                const unsigned long long fixed_lb = (unsigned long long)hdr_bits + fo * bps + ((unsigned long long)sm.lb_f + hf) + 6;
                skip_f = sm.bad_l == 0 && lpc_bits < fixed_lb && lpc_bits < 0xFFFFFFFFull;
                if (skip_f) break;   // (use16 holds here: nothing else is left to count)
            }
        } else if (stage < 2) {
            if (wsub == 0) {
                for (uint32_t t = lane; t < A3_SETS * MAX_PARTS; t += 32) {
                    (&sm.limb_lo[0][0])[t] = 0;
                    (&sm.limb_hi[0][0])[t] = 0;
                }
                if (lane < 5) sm.u[lane] = 0;
                if (lane == 0) {
                    if (stage == 0) sm.mask = 0;   // the OR mask is gathered in stage 0 only
                    sm.ovf = 0; sm.bad16 = 0; sm.bad_f = 0; sm.bad_l = 0; sm.bits_f = 0; sm.bits_l = 0;
                }
            }
            a3_pair_sync(cand);
        } else {
            // ---- between the passes (stage 2): fixed order, partition trees, Rice parameters (first warp of the candidate) ----
            if (sm.mask == 0) {   // all samples zero -> CONSTANT (:2870, :2883)
                if (wsub == 0 && lane == 0) {
                    rec->type = 0; rec->order = 0; rec->wasted = 0; rec->bps = (uint8_t)full_bps;
                    rec->bits = 8 + full_bps;
                }
                return;
            }
            bps = full_bps - wasted;
            if (wsub == 0) {   // first warp: the fixed predictor
                unsigned long long s0, s1, s2, s3, s4;
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
                    fo = 0;
                    if (kmax >= 1 && s1 < best) { best = s1; fo = 1; }
                    if (kmax >= 2 && s2 < best) { best = s2; fo = 2; }
                    if (kmax >= 3 && s3 < best) { best = s3; fo = 3; }
                    if (kmax >= 4 && s4 < best) { best = s4; fo = 4; }
                }
                aw_choose_partitions_flat(cfg, n, fo, p_max, sm.limb_lo[fo], sm.limb_hi[fo], sm.tree[0], sm.part_code[0], sm.choice[0], &sm.lb_f);
                if (lane == 0) sm.fo = fo;
            } else if (wsub == 1) {   // second warp: the LPC predictor
                lpc_ok = have_lpc && sm.ovf == 0;   // ResidualOverflow
                if (lpc_ok) {
                    aw_choose_partitions_flat(cfg, n, order, p_max, sm.limb_lo[5], sm.limb_hi[5], sm.tree[1], sm.part_code[1], sm.choice[1]);
                }
                if (lane == 0) sm.lpc_ok = lpc_ok ? 1u : 0u;
            }
            a3_pair_sync(cand);
            fo = sm.fo;
            lpc_ok = sm.lpc_ok != 0;
            use16 = sm.bad16 == 0;
            cpf = n >> sm.choice[0].porder_g;
            j0f = (1u << sm.choice[0].porder_g) - sm.choice[0].nparts;
            cpf16 = (cpf & 15u) == 0;
            dcpf = udiv_make(cpf);
            if (lpc_ok) {
                cpl = n >> sm.choice[1].porder_g;
                j0l = (1u << sm.choice[1].porder_g) - sm.choice[1].nparts;
                cpl16 = (cpl & 15u) == 0;
                dcpl = udiv_make(cpl);
            }
            {
                const RiceChoice& cf_ = sm.choice[0];
                const RiceChoice& cl_ = sm.choice[1];
                for (uint32_t j = lane; j < cf_.nparts; j += 32) hf += (cf_.rice[j] < 0x40) ? (cf_.method ? 5u : 4u) : (cf_.method ? 10u : 9u);
                if (lpc_ok)
                    for (uint32_t j = lane; j < cl_.nparts; j += 32) hl += (cl_.rice[j] < 0x40) ? (cl_.method ? 5u : 4u) : (cl_.method ? 10u : 9u);
                hf = __reduce_add_sync(0xffffffffu, hf);
                hl = __reduce_add_sync(0xffffffffu, hl);
            }
            swept_l = lpc_ok && use16;
            if (!swept_l) continue;   // everything is counted in stage 3
        }
        for (uint32_t rd = wsub; rd < rounds; rd += A3_WPC) {
            const uint32_t t = rd * 32 + lane, i0 = t * 16;
            const unsigned long long rb_f0 = bits_f, rb_l0 = bits_l;
            if (i0 < n) {
            const bool tail = i0 + 16 > n;   // tile cut by the block end: rare, sample-by-sample path
            const bool fir = lpc_ok && (stage < 2 || (stage == 3 && !use16));
            int32_t x[16], h[16];
#pragma unroll
            for (int e = 0; e < 16; e++) { x[e] = 0; h[e] = 0; }
            if (stage != 2) {   // (the LPC sweep reads the int16 copy only)
                a3_tile<STEREO>(planes, pslot, t, 0, x);
                if (t > 0) {   // the fixed differences look back 4 samples, the FIR HB
                    if (fir) a3_tile<STEREO>(planes, pslot, t - 1, 4 - HB / 4, h);
                    else a3_tile<STEREO>(planes, pslot, t - 1, 3, h);
                }
            }
            if (stage == 0) {
#pragma unroll
                for (int e = 0; e < 16; e++) mask |= (uint32_t)x[e];
            } else if (wasted) {   // :2878-2898
#pragma unroll
                for (int e = 0; e < 16; e++) { x[e] >>= wasted; h[e] >>= wasted; }
            }
            // ---- LPC residuals (:3174-3203) ----
            int32_t rl[16];
            if (fir) {
                uint32_t oacc = 0, sacc = 0;   // OR of the checked_sub overflow signs / of the bits that do not fit int16
#if FLACB200_A3_FIR == 0
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    long long sum = 0;
#pragma unroll
                    for (int j = 0; j < HB; j++) sum = mad_wide_s32(e - 1 - j >= 0 ? x[e - 1 - j >= 0 ? e - 1 - j : 0] : h[16 + e - 1 - j >= 0 ? 16 + e - 1 - j : 0], q[j], sum);
                    const int32_t pred = (int32_t)(uint32_t)(unsigned long long)(sum >> shift);   // `as i32`
                    const int32_t rr = (int32_t)((uint32_t)x[e] - (uint32_t)pred);
                    rl[e] = rr;
                    oacc |= (uint32_t)((x[e] ^ pred) & (x[e] ^ rr));   // checked_sub: sign bit set when it overflowed
                    sacc |= (uint32_t)rr + 0x8000u;
                }
#else
                // The i64 dot product (:3187) as DFMAs: |q| < 2^15, |x| < 2^25, <= 16 taps -- every partial sum is an integer
                // below 2^45 and exact in a double.  The accumulators start at 1.5 * 2^52, so the two's complement bits of the sum
                // sit in the mantissa (ulp = 1) and `(sum >> shift) as i32` is one funnel shift over the two words (shift <= 15
                // reaches bit 46 at most; the 2^51 of the bias is above that).  One instruction per tap on the FP64 pipe, which
                // this kernel leaves idle, instead of IMAD.WIDE + carry adds on the two integer pipes that bound it.  Inputs are
                // the outer loop: consecutive DFMAs go to different accumulators.
#pragma unroll
                for (int half = 0; half < 16; half += A3_FIR_OUT) {
                    double acc[A3_FIR_OUT];
#pragma unroll
                    for (int e = 0; e < A3_FIR_OUT; e++) acc[e] = 6755399441055744.0;
#pragma unroll
                    for (int i = half - HB; i < half + A3_FIR_OUT - 1; i++) {
                        const int32_t v = i >= 0 ? x[i >= 0 ? i : 0] : h[16 + i >= 0 ? 16 + i : 0];
                        const double dv = a3_i2d(v);
#pragma unroll
                        for (int e = 0; e < A3_FIR_OUT; e++) {
                            const int j = half + e - 1 - i;   // tap of output half + e that reads input i
                            if (j >= 0 && j < HB) acc[e] = fma(qd[j >= 0 && j < HB ? j : 0], dv, acc[e]);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < A3_FIR_OUT; e++) {
                        const int32_t pred = (int32_t)__funnelshift_r((uint32_t)__double2loint(acc[e]), (uint32_t)__double2hiint(acc[e]), shift);
                        const int32_t xe = x[half + e];
                        const int32_t rr = (int32_t)((uint32_t)xe - (uint32_t)pred);
                        rl[half + e] = rr;
                        oacc |= (uint32_t)((xe ^ pred) & (xe ^ rr));   // checked_sub: sign bit set when it overflowed
                        sacc |= (uint32_t)rr + 0x8000u;
                    }
                }
#endif
                if (stage < 2) {
                    if (((oacc >> 31) | (sacc >> 16)) != 0) {   // rare: look again, only samples in [order, n) count
                        const uint32_t first = order > i0 ? min(order - i0, 16u) : 0u, last = min(16u, n - i0);
#pragma unroll
                        for (int e = 0; e < 16; e++) {
                            if ((uint32_t)e < first || (uint32_t)e >= last) continue;
                            const int32_t pred = (int32_t)((uint32_t)x[e] - (uint32_t)rl[e]);
                            if (((x[e] ^ pred) & (x[e] ^ rl[e])) < 0) ovf = 1;
                            if (((uint32_t)rl[e] + 0x8000u) >> 16) b16 = 1;
                        }
                    }
                    // park the residuals as int16 pairs (two 16-byte chunks per tile, conflict-free for 32 consecutive tiles)
                    uint4 lo4, hi4;
                    lo4.x = __byte_perm((uint32_t)rl[0], (uint32_t)rl[1], 0x5410); lo4.y = __byte_perm((uint32_t)rl[2], (uint32_t)rl[3], 0x5410);
                    lo4.z = __byte_perm((uint32_t)rl[4], (uint32_t)rl[5], 0x5410); lo4.w = __byte_perm((uint32_t)rl[6], (uint32_t)rl[7], 0x5410);
                    hi4.x = __byte_perm((uint32_t)rl[8], (uint32_t)rl[9], 0x5410); hi4.y = __byte_perm((uint32_t)rl[10], (uint32_t)rl[11], 0x5410);
                    hi4.z = __byte_perm((uint32_t)rl[12], (uint32_t)rl[13], 0x5410); hi4.w = __byte_perm((uint32_t)rl[14], (uint32_t)rl[15], 0x5410);
                    res16[(rd * 2 + 0) * 32 + lane] = lo4;
                    res16[(rd * 2 + 1) * 32 + lane] = hi4;
                }
            } else if (stage == 2) {   // pass 2, LPC sweep: the int16 copy
                const uint4 lo4 = res16[(rd * 2 + 0) * 32 + lane], hi4 = res16[(rd * 2 + 1) * 32 + lane];
                const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    rl[2 * k] = (int32_t)(w[k] << 16) >> 16;
                    rl[2 * k + 1] = (int32_t)w[k] >> 16;
                }
            }
            // ---- fixed differences (:3039-3060); <= 28-bit samples cannot overflow i32 up to order 4 ----
            int32_t p1 = h[15] - h[14], p2 = p1 - (h[14] - h[13]), p3 = p2 - ((h[14] - h[13]) - (h[13] - h[12]));
            if (stage < 2) {
                const uint32_t chunk = udiv(i0, dcf);
                if (!tail && (cf16 || udiv(i0 + 15, dcf) == chunk)) {
                    // per-tile sums in 32 bits: |e_k| < 2^(28 + k) would overflow over 16 samples only for k >= 3, which are
                    // summed in two halves; the LPC residual is unbounded and keeps a 64-bit accumulator
                    uint32_t a0 = 0, a1 = 0, a2 = 0, a3a = 0, a3b = 0, a4a = 0, a4b = 0;
                    unsigned long long al = 0;
                    int32_t prev = h[15];
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const int32_t e1 = x[e] - prev, e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
                        prev = x[e]; p1 = e1; p2 = e2; p3 = e3;
                        a0 += uabs32(x[e]); a1 += uabs32(e1); a2 += uabs32(e2);
                        if (e < 8) { a3a += uabs32(e3); a4a += uabs32(e4); }
                        else { a3b += uabs32(e3); a4b += uabs32(e4); }
                        if (lpc_ok) al = acc_u32(al, uabs32(rl[e]));
                    }
                    unsigned long long s0 = a0, s1 = a1, s2 = a2, s3 = (unsigned long long)a3a + a3b, s4 = (unsigned long long)a4a + a4b;
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
                        s1 -= f0;
                        s2 -= f0 + f21;
                        s3 -= f0 + f31 + f32;
                        s4 -= f0 + f41 + f42 + f43;
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
                    aw_add_limbs(sm.limb_lo[0], sm.limb_hi[0], chunk, s0);
                    aw_add_limbs(sm.limb_lo[1], sm.limb_hi[1], chunk, s1);
                    aw_add_limbs(sm.limb_lo[2], sm.limb_hi[2], chunk, s2);
                    aw_add_limbs(sm.limb_lo[3], sm.limb_hi[3], chunk, s3);
                    aw_add_limbs(sm.limb_lo[4], sm.limb_hi[4], chunk, s4);
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
                if (stage == 3) {
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
                const uint32_t pf = udiv(i0, dcpf);
                const uint32_t codef = chf.rice[pf - j0f >= chf.nparts ? 0 : pf - j0f];
                if (!tail && i0 != 0 && (cpf16 || udiv(i0 + 15, dcpf) == pf) && codef < 0x40) {
                    uint32_t ta[4] = {0, 0, 0, 0};   // |rf| <= 2^28 here (25-bit samples): four shifted zig-zags fit 32 bits
#pragma unroll
                    for (int e = 0; e < 16; e++) ta[e >> 2] += zigzag32(rf[e]) >> codef;
                    bits_f += ((unsigned long long)ta[0] + ta[1]) + ((unsigned long long)ta[2] + ta[3]) + 16u * (1u + codef);
                } else if (!tail && i0 == 0 && cpf >= 16 && codef < 0x40) {
                    // the block's first tile: its first `fo` samples are warm-up, the rest belongs to partition 0 (one lane
                    // per candidate takes this branch; through the sample-by-sample fallback it held its warp for a whole
                    // extra round)
                    uint32_t ta[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int e = 0; e < 16; e++) ta[e >> 2] += (uint32_t)e >= fo ? zigzag32(rf[e]) >> codef : 0u;
                    bits_f += ((unsigned long long)ta[0] + ta[1]) + ((unsigned long long)ta[2] + ta[3]) + (16u - fo) * (1u + codef);
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
                }
                if (lpc_ok && (stage == 2 || !use16)) {
                    const RiceChoice& chl = sm.choice[1];
                    const uint32_t pl = udiv(i0, dcpl);
                    const uint32_t codel = chl.rice[pl - j0l >= chl.nparts ? 0 : pl - j0l];
                    if (!tail && i0 != 0 && (cpl16 || udiv(i0 + 15, dcpl) == pl) && codel < 0x40) {
                        if (use16) {   // |rl| < 2^15: the whole tile fits 32 bits
                            uint32_t ta = 0;
#pragma unroll
                            for (int e = 0; e < 16; e++) ta += zigzag32(rl[e]) >> codel;
                            bits_l += ta + 16u * (1u + codel);
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; e++) bits_l = acc_u32(bits_l, zigzag32(rl[e]) >> codel);
                            bits_l += 16u * (1u + codel);
                        }
                    } else if (!tail && i0 == 0 && cpl >= 16 && order <= 16 && codel < 0x40) {   // first tile: `order` warm-up samples
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
            if (stage >= 2 && rd < 8) {   // (all lanes of the warp are here: the body above is a plain `if`)
                const uint32_t rf_ = __reduce_add_sync(0xffffffffu, (uint32_t)(bits_f - rb_f0)), rl_ = __reduce_add_sync(0xffffffffu, (uint32_t)(bits_l - rb_l0));
                if (lane == 0) {
                    if (stage == 3) sm.round_bits[0][rd] = rf_;
                    if (stage == 2 || !swept_l) sm.round_bits[1][rd] = rl_;
                }
            }
        }
        if (stage < 2) {
            mask = __reduce_or_sync(0xffffffffu, mask);
            ovf = __reduce_or_sync(0xffffffffu, ovf);
            b16 = __reduce_or_sync(0xffffffffu, b16);
            if (lane == 0) {
                if (mask) atomicOr(&sm.mask, mask);
                if (ovf) atomicOr(&sm.ovf, 1u);
                if (b16) atomicOr(&sm.bad16, 1u);
            }
            a3_pair_sync(cand);
            if (stage == 0) {
                const uint32_t m = sm.mask;
                wasted = (m == 0 || (m & 1u)) ? 0u : (uint32_t)__ffs((int)m) - 1u;
                if (wasted) a3_pair_sync(cand);   // both warps have read the flags before the next stage clears them
            }
        }
    }
    // ---- totals of both warps ----
    bits_f = warp_sum_u64(bits_f);
    bits_l = warp_sum_u64(bits_l);
    bad_f = __any_sync(0xffffffffu, bad_f) ? 1u : 0u;
    bad_l = __any_sync(0xffffffffu, bad_l) ? 1u : 0u;
    if (lane == 0) {
        if (bits_f) atomicAdd(&sm.bits_f, bits_f);
        if (bits_l) atomicAdd(&sm.bits_l, bits_l);
        if (bad_f) atomicOr(&sm.bad_f, 1u);
        if (bad_l) atomicOr(&sm.bad_l, 1u);
    }
    a3_pair_sync(cand);
    // The LPC residuals as the int16 copy the first pass parked, for k_pack3, which reads them back instead of unpacking the PCM
    // and running the FIR a second time (a quarter of its instructions).  Both warps copy (8 KB per candidate) before the
    // first one goes on to the decision; whether the subframe really is an LPC subframe is said by CandRec::pad0 below.
    if (gres16 != nullptr && lpc_ok && use16)
        for (uint32_t k = wsub * 32 + lane; k < rounds * 64; k += 32 * A3_WPC) gres16[k] = res16[k];
    if (wsub != 0) return;
    const unsigned long long tot_f = sm.bits_f + hf, tot_l = sm.bits_l + hl;
    const bool fixed_ok = sm.bad_f == 0 && !skip_f;   // (skipped: its exact size is not known, only that it exceeds the LPC size)
    if (sm.bad_l) lpc_ok = false;
    const uint32_t hdr_bits = 8 + wasted;   // pad + type + wasted flag (+ unary(wasted - 1)) (src/stream.rs:1397)
    const uint32_t fixed_bits = hdr_bits + fo * bps + (uint32_t)tot_f + 6;
    const uint32_t lpc_bits = hdr_bits + order * bps + 4 + 5 + order * lp.precision + (uint32_t)tot_l + 6;
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
    // k_pack3 may use the stored residuals: an LPC subframe whose residuals all fit 16 bits
    if (lane == 0) rec->pad0 = (gres16 != nullptr && pick == 1 && use16) ? 1 : 0;
}

template <bool STEREO>
__host__ __device__ constexpr int a3_cands() { return STEREO ? 4 : 2; }

template <bool STEREO>
__host__ __device__ constexpr size_t a3_smem_bytes()
{
    return (size_t)2 * A3_PLANE * 4 + (size_t)a3_cands<STEREO>() * A3_PLANE * 2 + (size_t)a3_cands<STEREO>() * sizeof(A3Cand);
}

// STEREO: grid = frames, block = 256 (4 candidates x 2 warps).  Otherwise: grid = frames * ceil(channels / 2), block = 128.
// (a 112-register build of this kernel -- room for a persistent two-warp k_lpc4 CTA per SM beside two of its CTAs, FP64 work
// under the integer work -- spills 128 bytes per thread and runs 27 % slower by itself, 24.5 -> 31.0 ms per step, and the
// two LPC warps per SM need 5 ms per launch group, longer than the integer kernels they hide behind: measured and dropped)
template <int HB, bool STEREO>
__global__ void __launch_bounds__(32 * A3_WPC * (STEREO ? 4 : 2), STEREO ? 2 : 4)
    k_analyze3(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm, const LpcRec* __restrict__ lpcs,
               CandRec* __restrict__ out, unsigned long long* __restrict__ abssum, uint4* __restrict__ gres16)
{
    constexpr int NC = a3_cands<STEREO>();
    extern __shared__ __align__(16) uint8_t a3_dyn[];
    int32_t* planes = reinterpret_cast<int32_t*>(a3_dyn);
    uint4* res16_all = reinterpret_cast<uint4*>(a3_dyn + (size_t)2 * A3_PLANE * 4);
    A3Cand* cands_sm = reinterpret_cast<A3Cand*>(a3_dyn + (size_t)2 * A3_PLANE * 4 + (size_t)NC * A3_PLANE * 2);
    __shared__ unsigned long long abs4[4];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t groups = STEREO ? 1u : (cfg.channels + 1u) / 2u;
    const uint32_t f = blockIdx.x / groups, g = blockIdx.x % groups;
    const FrameDesc d = descs[f];
    const uint32_t n = d.n;
    const uint32_t ch0 = STEREO ? 0u : 2u * g;
    const bool fast_modes = STEREO && (cfg.mode == MODE_FAST_MID_SIDE || cfg.mode == MODE_FAST_SIDE);
    if (tid < 4) abs4[tid] = 0;
    if (fast_modes) __syncthreads();
#if FLACB200_A3_PF
    // The frame that will run in this CTA's place (two CTAs per SM, frames start in index order) is asked into L2 now: its CTA
    // otherwise begins with all eight warps waiting a DRAM round trip for the 24 KB of PCM (10 % of this kernel's stall samples).
    FrameDesc dnext;
    const bool pf = STEREO && f + A3_PF_DIST < cfg.nframes && tid < 192;
    if (pf) dnext = descs[f + A3_PF_DIST];
#endif
    // ---- unpack the two source channels once (Frame::fill_from_buf, src/audio.rs:149-187) ----
    {
        unsigned long long sl = 0, sr = 0, smid = 0, sside = 0;
        const uint32_t ntiles = (n + 15) / 16;
        for (uint32_t t = tid; t < ntiles; t += blockDim.x) {
            int32_t a[16], b[16];
            const uint32_t i0 = t * 16;
            if (STEREO || cfg.channels == 1) {
                if (STEREO) load_thread_samples<2>(cfg, d, pcm, i0, 0, a, b);
                else load_thread_samples<1>(cfg, d, pcm, i0, 0, a, b);
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const bool ok = i0 + e < n;
                    a[e] = ok ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch0) : 0;
                    b[e] = (ok && ch0 + 1 < cfg.channels) ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch0 + 1) : 0;
                }
            }
            const uint32_t w = a3_tile_base(t);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                *reinterpret_cast<int4*>(planes + w + c * 128) = make_int4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
                if (STEREO || cfg.channels > 1) *reinterpret_cast<int4*>(planes + A3_PLANE + w + c * 128) = make_int4(b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]);
            }
            if (fast_modes) {   // correlate_channels abs sums (:2475-2503)
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    sl += uabs32(a[e]); sr += uabs32(b[e]); smid += uabs32((a[e] + b[e]) >> 1); sside += uabs32(a[e] - b[e]);
                }
            }
        }
        if (fast_modes) {
            sl = warp_sum_u64(sl); sr = warp_sum_u64(sr); smid = warp_sum_u64(smid); sside = warp_sum_u64(sside);
            if (lane == 0) { atomicAdd(&abs4[0], sl); atomicAdd(&abs4[1], sr); atomicAdd(&abs4[2], smid); atomicAdd(&abs4[3], sside); }
        }
    }
#if FLACB200_A3_PF
    if (pf && cfg.pcm_kind <= 1) {
        const unsigned long long fb = (unsigned long long)cfg.channels * cfg.bytes_per_sample;
        const unsigned long long off = (unsigned long long)tid * 128u;
        if (off < dnext.n * fb) asm volatile("prefetch.global.L2 [%0];" ::"l"(pcm + dnext.pcm_off * fb + off));
    }
#endif
    __syncthreads();
    const uint32_t cand = wid / A3_WPC, wsub = wid % A3_WPC;
    const uint32_t slot = STEREO ? cand : ch0 + cand;
    if (!STEREO && slot >= cfg.channels) return;
    CandRec* rec = out + (size_t)f * cfg.nslots + slot;
    if (STEREO) {
        if (fast_modes) {
            unsigned long long sums[4] = {abs4[0], abs4[1], abs4[2], abs4[3]};
            if (tid < 4) abssum[(size_t)f * 4 + tid] = sums[tid];
            if (!slot_active(cfg, sums, slot)) {
                if (wsub == 0 && lane == 0) { rec->type = 0xFF; rec->bits = 0; }
                return;
            }
        } else if (cfg.mode == MODE_EXH_SIDE && slot == 2) {
            if (wsub == 0 && lane == 0) { rec->type = 0xFF; rec->bits = 0; }
            return;
        }
    }
    const uint32_t full_bps = STEREO ? cand_bps(cfg, slot) : cfg.bps;
    const LpcRec lp = lpcs[(size_t)f * cfg.nslots + slot];
    a3_candidate<HB, STEREO>(cfg, d, planes, res16_all + (size_t)cand * (A3_PLANE * 2 / 16), slot, STEREO ? slot : cand, cand, wsub, full_bps, lp,
                             cands_sm[cand], rec, gres16 ? gres16 + ((size_t)f * cfg.nslots + slot) * (A3_PLANE * 2 / 16) : nullptr);
}

// the 32-bit per-tile sums of the fixed residuals (passes 1 and 2) need |x| < 2^24: 24-bit stereo with its 25-bit side
bool analyze3_ok(const EncCfg& cfg)
{
    const uint32_t widest = cfg.bps + (cfg.mode != MODE_INDEPENDENT ? 1u : 0u);
    return analyze_fast_ok(cfg) && widest <= 25;
}

cudaError_t launch_analyze3(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const LpcRec* lpcs, CandRec* cands,
                            unsigned long long* abssum, uint4* gres16, cudaStream_t st)
{
    const uint32_t hb = cfg.max_lpc_order ? (cfg.max_lpc_order + 3u) >> 2 : 1u;
#ifndef FLACB200_A3_PAD
#define FLACB200_A3_PAD 0
#endif
    constexpr size_t pad_ = FLACB200_A3_PAD;   // occupancy experiments (tools/build_variant.sh)
#define FLACB200_A3(HBV, ST)                                                                                                          \
    do {                                                                                                                              \
        const size_t smem_ = a3_smem_bytes<ST>() + pad_;                                                                              \
        cudaError_t e_ = cudaFuncSetAttribute(k_analyze3<HBV, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);          \
        if (e_ != cudaSuccess) return e_;                                                                                             \
        const uint32_t groups_ = ST ? 1u : (cfg.channels + 1u) / 2u;                                                                  \
        count_launch(), k_analyze3<HBV, ST><<<cfg.nframes * groups_, 32 * A3_WPC * (ST ? 4 : 2), smem_, st>>>(cfg, descs, pcm, lpcs, cands, abssum, gres16);  \
    } while (0)
    if (cfg.mode != MODE_INDEPENDENT) {
        switch (hb) {
        case 1: FLACB200_A3(4, true); break;
        case 2: FLACB200_A3(8, true); break;
        case 3: FLACB200_A3(12, true); break;
        default: FLACB200_A3(16, true); break;
        }
    } else {
        switch (hb) {
        case 1: FLACB200_A3(4, false); break;
        case 2: FLACB200_A3(8, false); break;
        case 3: FLACB200_A3(12, false); break;
        default: FLACB200_A3(16, false); break;
        }
    }
#undef FLACB200_A3
    return cudaGetLastError();
}


// =====================================================================================================================
// k_frame4: encode_frame (src/encode.rs:2259-2439) for one stereo frame in ONE CTA -- analysis, decision and packing.
//
// k_analyze3 + k_decide + k_scan + k_pack3 in one kernel: the unpacked planes and the int16 copy of the LPC residuals stay
// in shared memory from the analysis through the packing, so the packer neither unpacks the PCM again nor runs the FIR a
// second time (k_pack3 spent a quarter of its instructions on those), and every warp of the CTA packs: the four warps of a
// subframe first add up the code lengths of their rounds, then emit at the bit positions that follow from the round totals.
//   1. planes (as k_analyze3), then a3_candidate for L, R, M, S: two warps each; the winning encodings land in shared memory
//   2. thread 0: channel assignment, frame header, CRC-8, subframe bit offsets, frame size (decide_frame)
//   3. the frame's size is published for the frames behind it (decoupled look-back: one 64-bit word per frame holding a flag
//      and either the frame's own size or the inclusive prefix of all sizes up to it)
//   4. bit image in shared memory -- on top of the analysis scratch, which is dead by now --, CRC-16 folded over it
//   5. warp 0 resolves the frame's byte offset: it sums the sizes of the frames in front of it that are still in flight
//      until it meets one that knows its prefix (CTAs start in index order, so every predecessor is running or done)
//   6. coalesced copy-out
// =====================================================================================================================
enum : unsigned long long { F4_FLAG_SIZE = 1ull << 62, F4_FLAG_PREFIX = 2ull << 62, F4_VALUE = (1ull << 62) - 1 };

struct F4Static {
    Crc16Fold tabs;
    CandRec cr[4];
    FrameRec fr;
    unsigned long long abs4[4];
    unsigned long long out_off;
    uint32_t bad16[4];
    uint32_t round_bits[2][8];   // per subframe: bits of every round of its residual block, partition headers included
    uint32_t crc_part[8];
    uint32_t placed;             // the look-back has succeeded
};

// sample i of candidate `slot` (0 L, 1 R, 2 mid, 3 side) from the tile-transposed planes
__device__ inline int32_t f4_sample(const int32_t* __restrict__ planes, uint32_t slot, uint32_t i)
{
    const uint32_t w = a3_tile_base(i >> 4) + ((i >> 2) & 3u) * 128u + (i & 3u);
    const int32_t a = planes[w], b = planes[A3_PLANE + w];
    return slot == 0 ? a : slot == 1 ? b : slot == 2 ? (a + b) >> 1 : a - b;
}

// the LPC residuals of tile t recomputed from the planes (only for a subframe whose residuals did not fit the int16 copy)
template <int HB>
static __device__ __noinline__ void f4_fir_tile(const int32_t* __restrict__ planes, uint32_t slot, uint32_t t, const int16_t* __restrict__ qc,
                                                uint32_t order, uint32_t shift, uint32_t wasted, int32_t* __restrict__ r)
{
    int32_t x[16], h[16], q[HB];
#pragma unroll
    for (int j = 0; j < HB; j++) q[j] = (uint32_t)j < order ? (int32_t)qc[j] : 0;
    a3_tile<true>(planes, slot, t, 0, x);
#pragma unroll
    for (int e = 0; e < 16; e++) h[e] = 0;
    if (t > 0) a3_tile<true>(planes, slot, t - 1, 4 - HB / 4, h);
#pragma unroll
    for (int e = 0; e < 16; e++) { x[e] >>= wasted; h[e] >>= wasted; }
#pragma unroll
    for (int e = 0; e < 16; e++) {
        long long sum = 0;
#pragma unroll
        for (int j = 0; j < HB; j++) sum = mad_wide_s32(e - 1 - j >= 0 ? x[e - 1 - j >= 0 ? e - 1 - j : 0] : h[16 + e - 1 - j >= 0 ? 16 + e - 1 - j : 0], q[j], sum);
        r[e] = (int32_t)((uint32_t)x[e] - (uint32_t)(unsigned long long)(sum >> shift));
    }
}

// One round (32 tiles, one per lane) of a subframe's residual block: residuals, code lengths, and -- EMIT -- the codes at
// the bit positions that start at `base`.  Returns the round's bit total (all lanes).
template <int HB, bool EMIT>
__device__ inline uint32_t f4_round(const int32_t* __restrict__ planes, const uint4* __restrict__ res16, uint32_t slot, const CandRec& cr, bool use16,
                                    uint32_t n, uint32_t rd, uint32_t words_sa, uint32_t base)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wasted = cr.wasted, order = cr.order;
    const bool lpc = cr.type == 3;
    const uint32_t cp = n >> cr.porder_g, j0 = (1u << cr.porder_g) - cr.nparts;
    const bool cp16 = (cp & 15u) == 0;
    const UDiv dcp = udiv_make(cp);
    const uint32_t hb = cr.method ? 5u : 4u, escape_code = cr.method ? 31u : 15u;
    const uint32_t t = rd * 32 + lane, i0 = t * 16;
    const bool live = i0 < n;
    int32_t r[16];
#pragma unroll
    for (int e = 0; e < 16; e++) r[e] = 0;
    if (live) {
        if (lpc) {
            if (use16) {   // the int16 copy k_analyze's first pass parked (two 16-byte chunks per tile)
                const uint4 lo4 = res16[(rd * 2 + 0) * 32 + lane], hi4 = res16[(rd * 2 + 1) * 32 + lane];
                const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    r[2 * k] = (int32_t)(w[k] << 16) >> 16;
                    r[2 * k + 1] = (int32_t)w[k] >> 16;
                }
            } else {
                int32_t tmp[16];
                f4_fir_tile<HB>(planes, slot, t, cr.q, order, cr.shift, wasted, tmp);
#pragma unroll
                for (int e = 0; e < 16; e++) r[e] = tmp[e];
            }
        } else {   // fixed differences (:3039-3060) from the planes
            int32_t x[16], h[16];
            a3_tile<true>(planes, slot, t, 0, x);
#pragma unroll
            for (int e = 0; e < 16; e++) h[e] = 0;
            if (t > 0) a3_tile<true>(planes, slot, t - 1, 3, h);
            int32_t x1 = h[15] >> wasted, x2 = h[14] >> wasted, x3 = h[13] >> wasted, x4 = h[12] >> wasted;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int32_t x0 = x[e] >> wasted;
                r[e] = order == 0 ? x0 : order == 1 ? x0 - x1 : order == 2 ? x0 - 2 * x1 + x2 : order == 3 ? x0 - 3 * x1 + 3 * x2 - x3
                                                                                                           : x0 - 4 * x1 + 6 * x2 - 4 * x3 + x4;
                x4 = x3; x3 = x2; x2 = x1; x1 = x0;
            }
        }
    }
    // ---- code lengths of this lane's tile (as k_pack3) ----
    const uint32_t lo_i = max(i0, order), hi_i = min(i0 + 16u, n);   // residuals exist for [lo_i, hi_i)
    const uint32_t pj = live ? udiv(i0, dcp) : 0u;
    const bool uniform = live && (i0 >= order || (i0 == 0 && order <= 16)) && i0 + 16 <= n && (cp16 || udiv(i0 + 15, dcp) == pj);
    const uint32_t cc0 = cr.rice[live ? min(pj - j0, (uint32_t)MAX_PARTS - 1) : 0u];
    const uint32_t first_res = max(pj * cp, order);
    const uint32_t skip = first_res > i0 ? min(first_res - i0, 16u) : 0u;
    const bool hdr_here = first_res >= i0 && first_res < i0 + 16;
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
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (!EMIT) return total;
    uint32_t p = base + incl - tsum;
    if (uniform && cc0 < 0x40) {
        if (hdr_here) {
            p3_put(words_sa, p, hb, cc0);
            p += hb;
        }
        const uint32_t stop = 1u << cc0, mask = stop - 1u;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const uint32_t u = zigzag32(r[e]);
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
    return total;
}

// Decoupled look-back (one warp): the byte offset of frame f = base + the sizes of all frames in front of it.  Every frame
// publishes its size as soon as it is known and its inclusive prefix as soon as its own look-back has succeeded; the walk
// goes back over published sizes until it meets a prefix.  spin = false: give up when a word that is needed has not been
// published yet (the caller tries again later -- and publishes its prefix at the first success, so resolution spreads
// forward while everybody is still packing).  Frames start in index order, so a spinning caller always makes progress.
__device__ inline bool f4_lookback(unsigned long long* state, uint32_t f, const unsigned long long* __restrict__ base_in, bool spin,
                                   unsigned long long* excl_out)
{
    const uint32_t lane = threadIdx.x & 31;
    if (f == 0) {
        *excl_out = *base_in;
        return true;
    }
    // 128 predecessors per step: lane l looks at the four frames at distances 4 l + 1 .. 4 l + 4 (nearest first) -- a whole
    // wave of resident CTAs is covered in two or three steps instead of ten
    unsigned long long excl = 0;
    long long i = (long long)f - 1;   // nearest frame not yet accounted for
    for (;;) {
        unsigned long long mine;      // sizes of this lane's frames in front of its first prefix (that prefix included)
        uint32_t have, stop;
        bool found;
        for (;;) {
            unsigned long long v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long idx = i - (long long)(4 * lane + k);
                v[k] = idx >= 0 ? *reinterpret_cast<volatile unsigned long long*>(state + idx) : F4_FLAG_PREFIX;   // (in front of frame 0: nothing)
            }
            mine = 0;
            found = false;
            bool missing = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t fl = (uint32_t)(v[k] >> 62);
                if (!found) {
                    if (fl == 0) missing = true;
                    else mine += v[k] & F4_VALUE;
                    if (fl == 2) found = true;
                }
            }
            have = __ballot_sync(0xffffffffu, found);
            stop = have ? (uint32_t)__ffs((int)have) - 1u : 32u;   // nearest lane that holds a prefix
            if (!__any_sync(0xffffffffu, lane <= stop && missing)) break;
            if (!spin) return false;
        }
        excl += warp_sum_u64(lane <= stop ? mine : 0ull);
        if (have) break;   // (the virtual frames in front of frame 0 carry the prefix flag with value 0: frame 0's own word is a prefix)
        i -= 128;
    }
    *excl_out = excl;
    return true;
}

// grid = frames of the launch group, block = 256 (4 candidates x 2 warps).  state: one word per frame, zeroed before the
// launch; base_in / total_out: running byte total of the call before / after this launch group.
template <int HB>
__global__ void __launch_bounds__(256, 2)
    k_frame4(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm, const LpcRec* __restrict__ lpcs,
             CandRec* __restrict__ cands_out, FrameRec* __restrict__ frecs_out, unsigned long long* __restrict__ abssum, unsigned long long* state,
             const unsigned long long* __restrict__ base_in, unsigned long long* __restrict__ total_out, unsigned long long* __restrict__ mapped_total,
             unsigned long long* __restrict__ err_word, uint32_t* __restrict__ frame_bytes_out, uint8_t* __restrict__ out)
{
    extern __shared__ __align__(16) uint8_t a3_dyn[];
    int32_t* planes = reinterpret_cast<int32_t*>(a3_dyn);
    uint4* res16_all = reinterpret_cast<uint4*>(a3_dyn + (size_t)2 * A3_PLANE * 4);
    A3Cand* cands_sm = reinterpret_cast<A3Cand*>(a3_dyn + (size_t)2 * A3_PLANE * 4 + (size_t)4 * A3_PLANE * 2);
    uint32_t* words = reinterpret_cast<uint32_t*>(cands_sm);   // the frame image reuses the analysis scratch
    constexpr uint32_t cap_words = (uint32_t)(4 * sizeof(A3Cand) / 4);
    __shared__ F4Static S;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t f = blockIdx.x;
    const FrameDesc d = descs[f];
    const uint32_t n = d.n;
    const bool fast_modes = cfg.mode == MODE_FAST_MID_SIDE || cfg.mode == MODE_FAST_SIDE;
    if (tid < 4) S.abs4[tid] = 0;
    for (uint32_t i = tid; i < sizeof(Crc16Fold) / 4; i += 256) reinterpret_cast<uint32_t*>(&S.tabs)[i] = reinterpret_cast<const uint32_t*>(&g_crc16_tabs)[i];
    if (fast_modes) __syncthreads();
    // ---- 1a. unpack the two source channels once (Frame::fill_from_buf, src/audio.rs:149-187) ----
    {
        unsigned long long sl = 0, sr = 0, smid = 0, sside = 0;
        const uint32_t ntiles = (n + 15) / 16;
        for (uint32_t t = tid; t < ntiles; t += 256) {
            int32_t a[16], b[16];
            load_thread_samples<2>(cfg, d, pcm, t * 16, 0, a, b);
            const uint32_t w = a3_tile_base(t);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                *reinterpret_cast<int4*>(planes + w + c * 128) = make_int4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
                *reinterpret_cast<int4*>(planes + A3_PLANE + w + c * 128) = make_int4(b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]);
            }
            if (fast_modes) {   // correlate_channels abs sums (:2475-2503)
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    sl += uabs32(a[e]); sr += uabs32(b[e]); smid += uabs32((a[e] + b[e]) >> 1); sside += uabs32(a[e] - b[e]);
                }
            }
        }
        if (fast_modes) {
            sl = warp_sum_u64(sl); sr = warp_sum_u64(sr); smid = warp_sum_u64(smid); sside = warp_sum_u64(sside);
            if (lane == 0) { atomicAdd(&S.abs4[0], sl); atomicAdd(&S.abs4[1], sr); atomicAdd(&S.abs4[2], smid); atomicAdd(&S.abs4[3], sside); }
        }
    }
    __syncthreads();
    // ---- 1b. the four candidates ----
    {
        const uint32_t cand = wid / A3_WPC, wsub = wid % A3_WPC, slot = cand;
        CandRec* rec = &S.cr[cand];
        bool active = true;
        if (fast_modes) {
            unsigned long long sums[4] = {S.abs4[0], S.abs4[1], S.abs4[2], S.abs4[3]};
            if (tid < 4 && abssum) abssum[(size_t)f * 4 + tid] = sums[tid];
            active = slot_active(cfg, sums, slot);
        } else if (cfg.mode == MODE_EXH_SIDE && slot == 2) {
            active = false;
        }
        if (!active) {
            if (wsub == 0 && lane == 0) { rec->type = 0xFF; rec->bits = 0; }
        } else {
            const LpcRec lp = lpcs[(size_t)f * 4 + slot];
            a3_candidate<HB, true>(cfg, d, planes, res16_all + (size_t)cand * (A3_PLANE * 2 / 16), slot, slot, cand, wsub, cand_bps(cfg, slot), lp,
                                   cands_sm[cand], rec);
        }
    }
    __syncthreads();
    // ---- 2. decisions of the frame ----
    if (tid < 4) S.bad16[tid] = cands_sm[tid].bad16;
    if (tid == 0) {
        S.placed = 0;
        decide_frame(cfg, d, S.cr, S.abs4, S.fr);
        const uint32_t need = (S.fr.frame_bytes + 3) / 4 + 2;
        if (need > cap_words) S.fr.err = 1;
        if (S.fr.err) *err_word = 1;   // sticky error word read back by the host
        frame_bytes_out[f] = S.fr.frame_bytes;
        // ---- 3. publish the size (frame 0 knows its prefix at once) ----
        const unsigned long long mine = S.fr.frame_bytes;
        *reinterpret_cast<volatile unsigned long long*>(state + f) = f == 0 ? (F4_FLAG_PREFIX | ((*base_in + mine) & F4_VALUE)) : (F4_FLAG_SIZE | mine);
    }
    if (cands_out)   // flacb200_encode_last_info
        for (uint32_t i = tid; i < 4 * sizeof(CandRec) / 4; i += 256) reinterpret_cast<uint32_t*>(cands_out + (size_t)f * 4)[i] = reinterpret_cast<const uint32_t*>(S.cr)[i];
    __syncthreads();
    if (tid < 16) {   // the code bits per round of the two chosen encodings, before the analysis scratch becomes the frame image
        const uint32_t sb = tid >> 3, rd = tid & 7;
        const CandRec& c = S.cr[S.fr.slot[sb]];
        S.round_bits[sb][rd] = c.type >= 2 ? cands_sm[S.fr.slot[sb]].round_bits[c.type == 3 ? 1 : 0][rd] : 0u;
    }
    __syncthreads();   // (also: every warp is done with the analysis scratch)
    const FrameRec& fr = S.fr;
    const uint32_t frame_bytes = fr.frame_bytes;
    // warp 0 tries to place the frame at several points of the packing; the first success publishes the prefix
    auto place = [&](bool spin) {
        if (S.placed) return;
        unsigned long long excl = 0;
        if (!f4_lookback(state, f, base_in, spin, &excl)) return;
        if (lane == 0) {
            const unsigned long long incl = excl + frame_bytes;
            if (f != 0) *reinterpret_cast<volatile unsigned long long*>(state + f) = F4_FLAG_PREFIX | (incl & F4_VALUE);
            S.out_off = excl;
            S.placed = 1;
            if (f + 1 == cfg.nframes) {
                *total_out = incl;
                if (mapped_total) *mapped_total = incl;
            }
            if (frecs_out) {
                FrameRec g = S.fr;
                g.out_off = excl;
                frecs_out[f] = g;
            }
        }
        __syncwarp();
    };
    if (wid == 0) place(false);
    const uint32_t words_sa = (uint32_t)__cvta_generic_to_shared(words);
    const uint32_t img_words = min((frame_bytes + 3) / 4 + 2, cap_words);
    const bool ok = fr.err == 0;
    // ---- 4. the bit image ----
    for (uint32_t i = tid; i < img_words; i += 256) words[i] = 0;
    __syncthreads();
    const uint32_t sub = wid >> 2, wq = wid & 3;   // subframe of this warp, warp within the subframe
    const CandRec& cr = S.cr[fr.slot[sub]];
    const uint32_t slot = fr.slot[sub];
    const uint32_t type = cr.type, order = type >= 2 ? cr.order : 0u, wasted = cr.wasted, bps = cr.bps;
    const bool use16 = S.bad16[slot] == 0;
    const uint32_t rounds = ((n + 15) / 16 + 31) / 32;
    uint32_t res_start = 0;
    if (ok) {
        if (tid < fr.hdr_len) p3_put(words_sa, 8 * tid, 8, fr.hdr[tid]);
        uint32_t pos = fr.sub_bit[sub];
        if (wq == 0 && lane == 0) {   // SubframeHeader (src/stream.rs:1397-1413)
            const uint32_t code = type == 0 ? 0u : type == 1 ? 1u : type == 2 ? 8u + order : 31u + order;
            p3_put_masked(words_sa, pos, 8, (code << 1) | (wasted ? 1u : 0u));
            if (wasted) p3_put(words_sa, pos + 8 + (wasted - 1), 1, 1);
        }
        pos += 8 + wasted;
        if (type == 0) {   // CONSTANT (:2982-2998)
            if (wq == 0 && lane == 0) p3_put_masked(words_sa, pos, bps, (uint32_t)(f4_sample(planes, slot, 0) >> wasted));
        } else if (type == 1) {   // VERBATIM (:3000-3018), all four warps
            for (uint32_t i = wq * 32 + lane; i < n; i += 128) p3_put_masked(words_sa, pos + i * bps, bps, (uint32_t)(f4_sample(planes, slot, i) >> wasted));
        } else {
            if (wq == 0 && lane < order) p3_put_masked(words_sa, pos + lane * bps, bps, (uint32_t)(f4_sample(planes, slot, lane) >> wasted));   // warm-up
            pos += order * bps;
            if (type == 3) {   // :3122-3133
                const uint32_t prec = cr.precision;
                if (wq == 0 && lane == 0) {
                    p3_put_masked(words_sa, pos, 4, prec - 1);
                    p3_put_masked(words_sa, pos + 4, 5, cr.shift);
                }
                if (wq == 1 && lane < order) p3_put_masked(words_sa, pos + 9 + lane * prec, prec, (uint32_t)(int32_t)cr.q[lane]);
                pos += 9 + order * prec;
            }
            if (wq == 2 && lane == 0) {   // residual block header (:3944-3961)
                p3_put_masked(words_sa, pos, 2, cr.method);
                p3_put_masked(words_sa, pos + 2, 4, cr.porder_w);
            }
            res_start = pos + 6;
            // the round totals of the analysis count the codes; the partition headers (src/stream.rs:1603-1619) ride in front of
            // the first residual of their partition
            if (wq == 3) {
                const uint32_t cp = n >> cr.porder_g, j0 = (1u << cr.porder_g) - cr.nparts, hb = cr.method ? 5u : 4u;
                for (uint32_t j = lane; j < cr.nparts; j += 32) {
                    const uint32_t first = max((j + j0) * cp, order);
                    if (first < n) atomicAdd(&S.round_bits[sub][first >> 9], cr.rice[j] < 0x40 ? hb : hb + 5u);
                }
            }
        }
    }
    __syncthreads();
    if (ok && type >= 2) {
        for (uint32_t rd = wq; rd < rounds; rd += 4) {
            uint32_t base = res_start;
            for (uint32_t k = 0; k < rd; k++) base += S.round_bits[sub][k];
            f4_round<HB, true>(planes, res16_all + (size_t)slot * (A3_PLANE * 2 / 16), slot, cr, use16, n, rd, words_sa, base);
        }
    }
    if (wid == 0) place(false);
    __syncthreads();
    // ---- CRC-16 over everything but the last two bytes (src/encode.rs:2408-2409) ----
    const uint32_t body = frame_bytes - 2, bw = body >> 2, btail = body & 3;
    const uint32_t per = (((bw + 7) / 8) + 63u) & ~63u;   // words per warp, whole rounds of 32 pairs
    if (ok) {
        const uint32_t a = min(wid * per, bw), b = min(a + per, bw);
        uint32_t part = p3_crc_words(S.tabs, words, a, b - a);
        if (lane == 0) {   // every warp shifts its own part to the end of the whole words: x^(32 * words behind it)
            const uint32_t after = bw - b;
            if (b > a && after) {
                part = gf16_mulmod(part, g_crc16_xblk[after >> 5]);
                if (after & 31u) part = gf16_mulmod(part, S.tabs.xd[after & 31u]);
            }
            S.crc_part[wid] = b > a ? part : 0u;
        }
    }
    __syncthreads();
    if (ok && tid == 0) {
        uint32_t crc = 0;
        for (uint32_t w = 0; w < 8; w++) crc ^= S.crc_part[w];
        for (uint32_t t = 0; t < btail; t++) crc = (S.tabs.T[0][((crc >> 8) ^ (words[bw] >> (24 - 8 * t))) & 0xff] ^ (crc << 8)) & 0xffffu;
        p3_put(words_sa, body * 8, 16, crc);
    }
    // ---- 5. where the frame goes ----
    if (wid == 0) place(true);
    __syncthreads();
    if (!ok) return;
    // ---- 6. copy out (as k_pack3) ----
    const unsigned long long o = S.out_off;
    const unsigned long long A = (o + 3) & ~3ull, B = (o + frame_bytes) & ~3ull;
    auto frame_byte = [&](uint32_t b) -> uint8_t { return (uint8_t)(words[b >> 2] >> (24 - 8 * (b & 3))); };
    if (A >= B) {
        for (uint32_t b = tid; b < frame_bytes; b += 256) out[o + b] = frame_byte(b);
        return;
    }
    const uint32_t head = (uint32_t)(A - o), tail0 = (uint32_t)(B - o);
    if (tid < head) out[o + tid] = frame_byte(tid);
    if (tid < frame_bytes - tail0) out[B + tid] = frame_byte(tail0 + tid);
    uint32_t* gw = reinterpret_cast<uint32_t*>(out + A);
    const uint32_t nfull = (uint32_t)((B - A) >> 2);
    const uint32_t sh = head & 3;
    for (uint32_t j = tid; j < nfull; j += 256) {
        const uint32_t b = head + 4 * j, idx = b >> 2;
        const uint32_t be = sh ? __funnelshift_l(words[idx + 1], words[idx], 8 * sh) : words[idx];
        gw[j] = __byte_perm(be, 0, 0x0123);
    }
}

void init_fused_tables(cudaStream_t st) { k_crc16_tables_init<<<1, 256, 0, st>>>(); }

// stereo frames the CTA-per-frame analysis covers (k_analyze3's shapes with four candidate slots)
bool frame4_ok(const EncCfg& cfg) { return analyze3_ok(cfg) && cfg.mode != MODE_INDEPENDENT && cfg.nslots == 4; }

cudaError_t launch_frame4(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const LpcRec* lpcs, CandRec* cands_out, FrameRec* frecs_out,
                          unsigned long long* abssum, unsigned long long* state, const unsigned long long* base_in, unsigned long long* total_out,
                          unsigned long long* mapped_total, unsigned long long* err_word, uint32_t* frame_bytes_out, uint8_t* out, cudaStream_t st)
{
    const uint32_t hb = cfg.max_lpc_order ? (cfg.max_lpc_order + 3u) >> 2 : 1u;
    const size_t smem = a3_smem_bytes<true>();
    cudaError_t e = cudaMemsetAsync(state, 0, (size_t)cfg.nframes * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
#define FLACB200_F4(HBV)                                                                                                             \
    do {                                                                                                                             \
        e = cudaFuncSetAttribute(k_frame4<HBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                             \
        if (e != cudaSuccess) return e;                                                                                              \
        count_launch(), k_frame4<HBV><<<cfg.nframes, 256, smem, st>>>(cfg, descs, pcm, lpcs, cands_out, frecs_out, abssum, state, base_in, total_out, \
                                                      mapped_total, err_word, frame_bytes_out, out);                                 \
    } while (0)
    switch (hb) {
    case 1: FLACB200_F4(4); break;
    case 2: FLACB200_F4(8); break;
    case 3: FLACB200_F4(12); break;
    default: FLACB200_F4(16); break;
    }
#undef FLACB200_F4
    return cudaGetLastError();
}

}   // namespace flacb200
