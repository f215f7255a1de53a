// encode_kernels.cu -- sm_100a kernels of the FLAC frame encoder.
//
// Pipeline per launch group (F frames, S candidate slots each), all on one stream:
//   k_planes    PCM bytes -> planar int32 candidates (+ mid/side), OR masks, abs sums        [HBM-bound]
//   k_lpc       warp per candidate: window, FP64 autocorrelation as strict sequential chains (one lag
//               per lane, bit-identical to the reference's left-to-right sums), Levinson-Durbin,
//               order estimate, coefficient quantisation                                     [FP64 latency]
//   k_residual  CTA per candidate: fixed orders 0-4, LPC residual (INT64 MAC), partitioned Rice
//               search, exact bit counts, fixed/LPC/verbatim/constant decision               [INT32/INT64]
//   k_decide    thread per frame: channel assignment by minimum bits, frame header + CRC-8, frame size
//   k_scan      exclusive scan of frame sizes -> output offsets (single CTA)
//   k_zero      clears the output range of the group (bit writes are ORs)
//   k_pack      CTA per emitted subframe: residual recomputation, scan of code lengths, parallel
//               bit packing through shared memory                                            [INT32/shared]
//   k_crc16     CTA per frame: chunked CRC-16 combined with x^(8 len) mod P
//
// Reference statements followed by each kernel are cited inline (paths relative to flac-codec 1.3.2).
#include <type_traits>

#include "common.cuh"
#include "glibc_log.cuh"
#include "frame_decide.cuh"
#include "crc.cuh"
#include "tiles.cuh"
#include "rice.cuh"

namespace flacb200 {

// grid (ceil(block_size / 256), F), block 256
__global__ void __launch_bounds__(256) k_planes(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
                                                int32_t* __restrict__ planes, uint32_t* __restrict__ ormask,
                                                unsigned long long* __restrict__ abssum)
{
    const uint32_t f = blockIdx.y;
    const FrameDesc d = descs[f];
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const bool valid = i < d.n;
    int32_t* base = planes + (size_t)f * cfg.nslots * cfg.bpad;
    if (cfg.mode == MODE_INDEPENDENT) {
        for (uint32_t c = 0; c < cfg.channels; c++) {
            int32_t v = valid ? load_pcm_sample(pcm, cfg, d.pcm_off + i, c) : 0;
            if (valid) base[(size_t)c * cfg.bpad + i] = v;
            uint32_t m = __reduce_or_sync(0xffffffffu, (uint32_t)v);
            if ((threadIdx.x & 31) == 0 && m) atomicOr(&ormask[f * cfg.nslots + c], m);
        }
        return;
    }
    // stereo with a side channel: L, R, M = (L + R) >> 1 (:2721), S = L - R (:2734)
    int32_t l = 0, r = 0;
    if (valid) {
        l = load_pcm_sample(pcm, cfg, d.pcm_off + i, 0);
        r = load_pcm_sample(pcm, cfg, d.pcm_off + i, 1);
    }
    const int32_t mid = (l + r) >> 1, side = l - r;
    const bool want_mid = cfg.mode == MODE_EXH_MID_SIDE || cfg.mode == MODE_FAST_MID_SIDE;
    if (valid) {
        base[i] = l;
        base[cfg.bpad + i] = r;
        if (want_mid) base[2 * (size_t)cfg.bpad + i] = mid;
        base[3 * (size_t)cfg.bpad + i] = side;
    }
    uint32_t ml = __reduce_or_sync(0xffffffffu, (uint32_t)l), mr = __reduce_or_sync(0xffffffffu, (uint32_t)r);
    uint32_t mm = __reduce_or_sync(0xffffffffu, (uint32_t)mid), ms = __reduce_or_sync(0xffffffffu, (uint32_t)side);
    const bool lane0 = (threadIdx.x & 31) == 0;
    if (lane0) {
        if (ml) atomicOr(&ormask[f * 4 + 0], ml);
        if (mr) atomicOr(&ormask[f * 4 + 1], mr);
        if (mm && want_mid) atomicOr(&ormask[f * 4 + 2], mm);
        if (ms) atomicOr(&ormask[f * 4 + 3], ms);
    }
    if (cfg.mode == MODE_FAST_MID_SIDE || cfg.mode == MODE_FAST_SIDE) {   // correlate_channels abs sums (:2475-2503)
        unsigned long long sl = warp_sum_u64(uabs32(l)), sr = warp_sum_u64(uabs32(r));
        unsigned long long sm = warp_sum_u64(uabs32(mid)), ss = warp_sum_u64(uabs32(side));
        if (lane0) {
            atomicAdd(&abssum[f * 4 + 0], sl);
            atomicAdd(&abssum[f * 4 + 1], sr);
            atomicAdd(&abssum[f * 4 + 2], sm);
            atomicAdd(&abssum[f * 4 + 3], ss);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_lpc: LpcParameters::best  (src/encode.rs:3292-3332)
// ------------------------------------------------------------------------------------------------
constexpr int LPC_WARPS = 4;
constexpr int RING = 128;   // doubles per warp: four 32-sample tiles

struct LpcWarpSmem {
    double ring[RING];
    double R[MAX_LPC + 1];
    double ca[MAX_LPC], cb[MAX_LPC], err[MAX_LPC];
};

// lp_coefficients (src/encode.rs:3536-3580) up to order `upto`; the order-`upto` set ends in *cur
__device__ inline void levinson(const double* R, int upto, double* ca, double* cb, double* err, double** cur)
{
    double k = __ddiv_rn(R[1], R[0]);                                        // :3545
    ca[0] = k;
    err[0] = __dmul_rn(R[0], __dsub_rn(1.0, __dmul_rn(k, k)));               // :3548
    double* a = ca;
    double* b = cb;
    for (int i = 1; i < upto; i++) {
        double s = -0.0;
        for (int j = 0; j < i; j++) s = __dadd_rn(s, __dmul_rn(R[i - j], a[j]));   // :3555-3561
        const double q = __dsub_rn(R[i + 1], s);
        k = __ddiv_rn(q, err[i - 1]);                                         // :3563
        for (int j = 0; j < i; j++) b[j] = __dsub_rn(a[j], __dmul_rn(k, a[i - 1 - j]));   // :3566-3569
        b[i] = k;
        err[i] = __dmul_rn(err[i - 1], __dsub_rn(1.0, __dmul_rn(k, k)));      // :3572
        double* t = a; a = b; b = t;
    }
    *cur = a;
}

// grid ceil(ncand / LPC_WARPS), block 32 * LPC_WARPS
__global__ void __launch_bounds__(32 * LPC_WARPS) k_lpc(EncCfg cfg, const FrameDesc* __restrict__ descs, const int32_t* __restrict__ planes,
                                                       const uint32_t* __restrict__ ormask, const unsigned long long* __restrict__ abssum,
                                                       const double* __restrict__ winpool, LpcRec* __restrict__ out, uint32_t ncand)
{
    __shared__ LpcWarpSmem sm_all[LPC_WARPS];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t cand = blockIdx.x * LPC_WARPS + wid;
    if (cand >= ncand) return;
    LpcWarpSmem& sm = sm_all[wid];
    const uint32_t f = cand / cfg.nslots, slot = cand % cfg.nslots;
    const FrameDesc d = descs[f];
    const uint32_t n = d.n, M = cfg.max_lpc_order;
    if (lane == 0) out[cand].ok = 0;
    const uint32_t mask = ormask[cand];
    // inactive candidate, LPC disabled, InsufficientLpcSamples (:3300), or CONSTANT (:2883)
    if (M == 0 || n <= M || mask == 0 || !slot_active(cfg, abssum + (size_t)f * 4, slot)) return;
    const uint32_t wasted = (mask & 1u) ? 0u : (uint32_t)__ffs((int)mask) - 1u;   // :2878-2898
    const uint32_t bps = cand_bps(cfg, slot) - wasted;
    const uint32_t precision = lpc_precision_for(n);
    const int32_t* x = planes + (size_t)cand * cfg.bpad;
    const double* win = winpool + d.win_off;

    // ---- windowing (Window::apply :1799) + autocorrelate (:3478-3501).  Lane l owns lag l and runs the
    // strict left-to-right sum  s = s + xw[i] * xw[i + l]  with separately rounded multiply and add, so
    // R[] is bit-identical to the reference.  Samples past the block are zero: adding x*0 leaves s unchanged.
    auto load_tile = [&](uint32_t tile) {
        const uint32_t idx = tile * 32 + lane;
        double v = 0.0;
        if (idx < n) v = __dmul_rn((double)(x[idx] >> wasted), win[idx]);
        sm.ring[idx & (RING - 1)] = v;
    };
    load_tile(0);
    load_tile(1);
    load_tile(2);
    __syncwarp();
    double acc = -0.0, acc32 = -0.0;   // Iterator::sum::<f64>() folds from -0.0
    const uint32_t ntiles = (n + 31) / 32;
    for (uint32_t t = 0; t < ntiles; t++) {
        const uint32_t base = t * 32;
#pragma unroll 8
        for (uint32_t s = 0; s < 32; s++) {
            const uint32_t i = base + s;
            const double a = sm.ring[i & (RING - 1)];
            const double b = sm.ring[(i + lane) & (RING - 1)];
            acc = __dadd_rn(acc, __dmul_rn(a, b));
            if (M == 32) {
                const double b2 = sm.ring[(i + 32) & (RING - 1)];
                acc32 = __dadd_rn(acc32, __dmul_rn(a, b2));
            }
        }
        __syncwarp();
        load_tile(t + 3);   // overwrites tile t - 1's slot; tiles t+1, t+2 are already resident
        __syncwarp();
    }
    if (lane <= M) sm.R[lane] = acc;
    if (M == 32 && lane == 0) sm.R[32] = acc32;
    __syncwarp();

    if (lane == 0) {
        // lp_coefficients over all orders, then compute_best_order (:3688-3702)
        double* cur;
        levinson(sm.R, (int)M, sm.ca, sm.cb, sm.err, &cur);
        const double error_scale = __ddiv_rn(0.5, (double)n);   // :3664
        const double divisor = 2.0 * 0.693147180559945309417232121458176568;   // (2.0 * LN_2).max(0.0)
        int best = 0;
        double best_bits = 0.0;
        for (uint32_t o = 1; o <= M; o++) {
            const double e = sm.err[o - 1];
            if (!(e > 0.0)) break;   // take_while  :3668
            const double bpr = __ddiv_rn(glibc_log(__dmul_rn(e, error_scale)), divisor);   // :3674-3675
            const double bits = fma(bpr, (double)(n - o), (double)(o * (bps + precision)));   // :3677
            if (best == 0 || total_key(bits) < total_key(best_bits)) {
                best = (int)o;
                best_bits = bits;
            }
        }
        if (best == 0) return;   // NoBestLpcOrder
        levinson(sm.R, best, sm.ca, sm.cb, sm.err, &cur);
        // quantize (:3334-3401)
        double l = fabs(cur[0]);
        for (int j = 1; j < best; j++) {
            const double a = fabs(cur[j]);
            if (total_key(a) >= total_key(l)) l = a;
        }
        if (!(l > 0.0)) return;   // ZeroLpCoefficients
        const int32_t max_coeff = (1 << (precision - 1)) - 1, min_coeff = -(1 << (precision - 1));
        const int32_t lg = f64_as_i32_sat(floor(glibc_log2(l)));
        long long sh = (long long)((int32_t)precision - 1) - (long long)lg - 1;   // :3360
        if (sh > 15) sh = 15;
        if (sh < -16) return;     // LpNegativeShiftError
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

// ------------------------------------------------------------------------------------------------
// k_residual: encode_subframe (src/encode.rs:2849-2980) without emitting bits
// ------------------------------------------------------------------------------------------------
constexpr int RES_THREADS = 256;

struct ResSmem {
    unsigned long long red[RES_THREADS / 32];
    uint32_t red32[RES_THREADS / 32];
    unsigned long long chunk_sum[MAX_PARTS];
    uint32_t part_est[128];
    uint8_t part_code[128];   // rice[] encoding, 0xFF = Partition::new returned None, 0xFE = empty
    RiceChoice fixed, lpc;
    uint32_t flags;
};

// best_partitions + try_reduce_rice (src/encode.rs:3865-3942) and the exact size of the residual block.
// r: L residuals of a block of n samples with predictor order o.  All threads of the CTA must call.
__device__ void rice_search(const int32_t* r, uint32_t L, uint32_t o, uint32_t n, const EncCfg& cfg, ResSmem& sm, RiceChoice& out)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t rice_max = cfg.use_rice2 ? 31u : 15u;
    uint32_t p_max = n ? (uint32_t)__ffs((int)n) - 1u : 0u;
    if (p_max > cfg.max_porder) p_max = cfg.max_porder;
    if (p_max > MAX_PORDER) p_max = MAX_PORDER;
    const uint32_t cf = n >> p_max;            // finest chunk
    const bool pow2 = (cf & (cf - 1)) == 0;
    const uint32_t cf_shift = 31u - (uint32_t)__clz((int)cf);
    if (tid < MAX_PARTS) sm.chunk_sum[tid] = 0;
    __syncthreads();
    // partition abs sums at the finest order; chunks are aligned on absolute sample index (rchunks from the end)
    for (uint32_t i0 = (tid & ~31u); i0 < L; i0 += RES_THREADS) {
        const uint32_t i = i0 + (tid & 31);
        const bool v = i < L;
        const uint32_t a = v ? uabs32(r[i]) : 0u;
        const uint32_t last = min(i0 + 31, L - 1);
        const uint32_t ia = (v ? i : last) + o;
        const uint32_t m = pow2 ? (ia >> cf_shift) : (ia / cf);
        const uint32_t m_lo = pow2 ? ((i0 + o) >> cf_shift) : ((i0 + o) / cf);
        const uint32_t m_hi = pow2 ? ((last + o) >> cf_shift) : ((last + o) / cf);
        if (m_lo == m_hi) {
            const unsigned long long s = warp_sum_u64(a);
            if ((tid & 31) == 0 && s) atomicAdd(&sm.chunk_sum[m_lo], s);
        } else if (a) {
            atomicAdd(&sm.chunk_sum[m], (unsigned long long)a);
        }
    }
    __syncthreads();
    // one thread per (order p, partition j)
    if (tid < 127) {
        const uint32_t p = 31u - (uint32_t)__clz((int)(tid + 1));
        const uint32_t j = tid + 1 - (1u << p);
        uint8_t code = 0xFE;
        uint32_t est = 0;
        if (p <= p_max) {
            const uint32_t cp = n >> p;
            const uint32_t lo = j * cp, hi = lo + cp;   // absolute sample range of the partition
            if (hi > o) {
                const uint32_t span = 1u << (p_max - p);
                unsigned long long s = 0;
                for (uint32_t m = j * span; m < (j + 1) * span; m++) s += sm.chunk_sum[m];
                code = partition_code(s, hi - max(lo, o), rice_max, &est);
            }
        }
        sm.part_code[tid] = code;
        sm.part_est[tid] = est;
    }
    __syncthreads();
    if (tid == 0) {
        bool have = false;
        uint32_t best_est = 0, best_p = 0, best_count = 0;
        for (uint32_t p = 0; p <= p_max; p++) {
            const uint32_t base = (1u << p) - 1;
            uint32_t count = 0, est = 0;
            bool ok = true;
            for (uint32_t j = 0; j < (1u << p); j++) {
                const uint8_t c = sm.part_code[base + j];
                if (c == 0xFE) continue;
                if (c == 0xFF) { ok = false; break; }
                count++;
                est += sm.part_est[base + j];
            }
            if (!ok || count == 0 || (count & (count - 1))) continue;   // :3880-3881
            if (!have || est < best_est) { have = true; best_est = est; best_p = p; best_count = count; }   // first minimum :3885
        }
        out.fail = 0;
        if (!have) {   // unwrap_or_else (:3887): one partition escaped at 31 bits
            out.porder_g = 0; out.porder_w = 0; out.nparts = 1;
            out.rice[0] = 0x40 | 31;
            out.method = cfg.use_rice2 ? 1 : 0;
        } else {
            out.porder_g = (uint8_t)best_p;
            out.nparts = (uint8_t)best_count;
            out.porder_w = (uint8_t)(31u - (uint32_t)__clz((int)best_count));   // partitions.len().ilog2() :3902
            const uint32_t base = (1u << best_p) - 1, j0 = (1u << best_p) - best_count;
            bool shrink = true;
            for (uint32_t j = 0; j < best_count; j++) {
                const uint8_t c = sm.part_code[base + j0 + j];
                out.rice[j] = c;
                if (c < 0x40 && c >= 15) shrink = false;
            }
            out.method = (cfg.use_rice2 && !shrink) ? 1 : 0;   // try_reduce_rice :3929-3942
        }
    }
    __syncthreads();
    // exact size: what Partition::to_writer will emit (:3834-3863)
    const uint32_t cp = n >> out.porder_g;
    const uint32_t j0 = (1u << out.porder_g) - out.nparts;
    const bool cp_pow2 = (cp & (cp - 1)) == 0;
    const uint32_t cp_shift = 31u - (uint32_t)__clz((int)cp);
    unsigned long long bits = 0;
    uint32_t bad = 0;
    for (uint32_t i = tid; i < L; i += RES_THREADS) {
        const uint32_t ia = i + o;
        const uint32_t j = (cp_pow2 ? (ia >> cp_shift) : (ia / cp)) - j0;
        const uint8_t c = out.rice[j];
        const int32_t s = r[i];
        if (c < 0x40) bits += (zigzag32(s) >> c) + 1u + c;
        else if (c & 0x40) {
            const uint32_t w = c & 31u;
            bits += w;
            if (w < 32 && (s < -(1 << (w - 1)) || s > (1 << (w - 1)) - 1)) bad = 1;   // write_signed_counted fails
        }
    }
    const uint32_t hdr = out.method ? 5u : 4u;
    unsigned long long total = block_sum_u64(bits, sm.red);
    const uint32_t anybad = block_or_u32(bad, sm.red32);
    if (tid == 0) {
        uint32_t extra = 2 + 4;   // coding method + partition order
        for (uint32_t j = 0; j < out.nparts; j++) extra += (out.rice[j] < 0x40) ? hdr : hdr + 5;
        out.resid_bits = (uint32_t)total + extra;
        out.fail = anybad;
    }
    __syncthreads();
}

// grid ncand, block RES_THREADS, dynamic smem: 2 * bpad int32 when SMEM
template <bool SMEM>
__global__ void __launch_bounds__(RES_THREADS) k_residual(EncCfg cfg, const FrameDesc* __restrict__ descs, const int32_t* __restrict__ planes,
                                                         const uint32_t* __restrict__ ormask, const unsigned long long* __restrict__ abssum,
                                                         const LpcRec* __restrict__ lpcs, CandRec* __restrict__ out, int32_t* __restrict__ scratch)
{
    extern __shared__ __align__(16) int32_t dyn[];
    __shared__ ResSmem sm;
    const uint32_t cand = blockIdx.x, tid = threadIdx.x;
    const uint32_t f = cand / cfg.nslots, slot = cand % cfg.nslots;
    const uint32_t n = descs[f].n;
    CandRec* rec = out + cand;
    if (!slot_active(cfg, abssum + (size_t)f * 4, slot)) {
        if (tid == 0) { rec->type = 0xFF; rec->bits = 0; }
        return;
    }
    const uint32_t full_bps = cand_bps(cfg, slot);
    const uint32_t mask = ormask[cand];
    if (mask == 0) {   // all samples zero -> CONSTANT (:2870, :2883)
        if (tid == 0) {
            rec->type = 0; rec->order = 0; rec->wasted = 0; rec->bps = (uint8_t)full_bps;
            rec->bits = 8 + full_bps;
        }
        return;
    }
    const uint32_t wasted = (mask & 1u) ? 0u : (uint32_t)__ffs((int)mask) - 1u;
    const uint32_t bps = full_bps - wasted;
    const int32_t* plane = planes + (size_t)cand * cfg.bpad;
    int32_t* xs = SMEM ? dyn : scratch + (size_t)cand * 2 * cfg.bpad;
    int32_t* rs = xs + cfg.bpad;
    for (uint32_t i = tid; i < n; i += RES_THREADS) xs[i] = plane[i] >> wasted;   // :2891
    __syncthreads();

    // ---- encode_fixed_subframe (:3020-3088): orders 0..4 by successive differences ----
    // d_k[i] computed exactly in 64 bits; level k "overflows" (checked_sub fails) if any d_k[i], i >= k, leaves i32
    const uint32_t kmax = min(4u, n - 1);
    unsigned long long s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    uint32_t ovf = 0;
    for (uint32_t i = tid; i < n; i += RES_THREADS) {
        const long long a0 = xs[i];
        const long long a1 = i >= 1 ? xs[i - 1] : 0, a2 = i >= 2 ? xs[i - 2] : 0, a3 = i >= 3 ? xs[i - 3] : 0, a4 = i >= 4 ? xs[i - 4] : 0;
        const long long d1 = a0 - a1, d2 = a0 - 2 * a1 + a2, d3 = a0 - 3 * a1 + 3 * a2 - a3, d4 = a0 - 4 * a1 + 6 * a2 - 4 * a3 + a4;
        if (i >= 1 && (d1 < INT32_MIN || d1 > INT32_MAX)) ovf |= 1u << 1;
        if (i >= 2 && (d2 < INT32_MIN || d2 > INT32_MAX)) ovf |= 1u << 2;
        if (i >= 3 && (d3 < INT32_MIN || d3 > INT32_MAX)) ovf |= 1u << 3;
        if (i >= 4 && (d4 < INT32_MIN || d4 > INT32_MAX)) ovf |= 1u << 4;
        if (i >= kmax) {
            s0 += (unsigned long long)(a0 < 0 ? -a0 : a0);
            s1 += (unsigned long long)(d1 < 0 ? -d1 : d1);
            s2 += (unsigned long long)(d2 < 0 ? -d2 : d2);
            s3 += (unsigned long long)(d3 < 0 ? -d3 : d3);
            s4 += (unsigned long long)(d4 < 0 ? -d4 : d4);
        }
    }
    ovf = block_or_u32(ovf, sm.red32);
    uint32_t K = kmax;
    if (ovf) {   // drop the overflowing order and everything above it (:3045-3050), then redo the common tail
        uint32_t first = (uint32_t)__ffs((int)ovf) - 1u;
        if (first - 1 < K) K = first - 1;
        s0 = s1 = s2 = s3 = s4 = 0;
        for (uint32_t i = tid; i < n; i += RES_THREADS) {
            if (i < K) continue;
            const long long a0 = xs[i];
            const long long a1 = i >= 1 ? xs[i - 1] : 0, a2 = i >= 2 ? xs[i - 2] : 0, a3 = i >= 3 ? xs[i - 3] : 0;
            const long long d1 = a0 - a1, d2 = a0 - 2 * a1 + a2, d3 = a0 - 3 * a1 + 3 * a2 - a3;
            s0 += (unsigned long long)(a0 < 0 ? -a0 : a0);
            s1 += (unsigned long long)(d1 < 0 ? -d1 : d1);
            s2 += (unsigned long long)(d2 < 0 ? -d2 : d2);
            s3 += (unsigned long long)(d3 < 0 ? -d3 : d3);
        }
    }
    s0 = block_sum_u64(s0, sm.red);
    s1 = block_sum_u64(s1, sm.red);
    s2 = block_sum_u64(s2, sm.red);
    s3 = block_sum_u64(s3, sm.red);
    s4 = block_sum_u64(s4, sm.red);
    uint32_t fo = 0;   // first minimum (:3065-3075)
    {
        unsigned long long best = s0;
        if (K >= 1 && s1 < best) { best = s1; fo = 1; }
        if (K >= 2 && s2 < best) { best = s2; fo = 2; }
        if (K >= 3 && s3 < best) { best = s3; fo = 3; }
        if (K >= 4 && s4 < best) { best = s4; fo = 4; }
    }
    for (uint32_t i = fo + tid; i < n; i += RES_THREADS) {
        const long long a0 = xs[i];
        long long d;
        switch (fo) {
        case 0: d = a0; break;
        case 1: d = a0 - xs[i - 1]; break;
        case 2: d = a0 - 2ll * xs[i - 1] + xs[i - 2]; break;
        case 3: d = a0 - 3ll * xs[i - 1] + 3ll * xs[i - 2] - xs[i - 3]; break;
        default: d = a0 - 4ll * xs[i - 1] + 6ll * xs[i - 2] - 4ll * xs[i - 3] + xs[i - 4]; break;
        }
        rs[i - fo] = (int32_t)d;
    }
    __syncthreads();
    rice_search(rs, n - fo, fo, n, cfg, sm, sm.fixed);
    const uint32_t hdr_bits = 8 + wasted;   // pad + type + wasted flag (+ unary(wasted - 1)) (src/stream.rs:1397)
    const bool fixed_ok = sm.fixed.fail == 0;
    const uint32_t fixed_bits = hdr_bits + fo * bps + sm.fixed.resid_bits;

    // ---- encode_lpc_subframe (:3090-3136) ----
    const LpcRec lp = lpcs[cand];
    bool lpc_ok = lp.ok != 0;
    uint32_t lpc_bits = 0;
    if (lpc_ok) {   // uniform across the CTA
        const uint32_t order = lp.order, shift = lp.shift;
        uint32_t bad = 0;
        for (uint32_t i = order + tid; i < n; i += RES_THREADS) {
            long long sum = 0;
            for (uint32_t j = 0; j < order; j++) sum = mad_wide_s32(xs[i - 1 - j], lp.q[j], sum);   // :3187-3192
            const int32_t pred = (int32_t)(uint32_t)(unsigned long long)(sum >> shift);             // `as i32`
            const long long rr = (long long)xs[i] - (long long)pred;
            if (rr < INT32_MIN || rr > INT32_MAX) bad = 1;                                         // checked_sub -> ResidualOverflow
            rs[i - order] = (int32_t)rr;
        }
        bad = block_or_u32(bad, sm.red32);
        if (bad) lpc_ok = false;
        else {
            rice_search(rs, n - order, order, n, cfg, sm, sm.lpc);
            if (sm.lpc.fail) lpc_ok = false;
            lpc_bits = hdr_bits + order * bps + 4 + 5 + order * lp.precision + sm.lpc.resid_bits;
        }
    }
    // ---- choose (:2929-2979): fixed wins ties; VERBATIM unless strictly smaller ----
    const uint32_t verbatim_len = n * bps;   // u32 like the reference
    int pick = -1;                            // 0 fixed, 1 lpc
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
}

// ------------------------------------------------------------------------------------------------
// k_decide: channel assignment + frame header  (src/encode.rs:2282-2406, :2747-2786; src/stream.rs:242-276)
// ------------------------------------------------------------------------------------------------
__global__ void k_decide(EncCfg cfg, const FrameDesc* __restrict__ descs, const CandRec* __restrict__ cands,
                         const unsigned long long* __restrict__ abssum, FrameRec* __restrict__ frecs)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= cfg.nframes) return;
    const FrameDesc d = descs[f];
    const CandRec* c = cands + (size_t)f * cfg.nslots;
    FrameRec fr;
    decide_frame(cfg, d, c, abssum + (size_t)f * 4, fr);
    frecs[f] = fr;
}

// single CTA of 1024 threads: out_off = totals[0] + exclusive scan of frame_bytes; then totals[0] += sum,
// totals[1] = first byte of this group, totals[2] = one past its last byte (read by k_zero)
__global__ void __launch_bounds__(1024) k_scan(uint32_t nframes, FrameRec* __restrict__ frecs, uint32_t* __restrict__ frame_bytes_out,
                                               unsigned long long* __restrict__ totals, unsigned long long* __restrict__ mapped_total)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long tile_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long carry = totals[0];
    const unsigned long long start = carry;
    // the sizes (and error flags, in bit 31) of up to 32 tiles are fetched up front with independent loads: the scan loop below
    // then runs on registers instead of paying a memory round trip per tile of 1024 frames (59 -> 20 us for a 32768-frame group)
    constexpr uint32_t PRE = 32;
    uint32_t pre[PRE];
#pragma unroll
    for (uint32_t k = 0; k < PRE; k++) {
        const uint32_t f = k * 1024 + tid;
        pre[k] = f < nframes ? (frecs[f].frame_bytes | (frecs[f].err ? 0x80000000u : 0u)) : 0u;
    }
#pragma unroll
    for (uint32_t k = 0; k < PRE; k++) {
        const uint32_t base = k * 1024;
        if (base >= nframes) break;
        const uint32_t f = base + tid;
        const unsigned long long v = pre[k] & 0x7FFFFFFFu;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = wsum[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o) wi += t;
            }
            wsum[lane] = wi - w;   // exclusive prefix over warps
            if (lane == 31) tile_total = wi;
        }
        __syncthreads();
        if (f < nframes) {
            frecs[f].out_off = carry + wsum[wid] + (incl - v);
            if (frame_bytes_out) frame_bytes_out[f] = (uint32_t)v;
            if (pre[k] >> 31) totals[3] = 1;   // sticky error word read back by the host
        }
        carry += tile_total;
        __syncthreads();
    }
    for (uint32_t base = PRE * 1024; base < nframes; base += 1024) {   // (launch groups beyond 32768 frames: not used today)
        const uint32_t f = base + tid;
        const unsigned long long v = f < nframes ? frecs[f].frame_bytes : 0;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = wsum[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o) wi += t;
            }
            wsum[lane] = wi - w;   // exclusive prefix over warps
            if (lane == 31) tile_total = wi;
        }
        __syncthreads();
        if (f < nframes) {
            frecs[f].out_off = carry + wsum[wid] + (incl - v);
            if (frame_bytes_out) frame_bytes_out[f] = (uint32_t)v;
            if (frecs[f].err) totals[3] = 1;   // sticky error word read back by the host
        }
        carry += tile_total;
        __syncthreads();
    }
    if (tid == 0) {
        totals[0] = carry;
        totals[1] = start;
        totals[2] = carry;
        if (mapped_total) *mapped_total = carry;   // host-mapped: the cumulative size reaches the host without a copy
    }
}

// clears out[totals[1] .. totals[2]) -- the packers OR their bits into the output
__global__ void __launch_bounds__(256) k_zero(uint8_t* __restrict__ out, const unsigned long long* __restrict__ totals)
{
    const unsigned long long start = totals[1], end = totals[2];
    const unsigned long long a = (start + 15) & ~15ull, b = end & ~15ull;
    const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long gsize = (unsigned long long)gridDim.x * blockDim.x;
    if (a >= b) {
        for (unsigned long long i = start + gtid; i < end; i += gsize) out[i] = 0;
        return;
    }
    if (blockIdx.x == 0) {
        for (unsigned long long i = start + threadIdx.x; i < a; i += blockDim.x) out[i] = 0;
        for (unsigned long long i = b + threadIdx.x; i < end; i += blockDim.x) out[i] = 0;
    }
    uint4* body = reinterpret_cast<uint4*>(out + a);
    const unsigned long long nvec = (b - a) >> 4;
    for (unsigned long long i = gtid; i < nvec; i += gsize) body[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// k_pack: Partition::to_writer / encode_*_subframe bit emission (src/encode.rs:2982-3136, :3834-3863)
// ------------------------------------------------------------------------------------------------
constexpr int PACK_THREADS = 256;
constexpr int PACK_EPT = 4;   // residuals per thread per tile

// grid (F * nsub_max), block PACK_THREADS; dynamic smem (SMEM): bpad int32 samples + cap_words bit words
template <bool SMEM>
__global__ void __launch_bounds__(PACK_THREADS) k_pack(EncCfg cfg, uint32_t nsub_max, uint32_t cap_words, const FrameDesc* __restrict__ descs,
                                                      const int32_t* __restrict__ planes, const CandRec* __restrict__ cands,
                                                      const FrameRec* __restrict__ frecs, uint8_t* __restrict__ out)
{
    extern __shared__ __align__(16) int32_t dyn[];
    __shared__ uint32_t warp_tot[PACK_THREADS / 32];
    __shared__ uint32_t tile_tot;
    __shared__ CandRec cr;
    const uint32_t f = blockIdx.x / nsub_max, c = blockIdx.x % nsub_max, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const FrameRec& fr = frecs[f];
    if (c >= fr.nsub) return;
    const uint32_t slot = fr.slot[c];
    const uint32_t cand = f * cfg.nslots + slot;
    if (tid < sizeof(CandRec) / 4) reinterpret_cast<uint32_t*>(&cr)[tid] = reinterpret_cast<const uint32_t*>(cands + cand)[tid];
    const uint32_t n = descs[f].n;
    const int32_t* plane = planes + (size_t)cand * cfg.bpad;
    int32_t* xs = dyn;
    uint32_t* wbuf = reinterpret_cast<uint32_t*>(dyn + cfg.bpad);
    // global bit range of this CTA: subframe c (plus the frame header for c == 0)
    const unsigned long long frame_bit0 = fr.out_off * 8ull;
    const unsigned long long g0 = frame_bit0 + (c == 0 ? 0u : fr.sub_bit[c]);
    __syncthreads();
    const unsigned long long g1 = frame_bit0 + fr.sub_bit[c] + cr.bits;
    const unsigned long long w0 = g0 >> 5, w1 = (g1 + 31) >> 5;
    const uint32_t nwords = (uint32_t)(w1 - w0);
    uint32_t* gwords = reinterpret_cast<uint32_t*>(out);
    uint32_t* words;          // where put_bits writes
    unsigned long long origin;   // bit position of words[0] in the same coordinates as `pos` below
    if (SMEM) {
        for (uint32_t i = tid; i < nwords + 1 && i < cap_words; i += PACK_THREADS) wbuf[i] = 0;
        words = wbuf;
        origin = w0 << 5;
    } else {
        words = gwords;
        origin = 0;
    }
    const uint32_t wasted = cr.wasted, bps = cr.bps, type = cr.type, order = (type >= 2) ? cr.order : 0;
    if (SMEM) {
        for (uint32_t i = tid; i < n; i += PACK_THREADS) xs[i] = plane[i] >> wasted;
    }
    __syncthreads();
    auto X = [&](uint32_t i) -> int32_t { return SMEM ? xs[i] : (plane[i] >> wasted); };
    auto put = [&](unsigned long long gpos, uint32_t nbits, uint32_t v) { put_bits<!SMEM>(words, gpos - origin, nbits, v); };

    unsigned long long pos = frame_bit0 + fr.sub_bit[c];
    if (tid == 0) {
        if (c == 0)
            for (uint32_t i = 0; i < fr.hdr_len; i++) put(frame_bit0 + 8ull * i, 8, fr.hdr[i]);
        // SubframeHeader (src/stream.rs:1397-1413): pad, 6-bit type, wasted flag, unary(wasted - 1)
        const uint32_t code = type == 0 ? 0u : type == 1 ? 1u : type == 2 ? 8u + order : 31u + order;
        put(pos, 8, (code << 1) | (wasted ? 1u : 0u));
        if (wasted) put(pos + 8 + (wasted - 1), 1, 1);
    }
    pos += 8 + wasted;
    if (type == 0) {   // CONSTANT: the sample is zero by construction (:2870-2887)
        if (tid == 0) put(pos, bps, (uint32_t)X(0));
    } else if (type == 1) {   // VERBATIM (:3000-3018)
        for (uint32_t i = tid; i < n; i += PACK_THREADS) put(pos + (unsigned long long)i * bps, bps, (uint32_t)X(i));
    } else {
        if (tid < order) put(pos + (unsigned long long)tid * bps, bps, (uint32_t)X(tid));   // warm-up (:3083, :3118)
        pos += (unsigned long long)order * bps;
        if (type == 3) {
            const uint32_t prec = cr.precision;
            if (tid == 0) {
                put(pos, 4, prec - 1);       // :3122
                put(pos + 4, 5, cr.shift);   // :3129
            }
            if (tid < order) put(pos + 9 + (unsigned long long)tid * prec, prec, (uint32_t)(int32_t)cr.q[tid]);   // :3131
            pos += 9 + (unsigned long long)order * prec;
        }
        // residual block (:3944-3961)
        if (tid == 0) {
            put(pos, 2, cr.method);
            put(pos + 2, 4, cr.porder_w);
        }
        pos += 6;
        const uint32_t L = n - order;
        const uint32_t cp = n >> cr.porder_g;
        const uint32_t j0 = (1u << cr.porder_g) - cr.nparts;
        const bool cp_pow2 = (cp & (cp - 1)) == 0;
        const uint32_t cp_shift = 31u - (uint32_t)__clz((int)cp);
        const uint32_t hb = cr.method ? 5u : 4u;
        const uint32_t escape_code = cr.method ? 31u : 15u;
        const uint32_t shift = cr.shift;
        uint32_t carry = 0;
        for (uint32_t base = 0; base < L; base += PACK_THREADS * PACK_EPT) {
            int32_t r[PACK_EPT];
            uint32_t len[PACK_EPT], code[PACK_EPT], first[PACK_EPT];
            uint32_t tsum = 0;
#pragma unroll
            for (int e = 0; e < PACK_EPT; e++) {
                const uint32_t i = base + tid * PACK_EPT + e;
                len[e] = 0; code[e] = 0; first[e] = 0; r[e] = 0;
                if (i < L) {
                    const uint32_t ia = i + order;
                    long long d;
                    if (type == 2) {
                        const long long a0 = X(ia);
                        switch (order) {
                        case 0: d = a0; break;
                        case 1: d = a0 - X(ia - 1); break;
                        case 2: d = a0 - 2ll * X(ia - 1) + X(ia - 2); break;
                        case 3: d = a0 - 3ll * X(ia - 1) + 3ll * X(ia - 2) - X(ia - 3); break;
                        default: d = a0 - 4ll * X(ia - 1) + 6ll * X(ia - 2) - 4ll * X(ia - 3) + X(ia - 4); break;
                        }
                    } else {
                        long long sum = 0;
                        for (uint32_t j = 0; j < order; j++) sum = mad_wide_s32(X(ia - 1 - j), cr.q[j], sum);
                        d = (long long)X(ia) - (long long)(int32_t)(uint32_t)(unsigned long long)(sum >> shift);
                    }
                    r[e] = (int32_t)d;
                    const uint32_t pj = (cp_pow2 ? (ia >> cp_shift) : (ia / cp));
                    const uint32_t j = pj - j0;
                    const uint32_t cc = cr.rice[j];
                    code[e] = cc;
                    const uint32_t pstart = pj * cp;   // absolute index of the partition's first sample
                    first[e] = (ia == (pstart > order ? pstart : order)) ? 1u : 0u;
                    uint32_t l = 0;
                    if (cc < 0x40) l = (zigzag32(r[e]) >> cc) + 1u + cc;
                    else if (cc & 0x40) l = cc & 31u;
                    if (first[e]) l += (cc < 0x40) ? hb : hb + 5;
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
            if (wid == 0) {
                const uint32_t w = lane < PACK_THREADS / 32 ? warp_tot[lane] : 0;
                uint32_t wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= (uint32_t)o) wi += t;
                }
                if (lane < PACK_THREADS / 32) warp_tot[lane] = wi - w;
                if (lane == 31) tile_tot = wi;
            }
            __syncthreads();
            unsigned long long p = pos + carry + warp_tot[wid] + (incl - tsum);
#pragma unroll
            for (int e = 0; e < PACK_EPT; e++) {
                const uint32_t i = base + tid * PACK_EPT + e;
                if (i < L) {
                    const uint32_t cc = code[e];
                    unsigned long long q = p;
                    if (first[e]) {   // ResidualPartitionHeader::to_writer (src/stream.rs:1603-1619)
                        if (cc < 0x40) { put(q, hb, cc); q += hb; }
                        else { put(q, hb, escape_code); put(q + hb, 5, (cc & 0x40) ? (cc & 31u) : 0u); q += hb + 5; }
                    }
                    if (cc < 0x40) {
                        const uint32_t u = zigzag32(r[e]);
                        const uint32_t msb = u >> cc;
                        put(q + msb, cc + 1, (1u << cc) | (u & ((1u << cc) - 1u)));   // unary stop bit + cc LSBs (:3850-3851)
                    } else if (cc & 0x40) {
                        put(q, cc & 31u, (uint32_t)r[e]);   // escaped: raw two's complement (:3857)
                    }
                    p += len[e];
                }
            }
            carry += tile_tot;
            __syncthreads();
        }
    }
    if (SMEM) {
        __syncthreads();
        // interior words are owned by this CTA: plain stores; the two boundary words are shared with neighbours
        for (uint32_t i = tid; i < nwords; i += PACK_THREADS) {
            const uint32_t v = __byte_perm(wbuf[i], 0, 0x0123);
            if (i == 0 || i == nwords - 1) { if (v) atomicOr(gwords + w0 + i, v); }
            else gwords[w0 + i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_crc16: CRC-16 over the frame (src/crc.rs:144-188, src/encode.rs:2408-2409)
// ------------------------------------------------------------------------------------------------
constexpr int CRC_THREADS = 128;

__global__ void __launch_bounds__(CRC_THREADS) k_crc16(const FrameRec* __restrict__ frecs, uint8_t* __restrict__ out)
{
    __shared__ uint16_t table[256];
    __shared__ uint32_t pw[24];   // x^(8 * 2^j) mod P
    __shared__ uint32_t wred[CRC_THREADS / 32];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 256; i += CRC_THREADS) table[i] = crc16_table_entry(i);
    if (tid == 0) {
        uint32_t v = 0x0100;   // x^8
        for (int j = 0; j < 24; j++) { pw[j] = v; v = gf16_mulmod(v, v); }
    }
    __syncthreads();
    const FrameRec& fr = frecs[blockIdx.x];
    const uint32_t total = fr.frame_bytes - 2;
    const uint8_t* p = out + fr.out_off;
    const uint32_t per = (total + CRC_THREADS - 1) / CRC_THREADS;
    const uint32_t a = min(tid * per, total), b = min(a + per, total);
    uint32_t crc = 0;
    for (uint32_t i = a; i < b; i++) crc = (table[((crc >> 8) ^ p[i]) & 0xff] ^ (crc << 8)) & 0xffffu;
    // shift by the bytes that follow: crc * x^(8 * (total - b)) mod P
    uint32_t rem = total - b;
    for (int j = 0; rem; j++, rem >>= 1)
        if (rem & 1u) crc = gf16_mulmod(crc, pw[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
    if ((tid & 31) == 0) wred[tid >> 5] = crc;
    __syncthreads();
    if (tid == 0) {
        uint32_t c = 0;
        for (int w = 0; w < CRC_THREADS / 32; w++) c ^= wred[w];
        out[fr.out_off + total] = (uint8_t)(c >> 8);
        out[fr.out_off + total + 1] = (uint8_t)c;
    }
}

uint32_t pack_cap_words(const EncCfg& cfg);
#include "encode_fast.inl"

// ------------------------------------------------------------------------------------------------
// launch wrappers (called from engine.cu)
// ------------------------------------------------------------------------------------------------
void launch_planes(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, int32_t* planes, uint32_t* ormask,
                   unsigned long long* abssum, cudaStream_t st)
{
    dim3 grid((cfg.block_size + 255) / 256, cfg.nframes);
    count_launch(), k_planes<<<grid, 256, 0, st>>>(cfg, descs, pcm, planes, ormask, abssum);
}

void launch_lpc(const EncCfg& cfg, const FrameDesc* descs, const int32_t* planes, const uint32_t* ormask,
                const unsigned long long* abssum, const double* winpool, LpcRec* lpcs, cudaStream_t st)
{
    const uint32_t ncand = cfg.nframes * cfg.nslots;
    count_launch(), k_lpc<<<(ncand + LPC_WARPS - 1) / LPC_WARPS, 32 * LPC_WARPS, 0, st>>>(cfg, descs, planes, ormask, abssum, winpool, lpcs, ncand);
}

bool residual_uses_smem(const EncCfg& cfg) { return (size_t)cfg.bpad * 8 <= 96 * 1024; }

cudaError_t launch_residual(const EncCfg& cfg, const FrameDesc* descs, const int32_t* planes, const uint32_t* ormask,
                            const unsigned long long* abssum, const LpcRec* lpcs, CandRec* cands, int32_t* scratch, cudaStream_t st)
{
    const uint32_t ncand = cfg.nframes * cfg.nslots;
    if (residual_uses_smem(cfg)) {
        const size_t smem = (size_t)cfg.bpad * 8;
        {
            cudaError_t e = cudaFuncSetAttribute(k_residual<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
        }
        count_launch(), k_residual<true><<<ncand, RES_THREADS, smem, st>>>(cfg, descs, planes, ormask, abssum, lpcs, cands, nullptr);
    } else {
        count_launch(), k_residual<false><<<ncand, RES_THREADS, 0, st>>>(cfg, descs, planes, ormask, abssum, lpcs, cands, scratch);
    }
    return cudaGetLastError();
}

void launch_decide_scan(const EncCfg& cfg, const FrameDesc* descs, const CandRec* cands, const unsigned long long* abssum,
                        FrameRec* frecs, uint32_t* frame_bytes_out, unsigned long long* totals, unsigned long long* mapped_total, uint8_t* out,
                        bool zero_output, cudaStream_t st)
{
    count_launch(), k_decide<<<(cfg.nframes + 127) / 128, 128, 0, st>>>(cfg, descs, cands, abssum, frecs);
    count_launch(), k_scan<<<1, 1024, 0, st>>>(cfg.nframes, frecs, frame_bytes_out, totals, mapped_total);
    if (zero_output) count_launch(), k_zero<<<148 * 4, 256, 0, st>>>(out, totals);   // the OR-ing packers need it; k_pack3 writes whole frames
}

uint32_t pack_cap_words(const EncCfg& cfg) { return (uint32_t)(((size_t)cfg.bpad * (cfg.bps + 1) + 512) / 32 + 8); }
bool pack_uses_smem(const EncCfg& cfg) { return (size_t)cfg.bpad * 4 + (size_t)pack_cap_words(cfg) * 4 <= 96 * 1024; }

cudaError_t launch_pack_crc(const EncCfg& cfg, const FrameDesc* descs, const int32_t* planes, const CandRec* cands,
                            const FrameRec* frecs, uint8_t* out, cudaStream_t st)
{
    const uint32_t nsub_max = cfg.mode == MODE_INDEPENDENT ? cfg.channels : 2;
    const uint32_t cap_words = pack_cap_words(cfg);
    if (pack_uses_smem(cfg)) {
        const size_t smem = (size_t)cfg.bpad * 4 + (size_t)cap_words * 4;
        {
            cudaError_t e = cudaFuncSetAttribute(k_pack<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
        }
        count_launch(), k_pack<true><<<cfg.nframes * nsub_max, PACK_THREADS, smem, st>>>(cfg, nsub_max, cap_words, descs, planes, cands, frecs, out);
    } else {
        count_launch(), k_pack<false><<<cfg.nframes * nsub_max, PACK_THREADS, 0, st>>>(cfg, nsub_max, cap_words, descs, planes, cands, frecs, out);
    }
    count_launch(), k_crc16<<<cfg.nframes, CRC_THREADS, 0, st>>>(frecs, out);
    return cudaGetLastError();
}

}   // namespace flacb200
