// encode_lpc.cu -- k_lpc3: LpcParameters::best (src/encode.rs:3292-3332) for stereo frames straight from the packed
// PCM, with the global loads taken out of the FP64 loop by an asynchronous staging pipeline.
//
// Same arithmetic as k_lpc2 (encode_fast.inl): the reference's autocorrelation is a strict left-to-right f64 sum per lag
// (:3491-3497), so each lag is one sequential chain of separately rounded DMUL + DADD; a lane owns FOUR consecutive lags
// of one candidate, four lanes make a candidate, a warp runs the 8 candidates (L, R, M, S of two frames) from per-candidate
// rings of windowed samples in shared memory (four 32-sample tiles + mirrors of two).
//
// What is new: a tile's raw PCM bytes (32 samples x 2 channels) and its window values travel global -> shared with
// cp.async (16-byte chunks, zero-filled past the block end) into a two-slot staging area per warp, two tiles ahead of the
// tile being computed; the lanes never wait on a global load inside the FP64 loop (k_lpc2 spent a third of its stall
// cycles there), and no registers are held for prefetched samples.
#include "common.cuh"
#include "glibc_log.cuh"
#include "tiles.cuh"

namespace flacb200 {

constexpr int L3_WARPS = 5;       // 5 warps x 14.4 KB: three CTAs per SM
constexpr int L3_RING = 192;      // 4 tiles + mirrors of tiles 0 and 1
constexpr int L3_CD = L3_RING + 2;   // doubles per candidate (+ 2: bank skew between candidates that keeps 16-byte alignment)
constexpr int L3_CANDS = 8;       // candidates per warp = two stereo frames
constexpr int L3_STAGE_BYTES = 2 * 2 * 512;   // 2 slots x 2 frames x (256 B PCM + 256 B window)

__device__ inline void l3_cp16(uint32_t dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ inline void l3_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ inline void l3_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// dynamic smem: L3_WARPS * (L3_CANDS * L3_CD * 8 + L3_STAGE_BYTES) bytes
__global__ void __launch_bounds__(32 * L3_WARPS, 3) k_lpc3(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
                                                         const double* __restrict__ winpool, LpcRec* __restrict__ out, uint32_t nframes)
{
    extern __shared__ __align__(16) uint8_t l3_dyn[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t M = cfg.max_lpc_order;
    // a warp takes pairs of frames; the grid may be smaller than the work (persistent launch: a few CTAs per SM that run
    // beside the integer kernels of the previous launch group, see flacb200_encode)
    for (uint32_t pair = blockIdx.x * L3_WARPS + wid; 2 * pair < nframes; pair += gridDim.x * L3_WARPS) {
    const uint32_t f0 = pair * 2;   // first frame of this warp
    const uint32_t nfr = min(2u, nframes - f0);
    uint8_t* wsm = l3_dyn + (size_t)wid * (L3_CANDS * L3_CD * 8 + L3_STAGE_BYTES);
    double* wbase = reinterpret_cast<double*>(wsm);
    uint8_t* stage = wsm + L3_CANDS * L3_CD * 8;
    const uint32_t stage_sa = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t g = lane >> 2, m = lane & 3;   // candidate, lag group
    const bool live = g < nfr * 4;
    const uint32_t B = cfg.bytes_per_sample, TB = 64 * B;   // bytes of one 32-sample stereo tile
    const bool big = cfg.pcm_kind == 1;
    // the warp's two frames
    const uint8_t* fp[2];
    const double* wp[2];
    uint32_t fn[2];
#pragma unroll
    for (int f = 0; f < 2; f++) {
        const FrameDesc d = descs[f0 + ((uint32_t)f < nfr ? f : 0)];
        fp[f] = pcm + d.pcm_off * (unsigned long long)(2 * B);
        wp[f] = winpool + d.win_off;
        fn[f] = ((uint32_t)f < nfr && d.n > M) ? d.n : 0u;   // n <= M: InsufficientLpcSamples (:3300), nothing to analyse
    }
    if (lane < nfr * 4) out[(size_t)f0 * 4 + lane].ok = 0;
    const uint32_t nmax = max(fn[0], fn[1]);
    const uint32_t ntiles = (nmax + 31) / 32;
    const uint32_t per_frame = 4 * B + 16;   // 16-byte chunks per frame and tile: PCM, then window

    // global -> staging slot (tile & 1); bytes past the end of the block are zero-filled.  A lane copies the same (at most
    // two) 16-byte chunks of every tile: chunk c of the 2 * per_frame chunks is PCM (k < 4 B) or window of frame c / per_frame.
    const uint8_t* csrc[2];
    uint32_t cdst[2], cstride[2], ctotal[2], coff[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t c = lane + 32 * j;
        const bool on = c < 2 * per_frame;
        const uint32_t f = c >= per_frame ? 1u : 0u, k = c - f * per_frame;
        const uint32_t nf = f ? fn[1] : fn[0];
        if (k < 4 * B) {
            csrc[j] = f ? fp[1] : fp[0];
            cdst[j] = f * 512 + k * 16; cstride[j] = TB; coff[j] = k * 16; ctotal[j] = on ? nf * 2 * B : 0u;
        } else {
            const uint32_t kk = k - 4 * B;
            csrc[j] = reinterpret_cast<const uint8_t*>(f ? wp[1] : wp[0]);
            cdst[j] = f * 512 + 256 + kk * 16; cstride[j] = 256; coff[j] = kk * 16; ctotal[j] = on ? nf * 8 : 0u;
        }
    }
    const bool two = lane + 32 < 2 * per_frame;
    auto issue = [&](uint32_t tile) {
        const uint32_t slot_sa = stage_sa + (tile & 1u) * 1024u;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (j == 1 && !two) break;
            const uint32_t off = tile * cstride[j] + coff[j];
            const uint32_t bytes = off < ctotal[j] ? min(16u, ctotal[j] - off) : 0u;
            l3_cp16(slot_sa + cdst[j], csrc[j] + (bytes ? off : 0u), bytes);
        }
        l3_commit();
    };

    uint32_t shift_mask = 0;   // bit c: candidate c must be redone with its wasted bits shifted out
    uint32_t wasted_of[L3_CANDS], masks[L3_CANDS];
#pragma unroll
    for (int c = 0; c < L3_CANDS; c++) { wasted_of[c] = 0; masks[c] = 0; }
    double acc0 = -0.0, acc1 = -0.0, acc2 = -0.0, acc3 = -0.0;   // Iterator::sum::<f64>() folds from -0.0
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && shift_mask == 0) break;
        // staging slot -> the 8 rings: Frame::fill_from_buf (src/audio.rs:149-187), decorrelation (:2721, :2734), Window::apply (:1799)
        auto store = [&](uint32_t tile) {
            const uint8_t* slot = stage + (tile & 1u) * 1024u;
            const uint32_t pos = (tile & 3) * 32 + lane;
            const uint32_t sh = 32 - 8 * B;
            const bool mirror = (tile & 3) < 2;   // mirrors of tiles 0 and 1 (mod 4)
#pragma unroll
            for (int f = 0; f < 2; f++) {
                const uint16_t* p16 = reinterpret_cast<const uint16_t*>(slot + f * 512 + lane * 2 * B);
                uint32_t raw[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((uint32_t)k < B) raw[k] = p16[k];
                const double wv = *reinterpret_cast<const double*>(slot + f * 512 + 256 + lane * 8);
                const unsigned long long wide = (unsigned long long)(raw[0] | (raw[1] << 16)) | ((unsigned long long)(raw[2] | (raw[3] << 16)) << 32);
                const uint32_t lw = (uint32_t)wide, rw = (uint32_t)(wide >> (8 * B));   // low B bytes: the sample in memory order
                int32_t l, r;
                if (big) {
                    l = (int32_t)__byte_perm(lw, 0, 0x0123) >> sh;
                    r = (int32_t)__byte_perm(rw, 0, 0x0123) >> sh;
                } else {
                    l = (int32_t)(lw << sh) >> sh;
                    r = (int32_t)(rw << sh) >> sh;
                }
                const int32_t px[4] = {l, r, (l + r) >> 1, l - r};
                if (pass == 0) {   // common pass: no wasted bits assumed, OR masks gathered on the fly
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = f * 4 + k;
                        masks[c] |= (uint32_t)px[k];
                        const double v = __dmul_rn((double)px[k], wv);
                        double* ring = wbase + (size_t)c * L3_CD;
                        ring[pos] = v;
                        if (mirror) ring[128 + pos] = v;
                    }
                } else {           // rare pass: only the candidates that have wasted bits, shifted (:2891)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = f * 4 + k;
                        if (!((shift_mask >> c) & 1u)) continue;
                        const double v = __dmul_rn((double)(px[k] >> wasted_of[c]), wv);
                        double* ring = wbase + (size_t)c * L3_CD;
                        ring[pos] = v;
                        if (mirror) ring[128 + pos] = v;
                    }
                }
            }
        };
        // prologue: tiles 0..2 into the rings, tile 3 in flight
        issue(0);
        issue(1);
        l3_wait<0>();
        __syncwarp();
        store(0);
        store(1);
        __syncwarp();
        issue(2);
        issue(3);
        l3_wait<1>();
        __syncwarp();
        store(2);
        __syncwarp();
        acc0 = acc1 = acc2 = acc3 = -0.0;
        const double* ring = wbase + (size_t)g * L3_CD;
        for (uint32_t t = 0; t < ntiles; t++) {
            issue(t + 4);   // lands two tiles of FP64 work later
            const double* pa = ring + (t & 3) * 32;
            const double* pb = pa + 4 * m;
            const double2 b01 = *reinterpret_cast<const double2*>(pb), b23 = *reinterpret_cast<const double2*>(pb + 2);
            double b0 = b01.x, b1 = b01.y, b2 = b23.x, b3 = b23.y;
#pragma unroll
            for (int s = 0; s < 32; s += 2) {   // autocorrelate :3491-3497, four lags per lane, two samples per 128-bit load
                const double2 a = *reinterpret_cast<const double2*>(pa + s);
                const double2 bn = *reinterpret_cast<const double2*>(pb + s + 4);
                acc0 = __dadd_rn(acc0, __dmul_rn(a.x, b0));
                acc1 = __dadd_rn(acc1, __dmul_rn(a.x, b1));
                acc2 = __dadd_rn(acc2, __dmul_rn(a.x, b2));
                acc3 = __dadd_rn(acc3, __dmul_rn(a.x, b3));
                acc0 = __dadd_rn(acc0, __dmul_rn(a.y, b1));
                acc1 = __dadd_rn(acc1, __dmul_rn(a.y, b2));
                acc2 = __dadd_rn(acc2, __dmul_rn(a.y, b3));
                acc3 = __dadd_rn(acc3, __dmul_rn(a.y, bn.x));
                b0 = b2; b1 = b3; b2 = bn.x; b3 = bn.y;
            }
            l3_wait<1>();   // tile t + 3 has landed (only tile t + 4 may still be in flight)
            __syncwarp();
            store(t + 3);   // replaces tile t - 1; tiles t + 1 and t + 2 stay resident
            __syncwarp();
        }
        l3_wait<0>();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < L3_CANDS; c++) masks[c] = __reduce_or_sync(0xffffffffu, masks[c]);
        // the rings are dead now: R[] goes on top of them
        const bool mine_redo = (shift_mask >> g) & 1u;
        if (live && (pass == 0 || mine_redo)) {
            double* Rg = wbase + (size_t)g * L3_CD;
            if (4 * m + 0 <= M) Rg[4 * m + 0] = acc0;
            if (4 * m + 1 <= M) Rg[4 * m + 1] = acc1;
            if (4 * m + 2 <= M) Rg[4 * m + 2] = acc2;
            if (4 * m + 3 <= M) Rg[4 * m + 3] = acc3;
        }
        if (pass == 0) {
#pragma unroll
            for (int c = 0; c < L3_CANDS; c++) {
                const uint32_t mk = masks[c];
                wasted_of[c] = (mk == 0 || (mk & 1u)) ? 0u : (uint32_t)__ffs((int)mk) - 1u;   // :2878-2898
                if (wasted_of[c]) shift_mask |= 1u << c;
            }
        }
        __syncwarp();
    }
    // a second pass rebuilds only the rings of the shifted candidates (store skips the others), so the R[] that the
    // untouched candidates parked on top of their own rings is still intact here
    __syncwarp();
    // ---- per candidate: Levinson-Durbin on the group's first lane, order estimate on all its lanes ----
    double* base = wbase + (size_t)g * L3_CD;
    double* Rv = base;              // M + 1
    double* errv = Rv + (M + 1);    // M
    double* bitsv = errv + M;       // M
    double* sets = bitsv + M;       // M (M + 1) / 2: coefficient set of every order, triangular
    uint32_t mask = 0, wasted = 0;
#pragma unroll
    for (int c = 0; c < L3_CANDS; c++)
        if ((uint32_t)c == g) { mask = masks[c]; wasted = wasted_of[c]; }
    const uint32_t cand = f0 * 4 + g;
    const uint32_t n = (g >> 2) ? fn[1] : fn[0];
    const bool run = live && mask != 0 && n > M;   // all-zero candidates become CONSTANT (:2883)
    const uint32_t bps = cand_bps(cfg, g & 3) - wasted;
    const uint32_t precision = lpc_precision_for(n);
    if (run && m == 0) {   // lp_coefficients (:3536-3580), every order's set kept
        const double* R = Rv;
        double* a = sets;   // order 1
        double k = __ddiv_rn(R[1], R[0]);
        a[0] = k;
        errv[0] = __dmul_rn(R[0], __dsub_rn(1.0, __dmul_rn(k, k)));
        for (uint32_t i = 1; i < M; i++) {
            double* b = a + i;   // the set of order i + 1 follows the i entries of order i
            double s = -0.0;
            for (uint32_t j = 0; j < i; j++) s = __dadd_rn(s, __dmul_rn(R[i - j], a[j]));
            const double q = __dsub_rn(R[i + 1], s);
            k = __ddiv_rn(q, errv[i - 1]);
            for (uint32_t j = 0; j < i; j++) b[j] = __dsub_rn(a[j], __dmul_rn(k, a[i - 1 - j]));
            b[i] = k;
            errv[i] = __dmul_rn(errv[i - 1], __dsub_rn(1.0, __dmul_rn(k, k)));
            a = b;
        }
    }
    __syncwarp();
    if (run) {   // subframe_bits_by_order (:3656-3686): this lane's orders 4m + 1 .. 4m + 4
        const double error_scale = __ddiv_rn(0.5, (double)n);
        const double divisor = 2.0 * 0.693147180559945309417232121458176568;
        for (uint32_t o = 4 * m + 1; o <= min(4 * m + 4, M); o++) {
            const double bpr = __ddiv_rn(glibc_log(__dmul_rn(errv[o - 1], error_scale)), divisor);
            bitsv[o - 1] = fma(bpr, (double)(n - o), (double)(o * (bps + precision)));
        }
    }
    __syncwarp();
    if (run && m == 0) {
        int best = 0;   // compute_best_order (:3688-3702): take_while(err > 0), first minimum under total_cmp
        double best_bits = 0.0;
        for (uint32_t o = 1; o <= M; o++) {
            if (!(errv[o - 1] > 0.0)) break;
            const double b = bitsv[o - 1];
            if (best == 0 || total_key(b) < total_key(best_bits)) { best = (int)o; best_bits = b; }
        }
        if (best != 0) {
            const double* cur = sets + (size_t)best * (best - 1) / 2;
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
    __syncwarp();   // the next pair's rings overwrite R[] and the coefficient sets
    }
}

// stereo frames (L, R, M, S slots), packed byte PCM, 16-byte aligned blocks, order <= 15 (four lanes x four lags)
bool lpc3_ok(const EncCfg& cfg, bool blocks_aligned16)
{
    return cfg.mode != MODE_INDEPENDENT && cfg.nslots == 4 && cfg.pcm_kind <= 1 && cfg.max_lpc_order >= 1 && cfg.max_lpc_order <= 15 &&
           blocks_aligned16;
}

// max_ctas: 0 = one warp per pair of frames; else a persistent grid of at most that many CTAs
cudaError_t launch_lpc3(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const double* winpool, LpcRec* lpcs, uint32_t max_ctas,
                        cudaStream_t st)
{
    const size_t smem = (size_t)L3_WARPS * (L3_CANDS * L3_CD * 8 + L3_STAGE_BYTES);
    cudaError_t e = cudaFuncSetAttribute(k_lpc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t nwarps = (cfg.nframes + 1) / 2;
    uint32_t grid = (nwarps + L3_WARPS - 1) / L3_WARPS;
    if (max_ctas && grid > max_ctas) grid = max_ctas;
    count_launch(), k_lpc3<<<grid, 32 * L3_WARPS, smem, st>>>(cfg, descs, pcm, winpool, lpcs, cfg.nframes);
    return cudaGetLastError();
}


// =====================================================================================================================
// k_lpc4: the same analysis with ONE LANE PER CANDIDATE -- a warp takes eight stereo frames (32 candidates).
//
// k_lpc3 spreads a candidate's lags over four lanes x four lags: 16 lag slots for the 13 lags of Options::best, a ring of
// windowed doubles in shared memory that every lane reads back, and an unpack/convert/store phase that costs as many issue
// slots as the FP64 loop itself (FP64 pipe 67 % busy).  Here a lane owns its candidate outright:
//   * all NL lags are NL accumulators of the lane (13 of 13 FP64 operations useful), the lagged samples are a circular
//     window of 16 doubles in registers with static indices (the tile loop is unrolled over 16 samples);
//   * the raw PCM bytes (16 samples x 2 channels) and the window values of the eight frames travel global -> shared with
//     cp.async, two tiles ahead.  Phase A of a tile is the warp's joint work: lane (q, f) takes samples 4 q .. 4 q + 3 of
//     frame f, extracts left and right once (one PRMT + shift each, static positions), forms all four channel
//     combinations, converts, multiplies by the window value and parks the 16 doubles in the value buffer
//     [sample][8 k + f]; phase B is the candidate lane's own: one 64-bit shared load, NL DMULs and NL DADDs per sample,
//     nothing else (an FP64 instruction holds the issue port for two cycles: the ~250 non-FP64 slots of a tile are what
//     there is to save beside its 432 x 2);
//   * the first tile and tiles that reach past the shortest frame of the warp run a predicated copy of phase B
//     (a lag only counts samples that exist: no zero terms are ever added, the sums start from -0.0 as in k_lpc3);
//   * Levinson-Durbin, the order estimate and the quantisation run on all 32 lanes at once, fully unrolled over the
//     template's order so that R[], the coefficient sets and the errors stay in registers; the set of the best order so
//     far is kept as the recursion proceeds (subframe_bits_by_order only needs the error of the order just finished).
// Arithmetic, operation for operation, is k_lpc3's (and the reference's :3478-3702).
// =====================================================================================================================
#ifndef FLACB200_L4_MINB
#define FLACB200_L4_MINB 4
#endif
constexpr int L4_SLOTS = 3;
constexpr int L4_STRIDE = 144;   // bytes per frame in a slot: <= 128 B of PCM (or 16 window doubles) + 16 B skew -> the 8 frames' LDS.128 hit distinct banks
constexpr int L4_SLOT_BYTES = 2 * 8 * L4_STRIDE;   // PCM of 8 frames, then their window values

// sample at byte offset O (static) of the words w[]: the B bytes moved to the top of a register by PRMT, then shifted down
template <int B, int O, int NW>
__device__ __forceinline__ int32_t l4_extract(const uint32_t (&w)[NW], const uint32_t (&sel)[4])
{
    constexpr int wi = O / 4, sh = O % 4;
    constexpr int wj = (wi + 1 < NW) ? wi + 1 : wi;   // the second word is only selected from when the sample straddles
    return (int32_t)__byte_perm(w[wi], w[wj], sel[sh]) >> (32 - 8 * B);
}

// phase A of a tile, the warp together: lane (q = lane / 8, f = lane % 8) takes samples 4 q .. 4 q + 3 of frame f -- it pulls their
// 8 B bytes with 64/128-bit loads, extracts left and right ONCE (PRMT + shift each, static positions), forms all four channel
// combinations, converts, multiplies by the window value and parks the 16 doubles in the columns of the frame's four candidates
// (value buffer [sample][8 k + f], rows of 34 doubles: the stores of a half-warp fall on 16 distinct bank pairs).
// SHIFT: the rare second pass, every candidate's wasted bits shifted out (:2891); the first pass gathers the OR masks instead.
constexpr int L4_VROW = 34;
template <int B, bool SHIFT>
__device__ __forceinline__ void l4_convert(const uint8_t* __restrict__ spcm, const uint8_t* __restrict__ swin, double* __restrict__ vout,
                                           const uint32_t (&sel)[4], const uint32_t (&wk)[4], uint32_t (&mk)[4])
{
    constexpr int NW = 2 * B;   // words of four stereo samples
    uint32_t w[NW];
    if (B == 2) {
        const uint4 a = *reinterpret_cast<const uint4*>(spcm);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
#pragma unroll
        for (int c = 0; c < NW / 2; c++) {
            const uint2 a = *reinterpret_cast<const uint2*>(spcm + 8 * c);
            w[2 * c] = a.x; w[2 * c + 1] = a.y;
        }
    }
    const double2 w01 = *reinterpret_cast<const double2*>(swin), w23 = *reinterpret_cast<const double2*>(swin + 16);
    const double wv[4] = {w01.x, w01.y, w23.x, w23.y};
    int32_t l[4], r[4];
    l[0] = l4_extract<B, 0 * B, NW>(w, sel); r[0] = l4_extract<B, 1 * B, NW>(w, sel);
    l[1] = l4_extract<B, 2 * B, NW>(w, sel); r[1] = l4_extract<B, 3 * B, NW>(w, sel);
    l[2] = l4_extract<B, 4 * B, NW>(w, sel); r[2] = l4_extract<B, 5 * B, NW>(w, sel);
    l[3] = l4_extract<B, 6 * B, NW>(w, sel); r[3] = l4_extract<B, 7 * B, NW>(w, sel);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int32_t px[4] = {l[i], r[i], (l[i] + r[i]) >> 1, l[i] - r[i]};   // L | R | mid (:2721) | side (:2734)
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const int32_t pv = SHIFT ? px[kk] >> wk[kk] : px[kk];
            if (!SHIFT) mk[kk] |= (uint32_t)pv;
            vout[i * L4_VROW + 8 * kk] = __dmul_rn((double)pv, wv[i]);   // Window::apply (:1799)
        }
    }
}

// phase B: autocorrelate (:3491-3497).  Sample s adds win[s - lag] * v_s to every lag's sum: NL independent DMULs, then NL
// DADDs -- nothing but FP64 work and one shared-memory load per sample.
template <int NL, bool EDGE>
__device__ __forceinline__ void l4_accumulate(const double* __restrict__ vcol, uint32_t idx0, uint32_t n, double (&acc)[NL], double (&win)[16])
{
    // software pipeline: the products of sample s + 1 are formed while the sums take in the products of sample s, so no DADD
    // ever waits on a DMUL that has just been issued
    double p[NL];
    {
        const double v = vcol[0];
        win[0] = v;
#pragma unroll
        for (int lag = 0; lag < NL; lag++) p[lag] = __dmul_rn(win[(0 - lag) & 15], v);
    }
#pragma unroll
    for (int s = 0; s < 16; s++) {
        double pn[NL];
        if (s + 1 < 16) {
            const double v = vcol[(s + 1) * L4_VROW];
            win[(s + 1) & 15] = v;
#pragma unroll
            for (int lag = 0; lag < NL; lag++) pn[lag] = __dmul_rn(win[(s + 1 - lag) & 15], v);
        }
        if (!EDGE) {
#pragma unroll
            for (int lag = 0; lag < NL; lag++) acc[lag] = __dadd_rn(acc[lag], p[lag]);
        } else {   // a lag only counts the samples that exist
            const uint32_t idx = idx0 + s;
#pragma unroll
            for (int lag = 0; lag < NL; lag++)
                if (idx < n && idx >= (uint32_t)lag) acc[lag] = __dadd_rn(acc[lag], p[lag]);
        }
        if (s + 1 < 16) {
#pragma unroll
            for (int lag = 0; lag < NL; lag++) p[lag] = pn[lag];
        }
    }
}

// WARPS = 4: one CTA per 32 frames, four CTAs per SM.  WARPS = 2: the PERSISTENT form -- a grid of one small CTA per SM (8192
// registers, 23 KB) that strides over the units and fits beside two CTAs of k_analyze3 built with a 112-register cap: the
// FP64 pipe, which the integer kernels leave idle, works on the next launch group meanwhile (flacb200_encode, option lpc_overlap)
template <int NL, int B, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, WARPS == 4 ? FLACB200_L4_MINB : 1) k_lpc4(EncCfg cfg, const FrameDesc* __restrict__ descs, const uint8_t* __restrict__ pcm,
                                                         const double* __restrict__ winpool, LpcRec* __restrict__ out, uint32_t nframes)
{
    constexpr int MM = NL - 1;   // largest order of this instantiation
    __shared__ __align__(16) uint8_t l4_sm[WARPS * L4_SLOTS * L4_SLOT_BYTES];
    __shared__ double l4_vals[WARPS][16][L4_VROW];   // windowed samples of the tile at hand: [sample][8 k + frame]
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t M = cfg.max_lpc_order;
    for (uint32_t unit = blockIdx.x * WARPS + wid; (unsigned long long)unit * 8 < nframes; unit += gridDim.x * WARPS) {
    __syncwarp();   // (persistent form: every lane has left the previous unit's staging slots)
    const uint32_t f0 = unit * 8;
    const uint32_t nfr = min(8u, nframes - f0);
    const uint32_t fl = lane >> 2, k = lane & 3;   // the lane's frame and channel combination
    const bool live = fl < nfr;
    const FrameDesc d = descs[f0 + (live ? fl : 0u)];
    const uint32_t n = (live && d.n > M) ? d.n : 0u;   // n <= M: InsufficientLpcSamples (:3300), nothing to analyse
    const uint8_t* fp = pcm + d.pcm_off * (unsigned long long)(2 * B);
    const uint8_t* wp = reinterpret_cast<const uint8_t*>(winpool + d.win_off);
    if (live) out[(size_t)f0 * 4 + lane].ok = 0;
    uint32_t nmax = n, nmin = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        nmin = min(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    }
    const uint32_t ntiles = (nmax + 15) / 16;
    uint8_t* wsm = l4_sm + (size_t)wid * (L4_SLOTS * L4_SLOT_BYTES);
    const uint32_t wsm_sa = (uint32_t)__cvta_generic_to_shared(wsm);
    const bool big = cfg.pcm_kind == 1;
    // PRMT selectors by the sample's byte offset within its word: the B bytes end up in the register's top bytes, in value order
    uint32_t sel[4];
#pragma unroll
    for (int sh = 0; sh < 4; sh++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int dst = big ? 3 - b : 4 - B + b;   // big endian: the first byte is the most significant
            v |= (uint32_t)(sh + b) << (4 * dst);
        }
        sel[sh] = v;
    }
    const uint32_t fa = lane & 7, qa = lane >> 3;   // phase A: the lane's frame and quarter of the tile

    // global -> slot: the four lanes of a frame copy its 2 B + 8 sixteen-byte chunks of the tile (PCM, then window values);
    // bytes past the end of the block are zero-filled
    constexpr uint32_t TB = 32 * B;   // PCM bytes of a tile
    auto issue = [&](uint32_t tile, uint32_t slot) {
        if (tile < ntiles) {
            const uint32_t base = wsm_sa + slot * L4_SLOT_BYTES + fl * L4_STRIDE;
#pragma unroll
            for (int j = 0; j < (2 * B + 8 + 3) / 4; j++) {
                const uint32_t c = k + 4 * j;
                if (c < 2 * B) {
                    const uint32_t off = tile * TB + c * 16, tot = n * 2 * B;
                    const uint32_t bytes = off < tot ? min(16u, tot - off) : 0u;
                    l3_cp16(base + c * 16, fp + (bytes ? off : 0u), bytes);
                } else if (c < 2 * B + 8) {
                    const uint32_t cc = c - 2 * B;
                    const uint32_t off = tile * 128 + cc * 16, tot = n * 8;
                    const uint32_t bytes = off < tot ? min(16u, tot - off) : 0u;
                    l3_cp16(base + 8 * L4_STRIDE + cc * 16, wp + (bytes ? off : 0u), bytes);
                }
            }
        }
        l3_commit();
    };

    double acc[NL];
    uint32_t mask = 0, wasted = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && !__any_sync(0xffffffffu, wasted != 0)) break;   // rare pass: wasted bits shifted out (:2878-2898)
        uint32_t wk[4], mk[4] = {0, 0, 0, 0};   // wasted bits / OR masks of the four candidates of the lane's phase-A frame
#pragma unroll
        for (int kk = 0; kk < 4; kk++) wk[kk] = __shfl_sync(0xffffffffu, wasted, 4 * fa + kk);
        double win[16];
#pragma unroll
        for (int i = 0; i < 16; i++) win[i] = 0.0;
#pragma unroll
        for (int lag = 0; lag < NL; lag++) acc[lag] = -0.0;   // Iterator::sum::<f64>() folds from -0.0
        issue(0, 0);
        issue(1, 1);
        uint32_t slot = 0;
        for (uint32_t t = 0; t < ntiles; t++) {
            issue(t + 2, slot >= 1 ? slot - 1 : 2);   // (slot + 2) % 3
            l3_wait<2>();
            __syncwarp();
            const uint8_t* spcm = wsm + slot * L4_SLOT_BYTES + fa * L4_STRIDE;
            double* vout = &l4_vals[wid][4 * qa][fa];
            if (pass == 0) l4_convert<B, false>(spcm + qa * (8 * B), spcm + 8 * L4_STRIDE + qa * 32, vout, sel, wk, mk);
            else l4_convert<B, true>(spcm + qa * (8 * B), spcm + 8 * L4_STRIDE + qa * 32, vout, sel, wk, mk);
            __syncwarp();   // the value buffer is complete (and the staging slot may be refilled)
            const double* vcol = &l4_vals[wid][0][8 * k + fl];
            if (t == 0 || (t + 1) * 16 > nmin) l4_accumulate<NL, true>(vcol, t * 16, n, acc, win);
            else l4_accumulate<NL, false>(vcol, t * 16, n, acc, win);
            slot = slot == 2 ? 0 : slot + 1;
        }
        l3_wait<0>();
        __syncwarp();
        if (pass == 0) {   // the masks of a frame's candidates: OR over its four phase-A lanes, then to the candidates' own lanes
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                mk[kk] |= __shfl_xor_sync(0xffffffffu, mk[kk], 8);
                mk[kk] |= __shfl_xor_sync(0xffffffffu, mk[kk], 16);
                const uint32_t t = __shfl_sync(0xffffffffu, mk[kk], fl);
                if (k == (uint32_t)kk) mask = t;
            }
            wasted = (mask == 0 || (mask & 1u)) ? 0u : (uint32_t)__ffs((int)mask) - 1u;
        }
    }
    // ---- Levinson-Durbin (:3536-3580) with the order estimate (:3656-3702) folded in: all lanes, everything in registers ----
    const bool run = n != 0 && mask != 0;   // all-zero candidates become CONSTANT (:2883)
    if (!run) continue;
    const uint32_t bps = cand_bps(cfg, k) - wasted;
    const uint32_t precision = lpc_precision_for(n);
    const double error_scale = __ddiv_rn(0.5, (double)n);
    const double divisor = 2.0 * 0.693147180559945309417232121458176568;
    double a[MM], best_a[MM];
#pragma unroll
    for (int j = 0; j < MM; j++) { a[j] = 0.0; best_a[j] = 0.0; }
    int best = 0;
    double best_bits = 0.0, err;
    bool alive = true;   // take_while(err > 0)
    {
        const double kk = __ddiv_rn(acc[1], acc[0]);
        a[0] = kk;
        err = __dmul_rn(acc[0], __dsub_rn(1.0, __dmul_rn(kk, kk)));
    }
#pragma unroll
    for (int i = 0; i < MM; i++) {   // here a[0..i] is the set of order i + 1 and err its error
        if ((uint32_t)i < M) {
            alive = alive && err > 0.0;
            if (alive) {
                const uint32_t o = i + 1;
                const double bpr = __ddiv_rn(glibc_log(__dmul_rn(err, error_scale)), divisor);
                const double bits = fma(bpr, (double)(n - o), (double)(o * (bps + precision)));
                if (best == 0 || total_key(bits) < total_key(best_bits)) {   // first minimum under total_cmp
                    best = (int)o;
                    best_bits = bits;
#pragma unroll
                    for (int j = 0; j <= i; j++) best_a[j] = a[j];
                }
            }
            if (i + 1 < MM && (uint32_t)(i + 1) < M) {   // the set of order i + 2
                double s = -0.0;
#pragma unroll
                for (int j = 0; j <= i; j++) s = __dadd_rn(s, __dmul_rn(acc[i + 1 - j], a[j]));
                const double q = __dsub_rn(acc[i + 2], s);
                const double kk = __ddiv_rn(q, err);
                double b[MM];
#pragma unroll
                for (int j = 0; j <= i; j++) b[j] = __dsub_rn(a[j], __dmul_rn(kk, a[i - j]));
#pragma unroll
                for (int j = 0; j <= i; j++) a[j] = b[j];
                a[i + 1] = kk;
                err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(kk, kk)));
            }
        }
    }
    if (best == 0) continue;
    // quantize (:3334-3401)
    double l = fabs(best_a[0]);
#pragma unroll
    for (int j = 1; j < MM; j++) {
        if (j < best) {
            const double aj = fabs(best_a[j]);
            if (total_key(aj) >= total_key(l)) l = aj;
        }
    }
    if (!(l > 0.0)) continue;
    const int32_t max_coeff = (1 << (precision - 1)) - 1, min_coeff = -(1 << (precision - 1));
    const int32_t lg = f64_as_i32_sat(floor(glibc_log2(l)));
    long long sh = (long long)((int32_t)precision - 1) - (long long)lg - 1;   // :3360
    if (sh > 15) sh = 15;
    if (sh < -16) continue;
    LpcRec rec;
    double error = 0.0;
    const double scale = (double)(1 << (sh >= 0 ? sh : -sh));
#pragma unroll
    for (int j = 0; j < MAX_LPC; j++) {
        int32_t q = 0;
        if (j < MM && j < best) {
            const double sum = sh >= 0 ? fma(best_a[j < MM ? j : 0], scale, error)                        // mul_add :3372
                                       : __dadd_rn(__ddiv_rn(best_a[j < MM ? j : 0], scale), error);   // :3391
            q = f64_as_i32_sat(round(sum));
            q = q < min_coeff ? min_coeff : (q > max_coeff ? max_coeff : q);
            error = __dsub_rn(sum, (double)q);
        }
        rec.q[j] = (int16_t)q;
    }
    rec.shift = (uint8_t)(sh >= 0 ? sh : 0);
    rec.ok = 1;
    rec.order = (uint8_t)best;
    rec.precision = (uint8_t)precision;
    rec.pad = 0;
    out[(size_t)f0 * 4 + lane] = rec;
    }
}

// stereo frames, 16- or 24-bit packed PCM on 16-byte boundaries, order <= 15
bool lpc4_ok(const EncCfg& cfg, bool blocks_aligned16)
{
    return lpc3_ok(cfg, blocks_aligned16) && (cfg.bytes_per_sample == 2 || cfg.bytes_per_sample == 3);
}

// max_ctas: 0 = a warp per eight frames, four warps per CTA; else the persistent two-warp form with at most that many CTAs
cudaError_t launch_lpc4(const EncCfg& cfg, const FrameDesc* descs, const uint8_t* pcm, const double* winpool, LpcRec* lpcs, uint32_t max_ctas,
                        cudaStream_t st)
{
    const uint32_t units = (cfg.nframes + 7) / 8;
    const uint32_t M = cfg.max_lpc_order;
#define FLACB200_L4(NLV, BV)                                                                                                                    \
    do {                                                                                                                                        \
        if (max_ctas) count_launch(), k_lpc4<NLV, BV, 2><<<min(max_ctas, (units + 1) / 2), 64, 0, st>>>(cfg, descs, pcm, winpool, lpcs, cfg.nframes); \
        else count_launch(), k_lpc4<NLV, BV, 4><<<(units + 3) / 4, 128, 0, st>>>(cfg, descs, pcm, winpool, lpcs, cfg.nframes);                  \
    } while (0)
    if (cfg.bytes_per_sample == 3) {
        if (M <= 8) FLACB200_L4(9, 3);
        else if (M <= 12) FLACB200_L4(13, 3);
        else FLACB200_L4(16, 3);
    } else {
        if (M <= 8) FLACB200_L4(9, 2);
        else if (M <= 12) FLACB200_L4(13, 2);
        else FLACB200_L4(16, 2);
    }
#undef FLACB200_L4
    return cudaGetLastError();
}

}   // namespace flacb200
