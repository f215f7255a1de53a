// rice.cuh -- Rice partition search pieces shared by the analysis kernels (k_residual, k_analyze, k_analyze3):
// Partition::new, best_partitions and try_reduce_rice of src/encode.rs:3747-3942.
#pragma once
#include "common.cuh"

namespace flacb200 {

struct RiceChoice {
    uint32_t resid_bits;   // bits of the whole residual block incl. method/order/partition headers
    uint32_t fail;
    uint8_t method, porder_w, porder_g, nparts;
    uint8_t rice[MAX_PARTS];
};

// Partition::new (src/encode.rs:3765-3831): code and estimated bits of one partition
__device__ inline uint8_t partition_code(unsigned long long sum, uint32_t len, uint32_t rice_max, uint32_t* est)
{
    *est = 0;
    const uint32_t samples = len & 0xffffu;   // `as u16`
    if (samples == 0) return 0xFF;
    if (sum == 0) return 0x80;                // all-zero partition (:3826)
    uint32_t rice = 0;
    if (sum > samples) {
        // ceil(log2(sum / samples)) == min{k : samples << k >= sum}; equality with the reference's f64 form
        // is checked in tests/test_oracle_kat.py::test_rice_parameter_integer_equivalence
        // start from the bit-length difference (at most one below the answer), then step
        const uint32_t lg_sum = 63u - (uint32_t)__clzll((long long)sum), lg_n = 31u - (uint32_t)__clz((int)samples);
        rice = lg_sum > lg_n ? lg_sum - lg_n : 0u;
        while (((unsigned long long)samples << rice) < sum) rice++;
        if (rice >= rice_max) {
            const uint32_t escape = (63u - (uint32_t)__clzll((long long)sum)) + 2u;   // ilog2(sum) + 2 (:3787)
            if (escape > 31) return 0xFF;
            *est = escape * samples;
            return (uint8_t)(0x40 | escape);
        }
    }
    const unsigned long long t = rice > 0 ? (sum >> (rice - 1)) : (sum << 1);   // :3811-3815
    if (t > 0xffffffffull) return 0xFF;
    *est = 4u + ((1u + rice) * samples) + (uint32_t)t - (samples / 2u);
    return (uint8_t)rice;
}

struct AwSmemLimbs {
    uint32_t* lo;
    uint32_t* hi;
};

// rare path: a 16-sample tile that straddles a partition boundary, contains samples before the predictor order, or is cut by the block end
static __device__ __noinline__ void aw_tile_sums_slow(const int32_t* r, uint32_t i0, uint32_t first, uint32_t end, uint32_t cf, uint32_t* lo, uint32_t* hi)
{
    for (uint32_t i = max(i0, first); i < min(i0 + 16u, end); i++) {
        const uint32_t v = uabs32(r[i - i0]);
        if (v) {
            atomicAdd(&lo[i / cf], v & 0xFFFFFFu);
            atomicAdd(&hi[i / cf], v >> 24);
        }
    }
}

static __device__ __noinline__ void aw_tile_bits_slow(const int32_t* r, uint32_t i0, uint32_t first, uint32_t end, uint32_t cp, uint32_t j0,
                                               const uint8_t* rice, unsigned long long* bits_io, uint32_t* bad_io)
{
    unsigned long long bits = *bits_io;
    uint32_t bad = *bad_io;
    for (uint32_t i = max(i0, first); i < min(i0 + 16u, end); i++) {
        const uint32_t c = rice[i / cp - j0];
        const int32_t s = r[i - i0];
        if (c < 0x40) bits += (zigzag32(s) >> c) + 1u + c;
        else if (c & 0x40) {
            const uint32_t w = c & 31u;
            bits += w;
            if (s < -(1 << (w - 1)) || s > (1 << (w - 1)) - 1) bad = 1;   // write_signed_counted fails
        }
    }
    *bits_io = bits;
    *bad_io = bad;
}

// pass-1 work of a rare tile (first tile of the block, tile cut by the block end, tile straddling a partition boundary):
// fixed orders 0..4 and the LPC residual rl, sample by sample
static __device__ __noinline__ void aw_pass1_tile_slow(const int32_t* x, const int32_t* h, const int32_t* rl, bool have_lpc, uint32_t order, uint32_t i0,
                                                uint32_t n, uint32_t kmax, uint32_t cf, AwSmemLimbs limbs, unsigned long long* u)
{
    int32_t r[5][16];
    int32_t p1 = h[15] - h[14], p2 = p1 - (h[14] - h[13]), p3 = p2 - ((h[14] - h[13]) - (h[13] - h[12]));
    int32_t prev = h[15];
    for (int e = 0; e < 16; e++) {
        const int32_t e1 = x[e] - prev, e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
        prev = x[e]; p1 = e1; p2 = e2; p3 = e3;
        r[0][e] = x[e]; r[1][e] = e1; r[2][e] = e2; r[3][e] = e3; r[4][e] = e4;
    }
    for (uint32_t k = 0; k < 5; k++) {
        aw_tile_sums_slow(r[k], i0, k, n, cf, limbs.lo + k * MAX_PARTS, limbs.hi + k * MAX_PARTS);
        if (i0 == 0) {   // what set k counts but the order comparison (:3062-3073, samples >= kmax only) does not
            unsigned long long v = 0;
            for (uint32_t i = k; i < kmax && i < 16 && i < n; i++) v += uabs32(r[k][i]);
            u[k] = v;
        }
    }
    if (have_lpc) aw_tile_sums_slow(rl, i0, order, n, cf, limbs.lo + 5 * MAX_PARTS, limbs.hi + 5 * MAX_PARTS);
}

__device__ inline void aw_add_limbs(uint32_t* lo, uint32_t* hi, uint32_t chunk, unsigned long long v)
{
    atomicAdd(&lo[chunk], (uint32_t)v & 0xFFFFFFu);   // unconditional: a branch costs more than an atomic that adds 0
    atomicAdd(&hi[chunk], (uint32_t)(v >> 24));
}

// best_partitions + try_reduce_rice (src/encode.rs:3865-3942) for one residual set, by one warp.
// tree[] holds the per-partition sums of every order; fills ch (rice[], geometry, method) -- not the size.
static __device__ void aw_choose_partitions(const EncCfg& cfg, uint32_t n, uint32_t o, uint32_t p_max, const unsigned long long* tree, uint32_t* part_est,
                                     uint8_t* part_code, RiceChoice& ch)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rice_max = cfg.use_rice2 ? 31u : 15u;
    for (uint32_t t = lane; t < 127; t += 32) {   // (order p, partition j)
        const uint32_t p = 31u - (uint32_t)__clz((int)(t + 1));
        const uint32_t j = t + 1 - (1u << p);
        uint8_t code = 0xFE;
        uint32_t est = 0;
        if (p <= p_max) {
            const uint32_t cp = n >> p;
            const uint32_t a = j * cp, b = a + cp;
            if (b > o) code = partition_code(tree[t], b - max(a, o), rice_max, &est);
        }
        part_code[t] = code;
        part_est[t] = est;
    }
    __syncwarp();
    // lane p ends up with the totals of order p; every order is summed by the whole warp
    uint32_t est = 0, cnt = 0, bad = 0;
    for (uint32_t p = 0; p <= p_max; p++) {
        const uint32_t base = (1u << p) - 1;
        uint32_t e1 = 0, c1 = 0, b1 = 0;
        for (uint32_t j = lane; j < (1u << p); j += 32) {
            const uint8_t c = part_code[base + j];
            if (c == 0xFE) continue;
            if (c == 0xFF) b1 = 1;
            c1++;
            e1 += part_est[base + j];
        }
        e1 = __reduce_add_sync(0xffffffffu, e1);
        c1 = __reduce_add_sync(0xffffffffu, c1);
        b1 = __reduce_or_sync(0xffffffffu, b1);
        if (lane == p) { est = e1; cnt = c1; bad = b1; }
    }
    const bool ok = lane <= p_max && !bad && cnt != 0 && (cnt & (cnt - 1)) == 0;   // :3880-3881
    const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
    if (okmask == 0) {   // unwrap_or_else (:3887): one partition escaped at 31 bits
        if (lane == 0) {
            ch.porder_g = 0; ch.porder_w = 0; ch.nparts = 1; ch.rice[0] = 0x40 | 31;
            ch.method = cfg.use_rice2 ? 1 : 0;
        }
        __syncwarp();
        return;
    }
    const uint32_t best_est = __reduce_min_sync(0xffffffffu, ok ? est : 0xFFFFFFFFu);
    const uint32_t best_p = (uint32_t)__ffs((int)(__ballot_sync(0xffffffffu, ok && est == best_est))) - 1u;   // first minimum :3885
    const uint32_t best_count = __shfl_sync(0xffffffffu, cnt, best_p);
    const uint32_t base = (1u << best_p) - 1, j0 = (1u << best_p) - best_count;
    uint32_t big = 0;
    for (uint32_t j = lane; j < best_count; j += 32) {
        const uint8_t c = part_code[base + j0 + j];
        ch.rice[j] = c;
        if (c < 0x40 && c >= 15) big = 1;
    }
    big = __any_sync(0xffffffffu, big);
    if (lane == 0) {
        ch.porder_g = (uint8_t)best_p;
        ch.nparts = (uint8_t)best_count;
        ch.porder_w = (uint8_t)(31u - (uint32_t)__clz((int)best_count));   // partitions.len().ilog2() :3902
        ch.method = (cfg.use_rice2 && big) ? 1 : 0;                         // try_reduce_rice :3929-3942
    }
    __syncwarp();
}


// best_partitions + try_reduce_rice (src/encode.rs:3865-3942) for one residual set, by one warp, from the sums of the FINEST
// partitions (lo/hi: 24-bit limbs per leaf, 1 << p_max leaves).  Same result as the tree walk of aw_choose_partitions, in a
// fraction of the instructions: the sum of partition j at order p is a difference of two entries of the prefix sums over
// the leaves, every (order, partition) node is evaluated once in registers (four per lane), the per-order totals come from
// warp reductions that all lanes receive -- no level-by-level tree, no per-order loops.  `pref` (65 entries) and `codes`
// (127 bytes) are the warp's scratch in shared memory.
// lb_out (optional): a LOWER BOUND of the exact size of the residual codes with the chosen parameters, from the partition sums
// alone -- zig-zag u is 2|r| or 2|r| - 1, so u >> k >= 2|r| / 2^k - 1 and a partition of len samples with parameter k costs at
// least max(0, ceil(2 S / 2^k) - len) + len (k + 1) bits; escaped and all-zero partitions are exact.
static __device__ void aw_choose_partitions_flat(const EncCfg& cfg, uint32_t n, uint32_t o, uint32_t p_max, const uint32_t* lo, const uint32_t* hi,
                                                unsigned long long* pref, uint8_t* codes, RiceChoice& ch, uint32_t* lb_out = nullptr)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rice_max = cfg.use_rice2 ? 31u : 15u;
    const uint32_t nleaf = 1u << p_max;
    // ---- prefix sums over the leaves: lane l owns leaves 2 l and 2 l + 1 ----
    const unsigned long long v0 = 2 * lane < nleaf ? (unsigned long long)lo[2 * lane] + ((unsigned long long)hi[2 * lane] << 24) : 0ull;
    const unsigned long long v1 = 2 * lane + 1 < nleaf ? (unsigned long long)lo[2 * lane + 1] + ((unsigned long long)hi[2 * lane + 1] << 24) : 0ull;
    unsigned long long incl = v0 + v1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    pref[2 * lane] = incl - v0 - v1;
    pref[2 * lane + 1] = incl - v1;
    if (lane == 31) pref[64] = incl;
    __syncwarp();
    // ---- every node (order p, partition j) once: t = (1 << p) - 1 + j, four nodes per lane ----
    uint32_t est[4];
    uint8_t code[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t t = lane + 32u * k;
        const uint32_t p = 31u - (uint32_t)__clz((int)(t + 1));
        const uint32_t j = t + 1 - (1u << p);
        code[k] = 0xFE;
        est[k] = 0;
        if (t < 127 && p <= p_max) {
            const uint32_t cp = n >> p, a = j * cp, b = a + cp, w = nleaf >> p;
            if (b > o) code[k] = partition_code(pref[(j + 1) * w] - pref[j * w], b - max(a, o), rice_max, &est[k]);
        }
        if (t < 127) codes[t] = code[k];
    }
    // ---- totals per order (nodes of order p: t in [2^p - 1, 2^(p+1) - 2]); every lane gets all of them ----
    // order 0..4 live in slot 0 (lanes 0, 1-2, 3-6, 7-14, 15-30), order 5 = slot 0 lane 31 + slot 1 lanes 0-30, order 6 = the rest
    const uint32_t p0 = 31u - (uint32_t)__clz((int)(lane + 1));   // order of the slot-0 node of this lane (5 for lane 31)
    uint32_t e_tot[7], cntw0 = 0, cntw1 = 0, badm = 0;
#pragma unroll
    for (int p = 0; p < 5; p++) e_tot[p] = __reduce_add_sync(0xffffffffu, p0 == (uint32_t)p ? est[0] : 0u);
    e_tot[5] = __reduce_add_sync(0xffffffffu, (lane == 31 ? est[0] : 0u) + (lane < 31 ? est[1] : 0u));
    e_tot[6] = __reduce_add_sync(0xffffffffu, (lane == 31 ? est[1] : 0u) + est[2] + (lane < 31 ? est[3] : 0u));
    {   // partition counts (8-bit fields: at most 64 per order) and "a partition of this order cannot be coded" flags
        auto tally = [&](uint8_t c, uint32_t p) {
            if (c == 0xFE) return;
            if (c == 0xFF) badm |= 1u << p;
            if (p < 4) cntw0 += 1u << (8 * p);
            else cntw1 += 1u << (8 * (p - 4));
        };
        tally(code[0], p0);
        tally(code[1], lane == 31 ? 6u : 5u);
        tally(code[2], 6u);
        if (lane < 31) tally(code[3], 6u);
        cntw0 = __reduce_add_sync(0xffffffffu, cntw0);
        cntw1 = __reduce_add_sync(0xffffffffu, cntw1);
        badm = __reduce_or_sync(0xffffffffu, badm);
    }
    uint32_t best_p = 0xFFFFFFFFu, best_est = 0, best_count = 0;
#pragma unroll
    for (int p = 0; p < 7; p++) {
        const uint32_t cnt = p < 4 ? (cntw0 >> (8 * p)) & 0xffu : (cntw1 >> (8 * (p - 4))) & 0xffu;
        const bool ok = (uint32_t)p <= p_max && !((badm >> p) & 1u) && cnt != 0 && (cnt & (cnt - 1)) == 0;   // :3880-3881
        if (ok && (best_p == 0xFFFFFFFFu || e_tot[p] < best_est)) {   // first minimum :3885
            best_p = (uint32_t)p;
            best_est = e_tot[p];
            best_count = cnt;
        }
    }
    __syncwarp();
    if (best_p == 0xFFFFFFFFu) {   // unwrap_or_else (:3887): one partition escaped at 31 bits
        if (lane == 0) {
            ch.porder_g = 0; ch.porder_w = 0; ch.nparts = 1; ch.rice[0] = 0x40 | 31;
            ch.method = cfg.use_rice2 ? 1 : 0;
            if (lb_out) *lb_out = 0;
        }
        __syncwarp();
        return;
    }
    const uint32_t base = (1u << best_p) - 1, j0 = (1u << best_p) - best_count;
    uint32_t big = 0;
    for (uint32_t j = lane; j < best_count; j += 32) {
        const uint8_t c = codes[base + j0 + j];
        ch.rice[j] = c;
        if (c < 0x40 && c >= 15) big = 1;
    }
    big = __any_sync(0xffffffffu, big);
    if (lane == 0) {
        ch.porder_g = (uint8_t)best_p;
        ch.nparts = (uint8_t)best_count;
        ch.porder_w = (uint8_t)(31u - (uint32_t)__clz((int)best_count));   // partitions.len().ilog2() :3902
        ch.method = (cfg.use_rice2 && big) ? 1 : 0;                         // try_reduce_rice :3929-3942
    }
    if (lb_out) {
        const uint32_t cp = n >> best_p, w = nleaf >> best_p;
        unsigned long long lb = 0;
        for (uint32_t j = lane; j < best_count; j += 32) {
            const uint32_t jj = j0 + j, a = jj * cp, b = a + cp, len = b - max(a, o);
            const uint32_t c = codes[base + jj];
            const unsigned long long S = pref[(jj + 1) * w] - pref[jj * w];
            if (c < 0x40) {
                const unsigned long long t = (2 * S + ((1ull << c) - 1)) >> c;
                lb += (t > len ? t - len : 0ull) + (unsigned long long)len * (c + 1u);
            } else if (c & 0x40) {
                lb += (unsigned long long)len * (c & 31u);
            }
        }
        lb = warp_sum_u64(lb);
        if (lane == 0) *lb_out = lb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)lb;
    }
    __syncwarp();
}

}   // namespace flacb200
