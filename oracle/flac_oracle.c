/*
 * flac_oracle.c -- CPU restatement of tuffy/flac-codec 1.3.2's frame encode/decode path.
 *
 * TEST INFRASTRUCTURE ONLY (see flac_oracle.h).  Every function cites the reference
 * file:line it follows (paths relative to the reference crate root).  Build with
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (see oracle/Makefile)
 * -ffp-contract=off matters: the reference (rustc) never fuses a*b+c unless mul_add is
 * written, and the f64 sums below must round exactly as the reference's do.
 */
#include "flac_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_LPC 32
#define MAX_PARTITIONS 64 /* src/encode.rs:3756 */
#define MAX_CHANNELS 8

/* ------------------------------------------------------------------------------------------
 * CRC-8 (poly 0x07) and CRC-16 (poly 0x8005), init 0, MSB first: src/crc.rs:100-188.
 * Tables are generated from the polynomials rather than transcribed.
 * ---------------------------------------------------------------------------------------- */
static uint8_t crc8_table[256];
static uint16_t crc16_table[256];
static int tables_ready = 0;

static void init_tables(void)
{
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        for (int b = 0; b < 8; b++) c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
        crc8_table[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int b = 0; b < 8; b++) d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1));
        crc16_table[i] = d;
    }
    tables_ready = 1;
}

void fo_libm(int fn, const double* in, double* out, size_t n)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) out[i] = fn == 0 ? log(in[i]) : log2(in[i]);
}

uint8_t fo_crc8(const uint8_t* p, size_t n)
{
    init_tables();
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) c = crc8_table[c ^ p[i]]; /* src/crc.rs:128 */
    return c;
}

uint16_t fo_crc16(const uint8_t* p, size_t n)
{
    init_tables();
    uint16_t c = 0;
    for (size_t i = 0; i < n; i++) c = (uint16_t)(crc16_table[(uint8_t)(c >> 8) ^ p[i]] ^ (uint16_t)(c << 8)); /* :181 */
    return c;
}

/* ------------------------------------------------------------------------------------------
 * MD5 (RFC 1321) -- the reference uses the `md5` crate for STREAMINFO (src/encode.rs:376, :2100)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t s[4];
    uint64_t len;
    uint8_t buf[64];
    uint32_t fill;
} md5_ctx;

static uint32_t md5_k[64];
static int md5_ready = 0;
static const uint8_t md5_r[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5,  9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4,  11, 16, 23, 6,  10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_init(md5_ctx* c)
{
    if (!md5_ready) {
        for (int i = 0; i < 64; i++) md5_k[i] = (uint32_t)(int64_t)floor(fabs(sin((double)(i + 1))) * 4294967296.0);
        md5_ready = 1;
    }
    c->s[0] = 0x67452301u;
    c->s[1] = 0xefcdab89u;
    c->s[2] = 0x98badcfeu;
    c->s[3] = 0x10325476u;
    c->len = 0;
    c->fill = 0;
}

static void md5_block(md5_ctx* c, const uint8_t* p)
{
    uint32_t w[16];
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = c->s[0], b = c->s[1], cc = c->s[2], d = c->s[3];
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) {
            f = (b & cc) | (~b & d);
            g = i;
        } else if (i < 32) {
            f = (d & b) | (~d & cc);
            g = (5 * i + 1) & 15;
        } else if (i < 48) {
            f = b ^ cc ^ d;
            g = (3 * i + 5) & 15;
        } else {
            f = cc ^ (b | ~d);
            g = (7 * i) & 15;
        }
        uint32_t t = d;
        d = cc;
        cc = b;
        uint32_t x = a + f + md5_k[i] + w[g];
        b = b + ((x << md5_r[i]) | (x >> (32 - md5_r[i])));
        a = t;
    }
    c->s[0] += a;
    c->s[1] += b;
    c->s[2] += cc;
    c->s[3] += d;
}

static void md5_update(md5_ctx* c, const uint8_t* p, size_t n)
{
    c->len += n;
    if (c->fill) {
        while (n && c->fill < 64) {
            c->buf[c->fill++] = *p++;
            n--;
        }
        if (c->fill == 64) {
            md5_block(c, c->buf);
            c->fill = 0;
        }
    }
    while (n >= 64) {
        md5_block(c, p);
        p += 64;
        n -= 64;
    }
    while (n) {
        c->buf[c->fill++] = *p++;
        n--;
    }
}

static void md5_final(md5_ctx* c, uint8_t out[16])
{
    uint64_t bits = c->len * 8;
    uint8_t pad = 0x80;
    md5_update(c, &pad, 1);
    pad = 0;
    while (c->fill != 56) md5_update(c, &pad, 1);
    uint8_t l[8];
    for (int i = 0; i < 8; i++) l[i] = (uint8_t)(bits >> (8 * i));
    md5_update(c, l, 8);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(c->s[i] >> (8 * j));
}

void fo_md5(const uint8_t* p, size_t n, uint8_t out[16])
{
    md5_ctx c;
    md5_init(&c);
    md5_update(&c, p, n);
    md5_final(&c, out);
}

/* ------------------------------------------------------------------------------------------
 * Bit recorder / writer.  Semantics of bitstream-io's BitWriter<_, BigEndian> / BitRecorder as
 * the reference uses them: MSB first; write_signed_counted(n, v) = n-bit
 * two's complement and fails if v does not fit; write_unary::<1>(q) = q zeros then a one;
 * byte_align pads with zero bits; written() = exact bit count.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t* data;
    size_t cap;     /* bytes */
    uint64_t nbits; /* bits written */
    int err;
} bits_t;

static void bits_reserve(bits_t* b, uint64_t more_bits)
{
    size_t need = (size_t)((b->nbits + more_bits + 7) / 8) + 8;
    if (need > b->cap) {
        size_t ncap = b->cap ? b->cap * 2 : 4096;
        while (ncap < need) ncap *= 2;
        b->data = (uint8_t*)realloc(b->data, ncap);
        memset(b->data + b->cap, 0, ncap - b->cap);
        b->cap = ncap;
    }
}

static void bits_clear(bits_t* b)
{
    size_t used = (size_t)((b->nbits + 7) / 8);
    if (b->data && used) memset(b->data, 0, used < b->cap ? used + 1 : b->cap);
    b->nbits = 0;
    b->err = 0;
}

/* write the low `n` bits of v (n <= 32), MSB first; buffer bytes past nbits are always zero */
static inline void bits_put(bits_t* b, uint32_t n, uint32_t v)
{
    if (n == 0) return;
    bits_reserve(b, n);
    if (n < 32) v &= (1u << n) - 1u;
    uint64_t pos = b->nbits;
    uint32_t left = n;
    while (left) {
        uint32_t bit_in_byte = (uint32_t)(pos & 7);
        uint32_t room = 8 - bit_in_byte;
        uint32_t take = left < room ? left : room;
        uint32_t chunk = (v >> (left - take)) & ((1u << take) - 1u);
        b->data[pos >> 3] |= (uint8_t)(chunk << (room - take));
        pos += take;
        left -= take;
    }
    b->nbits = pos;
}

static inline void bits_put64(bits_t* b, uint32_t n, uint64_t v)
{
    if (n > 32) {
        bits_put(b, n - 32, (uint32_t)(v >> 32));
        bits_put(b, 32, (uint32_t)v);
    } else {
        bits_put(b, n, (uint32_t)v);
    }
}

/* write_signed_counted: value must be representable in n bits (n in 1..=32) */
static inline void bits_put_signed(bits_t* b, uint32_t n, int32_t v)
{
    if (n < 32) {
        int32_t lo = -(int32_t)(1u << (n - 1)), hi = (int32_t)((1u << (n - 1)) - 1u);
        if (v < lo || v > hi) {
            b->err = FO_ERR_IO;
            return;
        }
    }
    bits_put(b, n, (uint32_t)v);
}

/* write_unary::<1>(q): q zero bits then a one bit */
static inline void bits_put_unary1(bits_t* b, uint32_t q)
{
    bits_reserve(b, (uint64_t)q + 1);
    b->nbits += q; /* buffer is pre-zeroed */
    bits_put(b, 1, 1);
}

/* write_unary::<0>(q): q one bits then a zero bit */
static inline void bits_put_unary0(bits_t* b, uint32_t q)
{
    for (uint32_t i = 0; i < q; i++) bits_put(b, 1, 1);
    bits_put(b, 1, 0);
}

static inline void bits_align(bits_t* b)
{
    uint32_t r = (uint32_t)(b->nbits & 7);
    if (r) {
        bits_reserve(b, 8 - r);
        b->nbits += 8 - r;
    }
}

/* BitRecorder::playback: append all recorded bits of src to dst */
static void bits_append(bits_t* dst, const bits_t* src)
{
    uint64_t n = src->nbits;
    bits_reserve(dst, n);
    size_t full = (size_t)(n >> 3);
    if ((dst->nbits & 7) == 0) {
        memcpy(dst->data + (dst->nbits >> 3), src->data, full);
        dst->nbits += (uint64_t)full * 8;
    } else {
        for (size_t i = 0; i < full; i++) bits_put(dst, 8, src->data[i]);
    }
    uint32_t rem = (uint32_t)(n & 7);
    if (rem) bits_put(dst, rem, (uint32_t)(src->data[full] >> (8 - rem)));
    if (src->err && !dst->err) dst->err = src->err;
}

/* ------------------------------------------------------------------------------------------
 * Options presets: src/encode.rs:1376-1408 (default), :1635-1644 (fast), :1649-1657 (best)
 * ---------------------------------------------------------------------------------------- */
void fo_options_default(fo_options* o)
{
    memset(o, 0, sizeof(*o));
    o->block_size = 4096;
    o->mid_side = 1;
    o->max_partition_order = 5;
    o->max_lpc_order = 8;
    o->window_kind = 2;
    o->tukey_p = 0.5f;
    o->exhaustive_channel_correlation = 1;
    o->seektable_kind = 1; /* SeekTableInterval::default() = Seconds(10), :1329 */
    o->seektable_n = 10;
    o->padding = 4096;
}

void fo_options_fast(fo_options* o)
{
    fo_options_default(o);
    o->block_size = 1152;
    o->mid_side = 0;
    o->max_partition_order = 3;
    o->max_lpc_order = 0;
    o->exhaustive_channel_correlation = 0;
}

void fo_options_best(fo_options* o)
{
    fo_options_default(o);
    o->block_size = 4096;
    o->mid_side = 1;
    o->max_partition_order = 6;
    o->max_lpc_order = 12;
}

/* ------------------------------------------------------------------------------------------
 * Window::generate  src/encode.rs:1725-1783
 * ---------------------------------------------------------------------------------------- */
static void window_hann(double* w, uint32_t n)
{
    double np = (double)n - 1.0; /* :1734 */
    for (uint32_t i = 0; i < n; i++) w[i] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)i / np); /* :1738 */
}

static void window_fill1(double* w, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) w[i] = 1.0;
}

static void window_tukey(double* w, uint32_t n, float p)
{
    if (p <= 0.0f) { /* ..=0.0  :1744 */
        window_fill1(w, n);
    } else if (p >= 1.0f) { /* 1.0..  :1747 */
        window_hann(w, n);
    } else if (p > 0.0f && p < 1.0f) { /* :1750 */
        double t = (double)p / 2.0 * (double)n;
        uint64_t tt = (uint64_t)t; /* `as usize` truncates */
        if (tt == 0) {             /* checked_sub(1) == None  :1773 */
            window_fill1(w, n);
            return;
        }
        uint64_t np = tt - 1;
        /* get_disjoint_mut([0..np, np..len-np, len-np..len]) fails when ranges overlap  :1769 */
        if (np > n || np > n - np) {
            window_fill1(w, n);
            return;
        }
        window_fill1(w, n);
        for (uint64_t k = 0; k < np; k++) {
            double x = 0.5 - 0.5 * cos(M_PI * (double)k / (double)np); /* :1764 */
            w[k] = x;
            w[n - 1 - k] = x; /* last.iter_mut().rev()  :1762 */
        }
    } else { /* NaN  :1778 */
        window_tukey(w, n, 0.5f);
    }
}

void fo_window(const fo_options* opt, uint32_t n, double* out)
{
    switch (opt->window_kind) {
    case 0: window_fill1(out, n); break;
    case 1: window_hann(out, n); break;
    default: window_tukey(out, n, opt->tukey_p); break;
    }
}

/* ------------------------------------------------------------------------------------------
 * autocorrelate  src/encode.rs:3478-3501.  Strict left-to-right f64 sums; Rust's
 * `Iterator::sum::<f64>()` folds from -0.0.
 * ---------------------------------------------------------------------------------------- */
int fo_autocorrelate(const double* windowed, uint32_t n, uint32_t max_lpc_order, double* out)
{
    int count = 0;
    for (uint32_t lag = 0; lag <= max_lpc_order; lag++) {
        if (lag >= n) return count; /* tail.is_empty()  :3492 */
        double s = -0.0;
        const double* tail = windowed + lag;
        uint32_t m = n - lag;
        for (uint32_t i = 0; i < m; i++) s = s + windowed[i] * tail[i]; /* :3495 */
        out[count++] = s;
    }
    return count;
}

/* ------------------------------------------------------------------------------------------
 * lp_coefficients (Levinson-Durbin, keeps every order)  src/encode.rs:3536-3580
 * coeffs[(o-1)*MAX_LPC + j] is coefficient j of the order-o predictor.
 * ---------------------------------------------------------------------------------------- */
int fo_lp_coefficients(const double* r, uint32_t n_autoc, double* coeffs, double* errors)
{
    if (n_autoc < 2) return 0; /* the reference panics; unreachable from the encoder */
    double k = r[1] / r[0]; /* :3545 */
    coeffs[0] = k;
    errors[0] = r[0] * (1.0 - k * k); /* k.powi(2)  :3548 */
    int orders = 1;
    for (uint32_t i = 1; i < n_autoc - 1; i++) { /* :3551 */
        const double* prev_c = coeffs + (size_t)(i - 1) * MAX_LPC;
        double* cur_c = coeffs + (size_t)i * MAX_LPC;
        double err = errors[i - 1];
        /* q = next - sum(prev.rev() zip coeffs)  :3555-3561; prev = r[0..=i], next = r[i+1] */
        double s = -0.0;
        for (uint32_t j = 0; j < i; j++) s = s + r[i - j] * prev_c[j];
        double q = r[i + 1] - s;
        k = q / err; /* :3563 */
        for (uint32_t j = 0; j < i; j++) cur_c[j] = prev_c[j] - k * prev_c[i - 1 - j]; /* :3566-3569 */
        cur_c[i] = k;
        errors[i] = err * (1.0 - k * k); /* :3572 */
        orders++;
    }
    return orders;
}

/* f64::total_cmp  (used by min_by at src/encode.rs:3699) */
static int total_cmp(double a, double b)
{
    int64_t x, y;
    memcpy(&x, &a, 8);
    memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
    y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return (x > y) - (x < y);
}

/* subframe_bits_by_order  src/encode.rs:3656-3684 (including the `.max(0.0)` precedence quirk:
 * the max applies to the divisor only). Returns the number of orders that pass take_while. */
int fo_subframe_bits_by_order(uint32_t bps, uint32_t precision, uint32_t sample_count, const double* errors,
                              uint32_t n_orders, double* bits_out)
{
    double error_scale = 0.5 / (double)sample_count; /* :3664 */
    int count = 0;
    for (uint32_t o = 1; o <= n_orders; o++) {
        double error = errors[o - 1];
        if (!(error > 0.0)) break; /* take_while  :3668 */
        uint32_t header_bits = o * (bps + precision); /* :3671 */
        double divisor = fmax(2.0 * M_LN2, 0.0);
        double bits_per_residual = log(error * error_scale) / divisor; /* :3674-3675 */
        bits_out[count++] = fma(bits_per_residual, (double)(sample_count - o), (double)header_bits); /* :3677 */
    }
    return count;
}

/* compute_best_order  src/encode.rs:3688-3702: first minimum under total_cmp; 0 = NoBestLpcOrder */
static uint32_t compute_best_order(uint32_t bps, uint32_t precision, uint32_t sample_count, const double* errors,
                                   uint32_t n_orders)
{
    double bits[MAX_LPC];
    int cnt = fo_subframe_bits_by_order(bps, precision, sample_count, errors, n_orders, bits);
    if (cnt == 0) return 0;
    int best = 0;
    for (int i = 1; i < cnt; i++)
        if (total_cmp(bits[i], bits[best]) < 0) best = i;
    return (uint32_t)best + 1;
}

/* saturating f64 -> i32 cast (Rust `as i32`) */
static int32_t f64_as_i32(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

/* LpcParameters::quantize  src/encode.rs:3334-3401 */
int fo_quantize(uint32_t order, const double* coeffs, uint32_t precision, int32_t* qcoefs, uint32_t* shift_out)
{
    const int32_t MAX_SHIFT = (1 << 4) - 1, MIN_SHIFT = -(1 << 4);
    int32_t max_coeff = (1 << (precision - 1)) - 1;
    int32_t min_coeff = -(1 << (precision - 1));

    /* max_by(total_cmp) over |c|  :3350-3356 */
    double l = fabs(coeffs[0]);
    for (uint32_t i = 1; i < order; i++) {
        double a = fabs(coeffs[i]);
        if (total_cmp(a, l) >= 0) l = a;
    }
    if (!(l > 0.0)) return FO_ERR_ZERO_LP_COEFFICIENTS;

    double error = 0.0;
    int32_t lg = f64_as_i32(floor(log2(l)));
    int64_t sh64 = (int64_t)((int32_t)precision - 1) - (int64_t)lg - 1; /* :3360 */
    int32_t shift = sh64 > MAX_SHIFT ? MAX_SHIFT : (sh64 < INT32_MIN ? INT32_MIN : (int32_t)sh64);
    if (shift >= 0) {
        for (uint32_t i = 0; i < order; i++) {
            double sum = fma(coeffs[i], (double)(1 << shift), error); /* mul_add  :3372 */
            int32_t q = f64_as_i32(round(sum));
            q = q < min_coeff ? min_coeff : (q > max_coeff ? max_coeff : q);
            error = sum - (double)q;
            qcoefs[i] = q;
        }
        *shift_out = (uint32_t)shift;
        return 0;
    } else if (shift >= MIN_SHIFT) { /* :3380 */
        uint32_t s = (uint32_t)(-shift);
        for (uint32_t i = 0; i < order; i++) {
            double sum = (coeffs[i] / (double)(1 << s)) + error; /* :3391 */
            int32_t q = f64_as_i32(round(sum));
            q = q < min_coeff ? min_coeff : (q > max_coeff ? max_coeff : q);
            error = sum - (double)q;
            qcoefs[i] = q;
        }
        *shift_out = 0;
        return 0;
    }
    return FO_ERR_LP_NEGATIVE_SHIFT; /* :3399 */
}

/* LpcSubframeParameters::encode_residuals  src/encode.rs:3174-3203 */
int fo_lpc_residuals(uint32_t order, uint32_t shift, const int32_t* qcoefs, const int32_t* x, uint32_t n, int32_t* res)
{
    for (uint32_t i = order; i < n; i++) {
        int64_t sum = 0;
        for (uint32_t j = 0; j < order; j++) sum += (int64_t)x[i - 1 - j] * (int64_t)qcoefs[j]; /* :3187-3192 */
        int32_t pred = (int32_t)(uint32_t)(uint64_t)(sum >> shift); /* `as i32` truncates  :3193 */
        int64_t r = (int64_t)x[i] - (int64_t)pred;
        if (r < INT32_MIN || r > INT32_MAX) return FO_ERR_RESIDUAL_OVERFLOW; /* checked_sub  :3186 */
        res[i - order] = (int32_t)r;
    }
    return 0;
}

/* ceil(log2(sum / samples)) exactly as src/encode.rs:3778-3780 evaluates it */
uint32_t fo_rice_parameter_f64(uint64_t sum, uint32_t samples)
{
    double v = ceil(log2((double)sum / (double)samples));
    if (v != v || v <= 0.0) return 0;
    if (v >= 4294967295.0) return UINT32_MAX;
    return (uint32_t)v;
}

/* ------------------------------------------------------------------------------------------
 * write_residuals  src/encode.rs:3747-3962
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t kind; /* 0 Standard{rice}, 1 Escaped{escape_size}, 2 Constant */
    uint8_t param;
    uint32_t start, len; /* slice of the residual array */
} partition_t;

static inline uint32_t ilog2_u64(uint64_t v) { return 63u - (uint32_t)__builtin_clzll(v); }

/* Partition::new  :3765-3831.  Returns 0 when the reference returns None. */
static int partition_new(const int32_t* res, uint32_t start, uint32_t len, uint32_t rice_max, uint32_t* estimated_bits,
                         partition_t* out)
{
    uint32_t partition_samples = (uint32_t)(uint16_t)len; /* `as u16`  :3766 */
    if (partition_samples == 0) return 0;
    uint64_t sum = 0;
    for (uint32_t i = 0; i < len; i++) {
        int32_t r = res[start + i];
        sum += (uint64_t)(r < 0 ? 0u - (uint32_t)r : (uint32_t)r); /* unsigned_abs  :3773 */
    }
    out->start = start;
    out->len = len;
    if (sum > 0) {
        uint32_t rice;
        if (sum > partition_samples) { /* :3777 */
            uint32_t bits_needed = fo_rice_parameter_f64(sum, partition_samples);
            if (bits_needed < rice_max) { /* BitCount::try_from + filter  :3782-3784 */
                rice = bits_needed;
            } else {
                uint32_t escape_size = ilog2_u64(sum) + 2; /* :3787-3792 */
                if (escape_size > 31) return 0;            /* try_into::<SignedBitCount<31>>  :3793 */
                *estimated_bits += escape_size * partition_samples; /* :3796 */
                out->kind = 1;
                out->param = (uint8_t)escape_size;
                return 1;
            }
        } else {
            rice = 0; /* :3806 */
        }
        uint64_t t = rice > 0 ? (sum >> (rice - 1)) : (sum << 1); /* :3811-3815 */
        if (t > UINT32_MAX) return 0;                            /* u32::try_from(..).ok()? */
        uint32_t partition_size = 4u + ((1u + rice) * partition_samples) + (uint32_t)t - (partition_samples / 2u);
        *estimated_bits += partition_size; /* :3818 */
        out->kind = 0;
        out->param = (uint8_t)rice;
        return 1;
    }
    out->kind = 2; /* all residuals 0  :3826 */
    out->param = 0;
    return 1;
}

/* best_partitions  :3865-3896.  Returns the partition count. */
static uint32_t best_partitions(const fo_options* opt, uint32_t rice_max, uint32_t block_size, const int32_t* res,
                                uint32_t n_res, partition_t* best)
{
    uint32_t tz = block_size ? (uint32_t)__builtin_ctz(block_size) : 32;
    uint32_t max_p = tz < opt->max_partition_order ? tz : opt->max_partition_order; /* :3870 */
    int have = 0;
    uint32_t best_bits = 0, best_count = 0;
    partition_t cand[MAX_PARTITIONS];
    for (uint32_t p = 0; p <= max_p; p++) {
        uint32_t partition_count = 1u << p;
        uint32_t chunk = block_size / partition_count; /* rchunks(block_size / partition_count)  :3877 */
        if (chunk == 0) break;                         /* rchunks(0) panics in the reference; unreachable (p <= tz) */
        /* rchunks(..).rev(): chunks aligned to the END; the first one may be short */
        uint32_t count = (n_res + chunk - 1) / chunk;
        if (count > MAX_PARTITIONS) continue; /* ArrayVec overflow: the reference panics (Appendix A.16) */
        uint32_t estimated_bits = 0;
        int ok = 1;
        uint32_t first_len = n_res - (count - 1) * chunk;
        uint32_t pos = 0;
        for (uint32_t j = 0; j < count; j++) {
            uint32_t len = j == 0 ? first_len : chunk;
            if (!partition_new(res, pos, len, rice_max, &estimated_bits, &cand[j])) {
                ok = 0;
                break;
            }
            pos += len;
        }
        if (!ok) continue;                                    /* collect::<Option<..>>  :3880 */
        if (count == 0 || (count & (count - 1)) != 0) continue; /* !is_empty && is_power_of_two  :3881 */
        if (!have || estimated_bits < best_bits) {            /* min_by_key, first wins  :3885 */
            have = 1;
            best_bits = estimated_bits;
            best_count = count;
            memcpy(best, cand, sizeof(partition_t) * count);
        }
    }
    if (!have) { /* unwrap_or_else  :3887-3895 */
        best[0].kind = 1;
        best[0].param = 31;
        best[0].start = 0;
        best[0].len = n_res;
        return 1;
    }
    return best_count;
}

/* Partition::to_writer  :3834-3863 and ResidualPartitionHeader::to_writer  src/stream.rs:1603-1619 */
static void write_partition(bits_t* w, const partition_t* p, const int32_t* res, uint32_t rice_max)
{
    uint32_t hdr_bits = rice_max == 15 ? 4 : 5;
    if (p->kind == 0) {
        bits_put(w, hdr_bits, p->param);
        uint32_t k = p->param;
        uint32_t mask = k ? ((k >= 32 ? 0xffffffffu : (1u << k) - 1u)) : 0;
        for (uint32_t i = 0; i < p->len; i++) {
            int32_t s = res[p->start + i];
            /* :3845-3849, u32 arithmetic exactly as written */
            uint32_t u = s < 0 ? ((((uint32_t)(-(int64_t)s)) - 1u) << 1) + 1u : ((uint32_t)s) << 1;
            bits_put_unary1(w, u >> k);
            bits_put(w, k, u & mask);
        }
    } else if (p->kind == 1) {
        bits_put(w, hdr_bits, rice_max);
        bits_put(w, 5, p->param);
        for (uint32_t i = 0; i < p->len; i++) bits_put_signed(w, p->param, res[p->start + i]);
    } else {
        bits_put(w, hdr_bits, rice_max);
        bits_put(w, 5, 0);
    }
}

static void write_residuals(const fo_options* opt, int use_rice2, bits_t* w, uint32_t predictor_order, const int32_t* res,
                            uint32_t n_res, fo_subframe_info* info)
{
    partition_t parts[MAX_PARTITIONS];
    uint32_t block_size = predictor_order + n_res; /* :3944 */
    uint32_t method, rice_max;
    uint32_t count;
    if (use_rice2) { /* :3946 */
        count = best_partitions(opt, 31, block_size, res, n_res, parts);
        /* try_reduce_rice  :3929-3942 */
        int shrink = 1;
        for (uint32_t j = 0; j < count; j++)
            if (parts[j].kind == 0 && parts[j].param >= 15) shrink = 0;
        method = shrink ? 0 : 1;
    } else {
        count = best_partitions(opt, 15, block_size, res, n_res, parts);
        method = 0;
    }
    rice_max = method ? 31 : 15;
    bits_put(w, 2, method);
    bits_put(w, 4, 31u - (uint32_t)__builtin_clz(count)); /* partitions.len().ilog2()  :3902 */
    for (uint32_t j = 0; j < count; j++) write_partition(w, &parts[j], res, rice_max);
    if (info) {
        info->coding_method = (int32_t)method;
        info->partition_order = (int32_t)(31u - (uint32_t)__builtin_clz(count));
        for (uint32_t j = 0; j < count && j < 64; j++) {
            info->rice[j] = parts[j].param;
            info->kind[j] = parts[j].kind;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Subframes
 * ---------------------------------------------------------------------------------------- */
/* SubframeHeader::to_writer  src/stream.rs:1397-1413; type codes :1555-1567 */
static void write_subframe_header(bits_t* w, uint32_t type_code, uint32_t wasted_bps)
{
    bits_put(w, 1, 0);
    bits_put(w, 6, type_code);
    if (wasted_bps == 0) {
        bits_put(w, 1, 0);
    } else {
        bits_put(w, 1, 1);
        bits_put_unary1(w, wasted_bps - 1);
    }
}

typedef struct {
    bits_t fixed_out, lpc_out, other_out; /* fixed_output / lpc_output / constant_output+verbatim_output */
    int32_t* wasted;                      /* samples >> wasted_bps */
    int32_t* fixed_buf[4];
    int32_t* lpc_res;
    double* window;
    double* windowed;
    uint32_t window_len;
    fo_options window_opt;
    uint32_t cap;
} chan_cache;

struct fo_encoder {
    chan_cache ch[MAX_CHANNELS]; /* channels[..]; for stereo-exhaustive: left, right, average, difference */
    int32_t* average;
    int32_t* difference;
    uint32_t cap;
    bits_t frame;
};

static void cache_reserve(chan_cache* c, uint32_t n)
{
    if (n <= c->cap) return;
    c->wasted = (int32_t*)realloc(c->wasted, sizeof(int32_t) * n);
    for (int i = 0; i < 4; i++) c->fixed_buf[i] = (int32_t*)realloc(c->fixed_buf[i], sizeof(int32_t) * n);
    c->lpc_res = (int32_t*)realloc(c->lpc_res, sizeof(int32_t) * n);
    c->window = (double*)realloc(c->window, sizeof(double) * n);
    c->windowed = (double*)realloc(c->windowed, sizeof(double) * n);
    c->window_len = 0;
    c->cap = n;
}

fo_encoder* fo_encoder_new(void)
{
    init_tables();
    return (fo_encoder*)calloc(1, sizeof(fo_encoder));
}

void fo_encoder_free(fo_encoder* e)
{
    if (!e) return;
    for (int i = 0; i < MAX_CHANNELS; i++) {
        chan_cache* c = &e->ch[i];
        free(c->fixed_out.data);
        free(c->lpc_out.data);
        free(c->other_out.data);
        free(c->wasted);
        for (int k = 0; k < 4; k++) free(c->fixed_buf[k]);
        free(c->lpc_res);
        free(c->window);
        free(c->windowed);
    }
    free(e->average);
    free(e->difference);
    free(e->frame.data);
    free(e);
}

/* encode_constant_subframe  src/encode.rs:2982-2998 */
static void encode_constant_subframe(bits_t* w, int32_t sample, uint32_t bps, uint32_t wasted)
{
    write_subframe_header(w, 0, wasted);
    bits_put_signed(w, bps, sample);
}

/* encode_verbatim_subframe  src/encode.rs:3000-3018 */
static void encode_verbatim_subframe(bits_t* w, const int32_t* ch, uint32_t n, uint32_t bps, uint32_t wasted)
{
    write_subframe_header(w, 1, wasted);
    for (uint32_t i = 0; i < n; i++) bits_put_signed(w, bps, ch[i]);
}

/* encode_fixed_subframe  src/encode.rs:3020-3088.  Returns 0 or an error code. */
static int encode_fixed_subframe(const fo_options* opt, int use_rice2, chan_cache* c, bits_t* w, const int32_t* ch,
                                 uint32_t n, uint32_t bps, uint32_t wasted, fo_subframe_info* info)
{
    const int32_t* orders[5];
    uint32_t lens[5];
    int n_orders = 1;
    orders[0] = ch;
    lens[0] = n;
    for (int b = 0; b < 4; b++) { /* :3039-3060 */
        const int32_t* prev = orders[n_orders - 1];
        uint32_t plen = lens[n_orders - 1];
        if (plen < 1) break; /* split_at_checked(1) == None */
        int32_t* buf = c->fixed_buf[b];
        uint32_t blen = 0;
        int overflow = 0;
        for (uint32_t i = 1; i < plen; i++) {
            int64_t v = (int64_t)prev[i] - (int64_t)prev[i - 1]; /* checked_sub  :3045 */
            if (v < INT32_MIN || v > INT32_MAX) {
                overflow = 1;
                break;
            }
            buf[blen++] = (int32_t)v;
        }
        if (overflow) break;   /* break 'outer */
        if (blen == 0) break;  /* buf.is_empty() */
        orders[n_orders] = buf;
        lens[n_orders] = blen;
        n_orders++;
    }
    uint32_t min_fixed = lens[n_orders - 1]; /* :3062 */
    int best = 0;
    uint64_t best_sum = 0;
    for (int o = 0; o < n_orders; o++) { /* :3065-3075, first minimum */
        uint64_t s = 0;
        const int32_t* r = orders[o] + (lens[o] - min_fixed);
        for (uint32_t i = 0; i < min_fixed; i++) s += (uint64_t)(r[i] < 0 ? 0u - (uint32_t)r[i] : (uint32_t)r[i]);
        if (o == 0 || s < best_sum) {
            best = o;
            best_sum = s;
        }
    }
    write_subframe_header(w, 8u + (uint32_t)best, wasted); /* Fixed{order}  :3078 */
    for (int i = 0; i < best; i++) bits_put_signed(w, bps, ch[i]); /* warm-up  :3083 */
    if (info) {
        info->type = 2;
        info->order = best;
    }
    write_residuals(opt, use_rice2, w, (uint32_t)best, orders[best], lens[best], info); /* :3087 */
    return w->err;
}

/* LpcParameters::best  src/encode.rs:3292-3332 */
static int lpc_parameters_best(const fo_options* opt, chan_cache* c, const int32_t* ch, uint32_t n, uint32_t bps,
                               uint32_t max_lpc_order, uint32_t* order_out, uint32_t* precision_out, uint32_t* shift_out,
                               int32_t* qcoefs)
{
    if (n <= max_lpc_order) return FO_ERR_INSUFFICIENT_LPC_SAMPLES; /* :3300 */
    uint32_t precision; /* :3305-3315 */
    if (n <= 192) precision = 7;
    else if (n <= 384) precision = 8;
    else if (n <= 576) precision = 9;
    else if (n <= 1152) precision = 10;
    else if (n <= 2304) precision = 11;
    else if (n <= 4608) precision = 12;
    else precision = 13;

    /* Window::apply  :1785-1801 (window regenerated when the length changes) */
    if (c->window_len != n || memcmp(&c->window_opt, opt, sizeof(*opt)) != 0) {
        fo_window(opt, n, c->window);
        c->window_len = n;
        c->window_opt = *opt;
    }
    for (uint32_t i = 0; i < n; i++) c->windowed[i] = (double)ch[i] * c->window[i]; /* :1799 */

    double autoc[MAX_LPC + 1];
    int n_autoc = fo_autocorrelate(c->windowed, n, max_lpc_order, autoc);
    double coeffs[MAX_LPC * MAX_LPC];
    double errors[MAX_LPC];
    int n_orders = fo_lp_coefficients(autoc, (uint32_t)n_autoc, coeffs, errors);
    uint32_t order = compute_best_order(bps, precision, n, errors, (uint32_t)n_orders);
    if (order == 0) return FO_ERR_NO_BEST_LPC_ORDER;
    int rc = fo_quantize(order, coeffs + (size_t)(order - 1) * MAX_LPC, precision, qcoefs, shift_out);
    if (rc) return rc;
    *order_out = order;
    *precision_out = precision;
    return 0;
}

/* encode_lpc_subframe  src/encode.rs:3090-3136 */
static int encode_lpc_subframe(const fo_options* opt, int use_rice2, chan_cache* c, bits_t* w, const int32_t* ch,
                               uint32_t n, uint32_t bps, uint32_t wasted, fo_subframe_info* info)
{
    uint32_t order, precision, shift;
    int32_t q[MAX_LPC];
    int rc = lpc_parameters_best(opt, c, ch, n, bps, opt->max_lpc_order, &order, &precision, &shift, q);
    if (rc) return rc;
    rc = fo_lpc_residuals(order, shift, q, ch, n, c->lpc_res); /* :3165 */
    if (rc) return rc;
    write_subframe_header(w, order + 31u, wasted); /* Lpc{order}  :3113 */
    for (uint32_t i = 0; i < order; i++) bits_put_signed(w, bps, ch[i]); /* :3118 */
    bits_put(w, 4, precision - 1);                                       /* :3122 */
    bits_put(w, 5, shift);                                               /* :3129 */
    for (uint32_t i = 0; i < order; i++) bits_put_signed(w, precision, q[i]); /* :3131 */
    if (info) {
        info->type = 3;
        info->order = (int32_t)order;
        info->precision = (int32_t)precision;
        info->shift = (int32_t)shift;
        memcpy(info->coefs, q, sizeof(int32_t) * order);
    }
    write_residuals(opt, use_rice2, w, order, c->lpc_res, n - order, info); /* :3135 */
    return w->err;
}

/* encode_subframe  src/encode.rs:2849-2980.  Returns the chosen recorder or NULL + *err. */
static const bits_t* encode_subframe(const fo_options* opt, int use_rice2, chan_cache* c, const int32_t* channel, uint32_t n,
                                     uint32_t bps, int all_0, int* err, fo_subframe_info* info)
{
    fo_subframe_info tmp_fixed, tmp_lpc;
    memset(&tmp_fixed, 0, sizeof(tmp_fixed));
    memset(&tmp_lpc, 0, sizeof(tmp_lpc));
    cache_reserve(c, n);
    *err = 0;
    if (info) memset(info, 0, sizeof(*info));
    if (all_0) { /* :2870 */
        bits_clear(&c->other_out);
        encode_constant_subframe(&c->other_out, channel[0], bps, 0);
        if (info) {
            info->type = 0;
            info->bps = (int32_t)bps;
            info->bits = c->other_out.nbits;
        }
        return &c->other_out;
    }
    /* wasted bits  :2878-2898 */
    uint32_t wasted_bps = 0;
    {
        uint32_t acc = 32;
        int none = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t tz = channel[i] == 0 ? 32u : (uint32_t)__builtin_ctz((uint32_t)channel[i]);
            if (tz == 0) { /* NonZero::new(0) == None aborts the fold */
                none = 1;
                break;
            }
            if (tz < acc) acc = tz;
        }
        if (!none) {
            if (acc == 32) { /* Some(WASTED_MAX)  :2883 */
                bits_clear(&c->other_out);
                encode_constant_subframe(&c->other_out, channel[0], bps, 0);
                if (info) {
                    info->type = 0;
                    info->bps = (int32_t)bps;
                    info->bits = c->other_out.nbits;
                }
                return &c->other_out;
            }
            wasted_bps = acc;
            for (uint32_t i = 0; i < n; i++) c->wasted[i] = channel[i] >> wasted_bps; /* :2891 */
            channel = c->wasted;
            bps = bps - wasted_bps; /* checked_sub(..).unwrap()  :2894 */
        }
    }

    bits_clear(&c->fixed_out);
    const bits_t* best = NULL;
    int best_is_lpc = 0;
    if (opt->max_lpc_order) { /* :2902 */
        bits_clear(&c->lpc_out);
        int rf = encode_fixed_subframe(opt, use_rice2, c, &c->fixed_out, channel, n, bps, wasted_bps, &tmp_fixed);
        int rl = encode_lpc_subframe(opt, use_rice2, c, &c->lpc_out, channel, n, bps, wasted_bps, &tmp_lpc);
        if (!rf && !rl) { /* min_by_key(written), fixed first  :2929-2932 */
            if (c->lpc_out.nbits < c->fixed_out.nbits) {
                best = &c->lpc_out;
                best_is_lpc = 1;
            } else {
                best = &c->fixed_out;
            }
        } else if (rf && !rl) {
            best = &c->lpc_out;
            best_is_lpc = 1;
        } else if (!rf && rl) {
            best = &c->fixed_out;
        }
    } else {
        int rf = encode_fixed_subframe(opt, use_rice2, c, &c->fixed_out, channel, n, bps, wasted_bps, &tmp_fixed);
        if (!rf) best = &c->fixed_out;
    }
    uint32_t verbatim_len = n * bps; /* u32  :2971 */
    if (best && best->nbits < (uint64_t)verbatim_len) {
        if (info) {
            *info = best_is_lpc ? tmp_lpc : tmp_fixed;
            info->wasted = (int32_t)wasted_bps;
            info->bps = (int32_t)bps;
            info->bits = best->nbits;
        }
        return best;
    }
    bits_clear(&c->other_out);
    encode_verbatim_subframe(&c->other_out, channel, n, bps, wasted_bps); /* :2936 / :2976 */
    if (c->other_out.err) {
        *err = c->other_out.err;
        return NULL;
    }
    if (info) {
        info->type = 1;
        info->wasted = (int32_t)wasted_bps;
        info->bps = (int32_t)bps;
        info->bits = c->other_out.nbits;
    }
    return &c->other_out;
}

/* ------------------------------------------------------------------------------------------
 * Frame header  src/stream.rs:242-276 (build), :537-560 (block size), :779-802 (sample rate),
 * :1136-1149 (bps), :1266-1326 (frame number)
 * ---------------------------------------------------------------------------------------- */
int fo_write_frame_number(uint64_t v, uint8_t out[7])
{
    if (v <= 0x7F) {
        out[0] = (uint8_t)v;
        return 1;
    }
    int bytes;
    if (v <= 0x7FF) bytes = 2;
    else if (v <= 0xFFFF) bytes = 3;
    else if (v <= 0x1FFFFF) bytes = 4;
    else if (v <= 0x3FFFFFF) bytes = 5;
    else if (v <= 0x7FFFFFFFull) bytes = 6;
    else if (v <= 0xFFFFFFFFFull) bytes = 7;
    else return -FO_ERR_INVALID_FRAME_NUMBER;
    /* write_unary::<0>(bytes) then (7 - bytes) bits of the top, then continuation bytes */
    uint32_t lead_bits = 7 - (uint32_t)bytes;
    uint8_t prefix = (uint8_t)(0xFFu << (8 - bytes));
    out[0] = (uint8_t)(prefix | (lead_bits ? (uint8_t)((v >> (6 * (bytes - 1))) & ((1u << lead_bits) - 1u)) : 0));
    for (int i = 1; i < bytes; i++) out[i] = (uint8_t)(0x80 | ((v >> (6 * (bytes - 1 - i))) & 0x3F));
    return bytes;
}

int fo_read_frame_number(const uint8_t* p, size_t n, uint64_t* v)
{
    if (n < 1) return -FO_ERR_IO;
    uint8_t b0 = p[0];
    int ones = 0;
    while (ones < 8 && (b0 & (0x80 >> ones))) ones++;
    if (ones == 0) {
        *v = b0 & 0x7F;
        return 1;
    }
    if (ones == 1 || ones > 7) return -FO_ERR_INVALID_FRAME_NUMBER; /* src/stream.rs:1250, :1260 */
    uint64_t frame = ones < 7 ? (uint64_t)(b0 & ((1u << (7 - ones)) - 1u)) : 0;
    if ((size_t)ones > n) return -FO_ERR_IO;
    for (int i = 1; i < ones; i++) {
        if ((p[i] & 0xC0) != 0x80) return -FO_ERR_INVALID_FRAME_NUMBER;
        frame = (frame << 6) | (p[i] & 0x3F);
    }
    *v = frame;
    return ones;
}

static int block_size_code(uint32_t bs, uint32_t* extra_bits)
{
    *extra_bits = 0;
    switch (bs) {
    case 192: return 1;
    case 576: return 2;
    case 1152: return 3;
    case 2304: return 4;
    case 4608: return 5;
    case 256: return 8;
    case 512: return 9;
    case 1024: return 10;
    case 2048: return 11;
    case 4096: return 12;
    case 8192: return 13;
    case 16384: return 14;
    case 32768: return 15;
    default: break;
    }
    if (bs <= 256) { /* Uncommon8  src/stream.rs:557 */
        *extra_bits = 8;
        return 6;
    }
    *extra_bits = 16;
    return 7;
}

/* returns the 4-bit code; kind: 0 none, 1 kHz (8 bits), 2 Hz (16 bits), 3 daHz (16 bits); -1 invalid */
static int sample_rate_code(uint32_t rate, int* kind)
{
    *kind = 0;
    switch (rate) {
    case 88200: return 1;
    case 176400: return 2;
    case 192000: return 3;
    case 8000: return 4;
    case 16000: return 5;
    case 22050: return 6;
    case 24000: return 7;
    case 32000: return 8;
    case 44100: return 9;
    case 48000: return 10;
    case 96000: return 11;
    default: break;
    }
    if (rate % 1000 == 0 && rate / 1000 < 255) { /* src/stream.rs:796 */
        *kind = 1;
        return 12;
    }
    if (rate % 10 == 0 && rate / 10 < 65535) {
        *kind = 3;
        return 14;
    }
    if (rate < 65535) {
        *kind = 2;
        return 13;
    }
    if (rate < (1u << 20)) return 0; /* Streaminfo(rate) */
    return -1;
}

static int bps_code(uint32_t bps)
{
    switch (bps) {
    case 8: return 1;
    case 12: return 2;
    case 16: return 4;
    case 20: return 5;
    case 24: return 6;
    case 32: return 7;
    default: return 0; /* Streaminfo  src/stream.rs:1147 */
    }
}

/* FrameHeader::write / write_subset: header bytes incl. CRC-8.  Returns length or negative error. */
static int write_frame_header(uint8_t* h, uint32_t block_size, uint32_t sample_rate, uint32_t bps, uint32_t assignment,
                              uint64_t frame_number, int subset)
{
    uint32_t bs_extra;
    int rate_kind;
    if (block_size == 0 || block_size > 65535) return -FO_ERR_INVALID_BLOCK_SIZE;
    int bsc = block_size_code(block_size, &bs_extra);
    int src = sample_rate_code(sample_rate, &rate_kind);
    if (src < 0) return -FO_ERR_INVALID_SAMPLE_RATE;
    int bpc = bps_code(bps);
    if (subset && src == 0) return -FO_ERR_NON_SUBSET_SAMPLE_RATE; /* src/encode.rs:1128-1131 */
    if (subset && bpc == 0) return -FO_ERR_NON_SUBSET_BPS;         /* :1134-1137 */
    int n = 0;
    h[n++] = 0xFF;                                           /* sync 111111111111100 + blocking strategy 0 */
    h[n++] = 0xF8;
    h[n++] = (uint8_t)((bsc << 4) | src);
    h[n++] = (uint8_t)((assignment << 4) | (bpc << 1));
    int fl = fo_write_frame_number(frame_number, h + n);
    if (fl < 0) return fl;
    n += fl;
    if (bs_extra == 8) h[n++] = (uint8_t)(block_size - 1);
    else if (bs_extra == 16) {
        h[n++] = (uint8_t)((block_size - 1) >> 8);
        h[n++] = (uint8_t)(block_size - 1);
    }
    if (rate_kind == 1) h[n++] = (uint8_t)(sample_rate / 1000);
    else if (rate_kind == 2) {
        h[n++] = (uint8_t)(sample_rate >> 8);
        h[n++] = (uint8_t)sample_rate;
    } else if (rate_kind == 3) {
        h[n++] = (uint8_t)((sample_rate / 10) >> 8);
        h[n++] = (uint8_t)(sample_rate / 10);
    }
    h[n] = fo_crc8(h, (size_t)n);
    return n + 1;
}

/* ------------------------------------------------------------------------------------------
 * encode_frame  src/encode.rs:2259-2439, correlate_channels :2463-2674,
 * correlate_channels_exhaustive :2676-2847
 * ---------------------------------------------------------------------------------------- */
static void enc_reserve(fo_encoder* e, uint32_t n)
{
    if (n <= e->cap) return;
    e->average = (int32_t*)realloc(e->average, sizeof(int32_t) * n);
    e->difference = (int32_t*)realloc(e->difference, sizeof(int32_t) * n);
    e->cap = n;
}

static inline uint64_t uabs64(int32_t v) { return (uint64_t)(v < 0 ? 0u - (uint32_t)v : (uint32_t)v); }

int64_t fo_encode_frame(fo_encoder* e, const fo_options* opt, uint32_t sample_rate, uint32_t bps, uint32_t channels,
                        uint64_t frame_number, const int32_t* const* planar, uint32_t n, int subset, uint8_t* out,
                        size_t out_cap, fo_frame_info* info)
{
    if (channels < 1 || channels > 8) return -FO_ERR_EXCESSIVE_CHANNELS;
    if (n == 0) return -FO_ERR_INVALID_BLOCK_SIZE;
    int use_rice2 = bps > 16; /* src/encode.rs:1965, :1115 */
    enc_reserve(e, n);
    const bits_t* sub[MAX_CHANNELS];
    uint32_t assignment;
    int err = 0;
    if (info) {
        memset(info, 0, sizeof(*info));
        info->channels = (int32_t)channels;
    }
    fo_subframe_info* si[4] = {NULL, NULL, NULL, NULL};
    fo_subframe_info infos4[4];

    if (channels == 1) { /* :2283 */
        int all0 = 1;
        for (uint32_t i = 0; i < n; i++)
            if (planar[0][i] != 0) {
                all0 = 0;
                break;
            }
        sub[0] = encode_subframe(opt, use_rice2, &e->ch[0], planar[0], n, bps, all0, &err, info ? &info->sub[0] : NULL);
        if (!sub[0]) return -err;
        assignment = 0;
    } else if (channels == 2 && opt->exhaustive_channel_correlation) { /* :2307, :2676 */
        const int32_t *left = planar[0], *right = planar[1];
        for (int k = 0; k < 4; k++) si[k] = info ? &infos4[k] : NULL;
        const bits_t* lrec = encode_subframe(opt, use_rice2, &e->ch[0], left, n, bps, 0, &err, si[0]);
        if (!lrec) return -err;
        const bits_t* rrec = encode_subframe(opt, use_rice2, &e->ch[1], right, n, bps, 0, &err, si[1]);
        if (!rrec) return -err;
        int pick[2] = {0, 1};
        assignment = 1;
        if (bps + 1 <= 32 && opt->mid_side) { /* :2716 */
            for (uint32_t i = 0; i < n; i++) e->average[i] = (left[i] + right[i]) >> 1; /* :2721 */
            const bits_t* arec = encode_subframe(opt, use_rice2, &e->ch[2], e->average, n, bps, 0, &err, si[2]);
            if (!arec) return -err;
            for (uint32_t i = 0; i < n; i++) e->difference[i] = left[i] - right[i]; /* :2734 */
            const bits_t* drec = encode_subframe(opt, use_rice2, &e->ch[3], e->difference, n, bps + 1, 0, &err, si[3]);
            if (!drec) return -err;
            /* [Independent, LeftSide, SideRight, MidSide], first minimum  :2747-2768 */
            uint64_t tot[4] = {lrec->nbits + rrec->nbits, lrec->nbits + drec->nbits, drec->nbits + rrec->nbits,
                               arec->nbits + drec->nbits};
            int b = 0;
            for (int k = 1; k < 4; k++)
                if (tot[k] < tot[b]) b = k;
            if (b == 0) { assignment = 1; sub[0] = lrec; sub[1] = rrec; pick[0] = 0; pick[1] = 1; }
            else if (b == 1) { assignment = 8; sub[0] = lrec; sub[1] = drec; pick[0] = 0; pick[1] = 3; }
            else if (b == 2) { assignment = 9; sub[0] = drec; sub[1] = rrec; pick[0] = 3; pick[1] = 1; }
            else { assignment = 10; sub[0] = arec; sub[1] = drec; pick[0] = 2; pick[1] = 3; }
        } else if (bps + 1 <= 32) { /* :2788 */
            for (uint32_t i = 0; i < n; i++) e->difference[i] = left[i] - right[i];
            const bits_t* drec = encode_subframe(opt, use_rice2, &e->ch[3], e->difference, n, bps + 1, 0, &err, si[3]);
            if (!drec) return -err;
            uint64_t tot[3] = {lrec->nbits + rrec->nbits, lrec->nbits + drec->nbits, drec->nbits + rrec->nbits}; /* :2803 */
            int b = 0;
            for (int k = 1; k < 3; k++)
                if (tot[k] < tot[b]) b = k;
            if (b == 0) { assignment = 1; sub[0] = lrec; sub[1] = rrec; pick[0] = 0; pick[1] = 1; }
            else if (b == 1) { assignment = 8; sub[0] = lrec; sub[1] = drec; pick[0] = 0; pick[1] = 3; }
            else { assignment = 9; sub[0] = drec; sub[1] = rrec; pick[0] = 3; pick[1] = 1; }
        } else { /* 32 bps: independent only  :2837 */
            sub[0] = lrec;
            sub[1] = rrec;
        }
        if (info) {
            info->sub[0] = infos4[pick[0]];
            info->sub[1] = infos4[pick[1]];
        }
    } else if (channels == 2) { /* correlate_channels  :2335, :2463 */
        const int32_t *left = planar[0], *right = planar[1];
        const int32_t* c0 = left;
        const int32_t* c1 = right;
        uint32_t b0 = bps, b1 = bps;
        int a0, a1;
        assignment = 1;
        if (bps + 1 <= 32 && opt->mid_side) {
            uint64_t ls = 0, rs = 0, ms = 0, ss = 0;
            for (uint32_t i = 0; i < n; i++) {
                ls += uabs64(left[i]);
                rs += uabs64(right[i]);
                e->average[i] = (left[i] + right[i]) >> 1;
                ms += uabs64(e->average[i]);
                e->difference[i] = left[i] - right[i];
                ss += uabs64(e->difference[i]);
            }
            /* [Independent, LeftSide, SideRight, MidSide]  :2506-2517 */
            uint64_t tot[4] = {ls + rs, ls + ss, ss + rs, ms + ss};
            int b = 0;
            for (int k = 1; k < 4; k++)
                if (tot[k] < tot[b]) b = k;
            if (b == 0) { assignment = 1; a0 = ls == 0; a1 = rs == 0; }
            else if (b == 1) { assignment = 8; c1 = e->difference; b1 = bps + 1; a0 = ls == 0; a1 = ss == 0; }
            else if (b == 2) { assignment = 9; c0 = e->difference; b0 = bps + 1; a0 = ss == 0; a1 = rs == 0; }
            else { assignment = 10; c0 = e->average; c1 = e->difference; b1 = bps + 1; a0 = ms == 0; a1 = ss == 0; }
        } else if (bps + 1 <= 32) {
            uint64_t ls = 0, rs = 0, ss = 0;
            for (uint32_t i = 0; i < n; i++) {
                ls += uabs64(left[i]);
                rs += uabs64(right[i]);
                e->difference[i] = left[i] - right[i];
                ss += uabs64(e->difference[i]);
            }
            /* [LeftSide, SideRight, Independent]  :2600-2607 */
            uint64_t tot[3] = {ls + ss, ss + rs, ls + rs};
            int b = 0;
            for (int k = 1; k < 3; k++)
                if (tot[k] < tot[b]) b = k;
            if (b == 0) { assignment = 8; c1 = e->difference; b1 = bps + 1; a0 = ls == 0; a1 = ss == 0; }
            else if (b == 1) { assignment = 9; c0 = e->difference; b0 = bps + 1; a0 = ss == 0; a1 = rs == 0; }
            else { assignment = 1; a0 = ls == 0; a1 = rs == 0; }
        } else {
            a0 = 1;
            a1 = 1;
            for (uint32_t i = 0; i < n; i++) {
                if (left[i]) a0 = 0;
                if (right[i]) a1 = 0;
            }
        }
        sub[0] = encode_subframe(opt, use_rice2, &e->ch[0], c0, n, b0, a0, &err, info ? &info->sub[0] : NULL);
        if (!sub[0]) return -err;
        sub[1] = encode_subframe(opt, use_rice2, &e->ch[1], c1, n, b1, a1, &err, info ? &info->sub[1] : NULL);
        if (!sub[1]) return -err;
    } else { /* :2370 */
        for (uint32_t c = 0; c < channels; c++) {
            int all0 = 1;
            for (uint32_t i = 0; i < n; i++)
                if (planar[c][i] != 0) {
                    all0 = 0;
                    break;
                }
            sub[c] = encode_subframe(opt, use_rice2, &e->ch[c], planar[c], n, bps, all0, &err, info ? &info->sub[c] : NULL);
            if (!sub[c]) return -err;
        }
        assignment = channels - 1;
    }

    uint8_t hdr[24];
    int hl = write_frame_header(hdr, n, sample_rate, bps, assignment, frame_number, subset);
    if (hl < 0) return hl;
    bits_t* w = &e->frame;
    bits_clear(w);
    for (int i = 0; i < hl; i++) bits_put(w, 8, hdr[i]);
    for (uint32_t c = 0; c < channels; c++) bits_append(w, sub[c]); /* playback  :2332 */
    if (w->err) return -w->err;
    bits_align(w); /* aligned_writer  :2408 */
    size_t nbytes = (size_t)(w->nbits >> 3);
    uint16_t crc = fo_crc16(w->data, nbytes);
    bits_put(w, 16, crc); /* :2409 */
    nbytes += 2;
    if (nbytes > out_cap) return -FO_ERR_IO;
    memcpy(out, w->data, nbytes);
    if (info) {
        info->channel_assignment = (int32_t)assignment;
        info->frame_bytes = (uint32_t)nbytes;
    }
    return (int64_t)nbytes;
}

/* ------------------------------------------------------------------------------------------
 * PCM bytes <-> samples  src/audio.rs:110-187, src/byteorder.rs:48-186
 * ---------------------------------------------------------------------------------------- */
void fo_bytes_to_samples(const uint8_t* b, size_t n, uint32_t bytes_per_sample, int big_endian, int32_t* out)
{
    for (size_t i = 0; i < n; i++) {
        const uint8_t* p = b + i * bytes_per_sample;
        uint32_t v = 0;
        if (big_endian)
            for (uint32_t k = 0; k < bytes_per_sample; k++) v = (v << 8) | p[k];
        else
            for (uint32_t k = 0; k < bytes_per_sample; k++) v |= (uint32_t)p[k] << (8 * k);
        uint32_t sh = 32 - 8 * bytes_per_sample;
        out[i] = (int32_t)(v << sh) >> sh;
    }
}

void fo_samples_to_bytes(const int32_t* s, size_t n, uint32_t bytes_per_sample, int big_endian, uint8_t* out)
{
    for (size_t i = 0; i < n; i++) {
        uint32_t v = (uint32_t)s[i];
        uint8_t* p = out + i * bytes_per_sample;
        for (uint32_t k = 0; k < bytes_per_sample; k++) {
            uint8_t byte = (uint8_t)(v >> (8 * k));
            if (big_endian) p[bytes_per_sample - 1 - k] = byte;
            else p[k] = byte;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Stream level: Encoder::new / encode / finalize_inner  src/encode.rs:1882-2110,
 * write_blocks src/metadata/mod.rs:904-976, STREAMINFO :1742-1760, SEEKTABLE :2118-2139
 * ---------------------------------------------------------------------------------------- */
static void put_be(uint8_t* p, uint64_t v, int bytes)
{
    for (int i = 0; i < bytes; i++) p[i] = (uint8_t)(v >> (8 * (bytes - 1 - i)));
}

static void write_streaminfo_body(uint8_t* p, uint32_t block_size, uint32_t min_frame, uint32_t max_frame,
                                  uint32_t sample_rate, uint32_t channels, uint32_t bps, uint64_t total,
                                  const uint8_t md5[16])
{
    put_be(p, block_size, 2);
    put_be(p + 2, block_size, 2);
    put_be(p + 4, min_frame, 3);
    put_be(p + 7, max_frame, 3);
    /* 20 bits rate, 3 bits channels-1, 5 bits bps-1, 36 bits total */
    uint64_t v = ((uint64_t)sample_rate << 44) | ((uint64_t)(channels - 1) << 41) | ((uint64_t)(bps - 1) << 36) | (total & 0xFFFFFFFFFull);
    put_be(p + 10, v, 8);
    memcpy(p + 18, md5, 16);
}

typedef struct {
    uint64_t sample_offset, byte_offset;
    uint16_t frame_samples;
    int defined;
} seekpoint_t;

int64_t fo_encode_frames_only(const fo_options* opt, uint32_t sample_rate, uint32_t bps, uint32_t channels,
                              const int32_t* interleaved, uint64_t n_pcm_frames, uint64_t first_frame_number, int nthreads,
                              uint8_t* out, size_t out_cap, uint32_t* frame_sizes, size_t frame_sizes_cap,
                              uint64_t* n_frames_out, fo_frame_info* infos)
{
    uint32_t bs = opt->block_size;
    uint64_t n_frames = (n_pcm_frames + bs - 1) / bs;
    if (n_frames_out) *n_frames_out = n_frames;
    if (n_frames == 0) return 0;
    /* worst-case bytes of one frame: verbatim + headers */
    size_t frame_cap = (size_t)bs * channels * 5 + 256;
    uint8_t* tmp = (uint8_t*)malloc(frame_cap * n_frames);
    int64_t* sizes = (int64_t*)malloc(sizeof(int64_t) * n_frames);
    if (!tmp || !sizes) {
        free(tmp);
        free(sizes);
        return -FO_ERR_IO;
    }
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        fo_encoder* e = fo_encoder_new();
        int32_t* planes = (int32_t*)malloc(sizeof(int32_t) * (size_t)bs * channels);
        const int32_t* ptrs[MAX_CHANNELS];
#pragma omp for schedule(dynamic, 4)
        for (int64_t f = 0; f < (int64_t)n_frames; f++) {
            uint64_t start = (uint64_t)f * bs;
            uint32_t n = (uint32_t)((n_pcm_frames - start) < bs ? (n_pcm_frames - start) : bs);
            /* Frame::fill_from_samples de-interleave  src/audio.rs:190-203 */
            for (uint32_t c = 0; c < channels; c++) {
                int32_t* dst = planes + (size_t)c * bs;
                const int32_t* src = interleaved + start * channels + c;
                for (uint32_t i = 0; i < n; i++) dst[i] = src[(size_t)i * channels];
                ptrs[c] = dst;
            }
            sizes[f] = fo_encode_frame(e, opt, sample_rate, bps, channels, first_frame_number + (uint64_t)f, ptrs, n, 0,
                                       tmp + (size_t)f * frame_cap, frame_cap, infos ? &infos[f] : NULL);
        }
        free(planes);
        fo_encoder_free(e);
    }
    int64_t total = 0;
    for (uint64_t f = 0; f < n_frames; f++) {
        if (sizes[f] < 0) {
            total = sizes[f];
            break;
        }
        if ((size_t)(total + sizes[f]) > out_cap) {
            total = -FO_ERR_IO;
            break;
        }
        memcpy(out + total, tmp + (size_t)f * frame_cap, (size_t)sizes[f]);
        if (frame_sizes && f < frame_sizes_cap) frame_sizes[f] = (uint32_t)sizes[f];
        total += sizes[f];
    }
    free(tmp);
    free(sizes);
    return total;
}

/* SeekTableInterval::filter  src/encode.rs:1338-1358 applied to `pts`, writes kept indices */
static size_t seek_filter(const fo_options* opt, uint32_t sample_rate, const seekpoint_t* pts, size_t n, size_t* keep)
{
    size_t k = 0;
    if (opt->seektable_kind == 1) {
        uint64_t nth = (uint64_t)((uint32_t)(uint8_t)opt->seektable_n * sample_rate); /* u32 multiply  :1345 */
        uint64_t offset = 0;
        for (size_t i = 0; i < n; i++) {
            if (offset >= pts[i].sample_offset && offset < pts[i].sample_offset + pts[i].frame_samples) {
                offset += nth;
                keep[k++] = i;
            }
        }
    } else if (opt->seektable_kind == 2) {
        size_t step = opt->seektable_n ? opt->seektable_n : 1;
        for (size_t i = 0; i < n; i += step) keep[k++] = i;
    }
    return k;
}

int64_t fo_encode_stream(const fo_options* opt, uint32_t sample_rate, uint32_t bps, uint32_t channels,
                         const int32_t* interleaved, uint64_t n_pcm_frames, int total_known, int nthreads, uint8_t* out,
                         size_t out_cap, uint32_t* frame_sizes, size_t frame_sizes_cap, uint64_t* n_frames_out)
{
    const size_t MAX_POINTS = (1u << 24) / 18; /* src/metadata/mod.rs:1989 */
    if (sample_rate >= 1048576) return -FO_ERR_INVALID_SAMPLE_RATE; /* :1899-1902 */
    if (channels < 1 || channels > 8) return -FO_ERR_EXCESSIVE_CHANNELS;
    if (bps < 1 || bps > 32) return -FO_ERR_INVALID_BPS;
    if (n_pcm_frames == 0) return -FO_ERR_NO_SAMPLES;
    uint32_t bs = opt->block_size;
    uint64_t n_frames = (n_pcm_frames + bs - 1) / bs;

    /* placeholder SEEKTABLE sized from total_samples  :1920-1939 */
    size_t n_seek_slots = 0;
    seekpoint_t* pts = (seekpoint_t*)malloc(sizeof(seekpoint_t) * (n_frames + 1));
    size_t* keep = (size_t*)malloc(sizeof(size_t) * (n_frames + 1));
    int have_seektable = 0;
    if (total_known && opt->seektable_kind) {
        for (uint64_t f = 0; f < n_frames; f++) { /* EncoderSeekPoint::placeholders  :2131 */
            pts[f].sample_offset = f * bs;
            uint64_t rem = n_pcm_frames - f * bs;
            pts[f].frame_samples = (uint16_t)(rem < bs ? rem : bs);
            pts[f].defined = 0;
        }
        n_seek_slots = seek_filter(opt, sample_rate, pts, n_frames, keep);
        if (n_seek_slots > MAX_POINTS) n_seek_slots = MAX_POINTS;
        have_seektable = 1;
    }
    /* header layout: fLaC, STREAMINFO, [SEEKTABLE], [PADDING]  (sort order :1944-1951) */
    size_t seek_bytes = have_seektable ? 4 + 18 * n_seek_slots : 0;
    size_t pad_bytes = opt->padding >= 0 ? 4 + (size_t)opt->padding : 0;
    size_t header_len = 4 + 4 + 34 + seek_bytes + pad_bytes;
    if (header_len > out_cap) {
        free(pts);
        free(keep);
        return -FO_ERR_IO;
    }
    uint32_t* sizes = frame_sizes;
    uint32_t* own_sizes = NULL;
    if (!sizes || frame_sizes_cap < n_frames) {
        own_sizes = (uint32_t*)malloc(sizeof(uint32_t) * n_frames);
        sizes = own_sizes;
    }
    uint64_t nf = 0;
    int64_t frames_len = fo_encode_frames_only(opt, sample_rate, bps, channels, interleaved, n_pcm_frames, 0, nthreads,
                                               out + header_len, out_cap - header_len, sizes, n_frames, &nf, NULL);
    if (frames_len < 0) {
        free(pts);
        free(keep);
        free(own_sizes);
        return frames_len;
    }
    if (n_frames_out) *n_frames_out = nf;

    /* STREAMINFO min/max frame size  :2413-2436 */
    uint32_t minf = 0, maxf = 0;
    uint64_t off = 0;
    for (uint64_t f = 0; f < n_frames; f++) {
        uint32_t s = sizes[f];
        pts[f].sample_offset = f * bs; /* Encoder::encode seekpoints  :1999-2003 */
        pts[f].byte_offset = off;
        uint64_t rem = n_pcm_frames - f * bs;
        pts[f].frame_samples = (uint16_t)(rem < bs ? rem : bs);
        pts[f].defined = 1;
        off += s;
        if (s < 0xFFFFFF && s != 0) { /* Streaminfo::MAX_FRAME_SIZE = 2^24 - 1 */
            minf = minf == 0 ? s : (s < minf ? s : minf);
            maxf = maxf == 0 ? s : (s > maxf ? s : maxf);
        }
    }
    /* MD5 over little-endian interleaved samples  :1292-1318 */
    uint8_t md5[16];
    {
        uint32_t bytes_per_sample = (bps + 7) / 8;
        md5_ctx c;
        md5_init(&c);
        uint8_t buf[4096 * 4];
        size_t total = (size_t)n_pcm_frames * channels, done = 0;
        while (done < total) {
            size_t m = total - done < 4096 ? total - done : 4096;
            fo_samples_to_bytes(interleaved + done, m, bytes_per_sample, 0, buf);
            md5_update(&c, buf, m * bytes_per_sample);
            done += m;
        }
        md5_final(&c, md5);
    }
    /* finalize_inner  :2024-2110 */
    size_t n_final_points = 0;
    size_t* final_keep = (size_t*)malloc(sizeof(size_t) * (n_frames + 1));
    int insert_seektable_after_padding = 0;
    if (opt->seektable_kind) {
        n_final_points = seek_filter(opt, sample_rate, pts, n_frames, final_keep);
        if (!have_seektable && opt->padding >= 0) { /* (None, Some(Padding))  :2053 */
            if (n_final_points > MAX_POINTS) n_final_points = MAX_POINTS;
            size_t seektable_size = 4 + 18 * n_final_points; /* total_size(): header + body */
            if ((size_t)opt->padding >= seektable_size) {
                insert_seektable_after_padding = 1;
                pad_bytes = 4 + ((size_t)opt->padding - seektable_size);
            }
        }
    }
    uint8_t* p = out;
    memcpy(p, "fLaC", 4);
    p += 4;
    int si_last = !(have_seektable || opt->padding >= 0);
    p[0] = (uint8_t)((si_last ? 0x80 : 0) | 0);
    put_be(p + 1, 34, 3);
    write_streaminfo_body(p + 4, bs, minf, maxf, sample_rate, channels, bps, n_pcm_frames, md5);
    p += 38;
    if (have_seektable) {
        int last = !(opt->padding >= 0);
        p[0] = (uint8_t)((last ? 0x80 : 0) | 3);
        put_be(p + 1, 18 * n_seek_slots, 3);
        p += 4;
        for (size_t i = 0; i < n_seek_slots; i++) { /* :2041-2051: defined points then placeholders */
            if (i < n_final_points) {
                const seekpoint_t* s = &pts[final_keep[i]];
                put_be(p, s->sample_offset, 8);
                put_be(p + 8, s->byte_offset, 8);
                put_be(p + 16, s->frame_samples, 2);
            } else {
                put_be(p, UINT64_MAX, 8);
                put_be(p + 8, 0, 8);
                put_be(p + 16, 0, 2);
            }
            p += 18;
        }
    }
    if (opt->padding >= 0) {
        int last = !insert_seektable_after_padding;
        size_t body = pad_bytes - 4;
        p[0] = (uint8_t)((last ? 0x80 : 0) | 1);
        put_be(p + 1, body, 3);
        memset(p + 4, 0, body);
        p += pad_bytes;
        if (insert_seektable_after_padding) { /* blocks.insert pushes after PADDING  :2071 */
            p[0] = (uint8_t)(0x80 | 3);
            put_be(p + 1, 18 * n_final_points, 3);
            p += 4;
            for (size_t i = 0; i < n_final_points; i++) {
                const seekpoint_t* s = &pts[final_keep[i]];
                put_be(p, s->sample_offset, 8);
                put_be(p + 8, s->byte_offset, 8);
                put_be(p + 16, s->frame_samples, 2);
                p += 18;
            }
        }
    }
    free(pts);
    free(keep);
    free(final_keep);
    free(own_sizes);
    return (int64_t)header_len + frames_len;
}

/* ==========================================================================================
 * DECODE
 * ======================================================================================== */
typedef struct {
    const uint8_t* data;
    size_t len;  /* bytes */
    uint64_t pos; /* bits */
    int err;
} bitreader;

static inline uint32_t br_read(bitreader* r, uint32_t n)
{
    if (n == 0) return 0;
    if (r->pos + n > (uint64_t)r->len * 8) {
        r->err = FO_ERR_IO; /* UnexpectedEof */
        r->pos = (uint64_t)r->len * 8;
        return 0;
    }
    uint64_t v = 0;
    uint64_t pos = r->pos;
    uint32_t left = n;
    while (left) {
        uint32_t bit_in_byte = (uint32_t)(pos & 7);
        uint32_t room = 8 - bit_in_byte;
        uint32_t take = left < room ? left : room;
        uint32_t chunk = (r->data[pos >> 3] >> (room - take)) & ((1u << take) - 1u);
        v = (v << take) | chunk;
        pos += take;
        left -= take;
    }
    r->pos = pos;
    return (uint32_t)v;
}

static inline int64_t br_read_signed(bitreader* r, uint32_t n) /* n in 1..=33 */
{
    uint64_t v;
    if (n > 32) {
        uint64_t hi = br_read(r, n - 32);
        v = (hi << 32) | br_read(r, 32);
    } else {
        v = br_read(r, n);
    }
    uint32_t sh = 64 - n;
    return (int64_t)(v << sh) >> sh;
}

/* read_unary::<1>: count zero bits up to the next one bit */
static inline uint32_t br_read_unary1(bitreader* r)
{
    uint32_t q = 0;
    uint64_t end = (uint64_t)r->len * 8;
    while (r->pos < end) {
        uint32_t bit_in_byte = (uint32_t)(r->pos & 7);
        uint8_t byte = (uint8_t)(r->data[r->pos >> 3] << bit_in_byte);
        if (byte == 0) {
            q += 8 - bit_in_byte;
            r->pos += 8 - bit_in_byte;
            continue;
        }
        uint32_t lz = (uint32_t)__builtin_clz((uint32_t)byte) - 24;
        q += lz;
        r->pos += lz + 1;
        return q;
    }
    r->err = FO_ERR_IO;
    return q;
}

/* predict  src/decode.rs:1738-1752 (i32 lane wraps like a release build; i64 lane for 33-bit side) */
void fo_predict(const int64_t* coefficients, uint32_t order, uint32_t shift, int32_t* ch, uint32_t n)
{
    for (uint32_t i = order; i < n; i++) {
        int64_t sum = 0;
        for (uint32_t j = 0; j < order; j++) sum += (int64_t)ch[i - 1 - j] * coefficients[j];
        ch[i] = (int32_t)((uint32_t)ch[i] + (uint32_t)(uint64_t)(sum >> shift));
    }
}

static void predict64(const int64_t* coefficients, uint32_t order, uint32_t shift, int64_t* ch, uint32_t n, int wide)
{
    for (uint32_t i = order; i < n; i++) {
        int64_t sum = 0;
        for (uint32_t j = 0; j < order; j++) sum += ch[i - 1 - j] * coefficients[j];
        if (wide) ch[i] = ch[i] + (sum >> shift);
        else ch[i] = (int64_t)(int32_t)((uint32_t)ch[i] + (uint32_t)(uint64_t)(sum >> shift));
    }
}

/* read_residuals  src/decode.rs:1800-1856 */
static int read_residuals(bitreader* r, uint32_t predictor_order, int64_t* res, uint32_t n_res)
{
    uint32_t method = br_read(r, 2);
    if (r->err) return r->err;
    if (method > 1) return FO_ERR_INVALID_CODING_METHOD;
    uint32_t rice_max = method ? 31 : 15;
    uint32_t hdr_bits = method ? 5 : 4;
    uint32_t block_size = predictor_order + n_res;
    uint32_t partition_order = br_read(r, 4);
    if (r->err) return r->err;
    uint32_t partition_count = 1u << partition_order;
    uint32_t chunk = block_size / partition_count;
    if (chunk == 0) return FO_ERR_INVALID_PARTITION_ORDER; /* rchunks_mut(0) panics in the reference */
    uint32_t count = (n_res + chunk - 1) / chunk;
    if (count != partition_count) return FO_ERR_INVALID_PARTITION_ORDER; /* :1818 */
    uint32_t pos = 0;
    uint32_t first_len = n_res - (count - 1) * chunk;
    for (uint32_t j = 0; j < count; j++) {
        uint32_t len = j == 0 ? first_len : chunk;
        uint32_t rice = br_read(r, hdr_bits);
        if (r->err) return r->err;
        if (rice == rice_max) { /* src/stream.rs:1590-1596 */
            uint32_t esc = br_read(r, 5);
            if (r->err) return r->err;
            if (esc) {
                for (uint32_t i = 0; i < len; i++) res[pos + i] = br_read_signed(r, esc); /* :1836 */
            } else {
                for (uint32_t i = 0; i < len; i++) res[pos + i] = 0; /* :1842 */
            }
        } else {
            for (uint32_t i = 0; i < len; i++) { /* :1823-1834 */
                uint32_t msb = br_read_unary1(r);
                uint32_t lsb = br_read(r, rice);
                uint32_t u = (msb << rice) | lsb;
                res[pos + i] = (u & 1) ? -(int64_t)(u >> 1) - 1 : (int64_t)(u >> 1);
            }
        }
        if (r->err) return r->err;
        pos += len;
    }
    return 0;
}

static const int64_t FIXED_COEFFS[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}}; /* src/stream.rs:1534 */

/* read_subframe  src/decode.rs:1635-1676; bps up to 33 (wide != 0 keeps i64 arithmetic) */
static int read_subframe(bitreader* r, uint32_t bps, int64_t* ch, uint32_t n, int wide)
{
    if (br_read(r, 1) != 0) return r->err ? r->err : FO_ERR_INVALID_SUBFRAME_HEADER; /* src/stream.rs:1385 */
    uint32_t type = br_read(r, 6);
    uint32_t wasted = 0;
    if (br_read(r, 1)) wasted = br_read_unary1(r) + 1;
    if (r->err) return r->err;
    int kind, order = 0;
    if (type == 0) kind = 0;
    else if (type == 1) kind = 1;
    else if (type >= 8 && type <= 12) {
        kind = 2;
        order = (int)type - 8;
    } else if (type >= 32) {
        kind = 3;
        order = (int)type - 31;
    } else
        return FO_ERR_INVALID_SUBFRAME_HEADER_TYPE; /* src/stream.rs:1550 */
    if (wasted > bps || bps - wasted == 0) return FO_ERR_EXCESSIVE_WASTED_BITS; /* checked_sub on SignedBitCount  :1644 */
    uint32_t ebps = bps - wasted;
    if (kind == 0) {
        int64_t v = br_read_signed(r, ebps);
        for (uint32_t i = 0; i < n; i++) ch[i] = v;
    } else if (kind == 1) {
        for (uint32_t i = 0; i < n; i++) ch[i] = br_read_signed(r, ebps);
    } else if (kind == 2) {
        if ((uint32_t)order > n) return FO_ERR_INVALID_FIXED_ORDER; /* :1684 */
        for (int i = 0; i < order; i++) ch[i] = br_read_signed(r, ebps);
        if (r->err) return r->err;
        int rc = read_residuals(r, (uint32_t)order, ch + order, n - (uint32_t)order);
        if (rc) return rc;
        predict64(FIXED_COEFFS[order], (uint32_t)order, 0, ch, n, wide);
    } else {
        if ((uint32_t)order > n) return FO_ERR_INVALID_LPC_ORDER; /* :1706 */
        for (int i = 0; i < order; i++) ch[i] = br_read_signed(r, ebps);
        uint32_t prec = br_read(r, 4) + 1;
        if (r->err) return r->err;
        if (prec > 15) return FO_ERR_INVALID_QLP_PRECISION; /* :1715-1719 */
        int64_t shift = br_read_signed(r, 5);
        if (r->err) return r->err;
        if (shift < 0) return FO_ERR_NEGATIVE_LPC_SHIFT; /* :1721-1724 */
        int64_t coefs[32];
        for (int i = 0; i < order; i++) coefs[i] = br_read_signed(r, prec);
        if (r->err) return r->err;
        int rc = read_residuals(r, (uint32_t)order, ch + order, n - (uint32_t)order);
        if (rc) return rc;
        predict64(coefs, (uint32_t)order, (uint32_t)shift, ch, n, wide);
    }
    if (r->err) return r->err;
    if (wasted) { /* :1671 */
        if (wide) for (uint32_t i = 0; i < n; i++) ch[i] = (int64_t)((uint64_t)ch[i] << wasted);
        else for (uint32_t i = 0; i < n; i++) ch[i] = (int64_t)(int32_t)((uint32_t)ch[i] << wasted);
    }
    return 0;
}

/* FrameHeader::parse + STREAMINFO cross checks  src/stream.rs:214-240, :279-313 */
static int64_t parse_frame_header(const uint8_t* d, size_t len, const fo_streaminfo* si, fo_frame_header* h)
{
    if (len < 4) return -FO_ERR_IO;
    if (d[0] != 0xFF || (d[1] & 0xFE) != 0xF8) return -FO_ERR_INVALID_SYNC_CODE;
    h->blocking_strategy = d[1] & 1;
    uint32_t bsc = d[2] >> 4, src = d[2] & 15, ca = d[3] >> 4, bpc = (d[3] >> 1) & 7;
    if (bsc == 0) return -FO_ERR_INVALID_BLOCK_SIZE;
    uint32_t rate = 0;
    int rate_kind = 0;
    static const uint32_t rates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    if (src == 0) {
        if (!si) return -FO_ERR_NON_SUBSET_SAMPLE_RATE;
        rate = si->sample_rate;
    } else if (src <= 11) rate = rates[src];
    else if (src == 12) rate_kind = 1;
    else if (src == 13) rate_kind = 2;
    else if (src == 14) rate_kind = 3;
    else return -FO_ERR_INVALID_SAMPLE_RATE;
    if (ca > 10) return -FO_ERR_INVALID_CHANNELS;
    uint32_t bps;
    switch (bpc) {
    case 0:
        if (!si) return -FO_ERR_NON_SUBSET_BPS;
        bps = si->bps;
        break;
    case 1: bps = 8; break;
    case 2: bps = 12; break;
    case 3: return -FO_ERR_INVALID_BPS;
    case 4: bps = 16; break;
    case 5: bps = 20; break;
    case 6: bps = 24; break;
    default: bps = 32; break;
    }
    /* r.skip(1): the reserved bit is not checked by the reference  :226 */
    size_t n = 4;
    uint64_t fnum;
    int fl = fo_read_frame_number(d + n, len - n, &fnum);
    if (fl < 0) return fl;
    n += (size_t)fl;
    uint32_t block_size;
    static const uint32_t bsizes[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    if (bsc == 6) {
        if (n + 1 > len) return -FO_ERR_IO;
        block_size = (uint32_t)d[n] + 1;
        n += 1;
    } else if (bsc == 7) {
        if (n + 2 > len) return -FO_ERR_IO;
        uint32_t v = ((uint32_t)d[n] << 8) | d[n + 1];
        if (v == 0xFFFF) return -FO_ERR_INVALID_BLOCK_SIZE; /* checked_add(1)  src/stream.rs:535 */
        block_size = v + 1;
        n += 2;
    } else
        block_size = bsizes[bsc];
    if (rate_kind == 1) {
        if (n + 1 > len) return -FO_ERR_IO;
        rate = (uint32_t)d[n] * 1000;
        n += 1;
    } else if (rate_kind) {
        if (n + 2 > len) return -FO_ERR_IO;
        rate = ((uint32_t)d[n] << 8) | d[n + 1];
        if (rate_kind == 3) rate *= 10;
        n += 2;
    }
    if (n + 1 > len) return -FO_ERR_IO;
    n += 1; /* CRC-8 byte */
    h->block_size = block_size;
    h->sample_rate = rate;
    h->bps = bps;
    h->channel_assignment = ca;
    h->channels = ca <= 7 ? ca + 1 : 2;
    h->frame_number = fnum;
    h->header_bytes = (uint32_t)n;
    if (si) { /* src/stream.rs:291-312, in this order */
        if (block_size > si->max_block_size) return -FO_ERR_BLOCK_SIZE_MISMATCH;
        if (rate != si->sample_rate) return -FO_ERR_SAMPLE_RATE_MISMATCH;
        if (h->channels != si->channels) return -FO_ERR_CHANNELS_MISMATCH;
        if (bps != si->bps) return -FO_ERR_BPS_MISMATCH;
    }
    if (fo_crc8(d, n) != 0) return -FO_ERR_CRC8_MISMATCH; /* src/stream.rs:158-163 */
    return (int64_t)n;
}

int64_t fo_decode_frame(const uint8_t* data, size_t len, const fo_streaminfo* si, uint64_t remaining, int32_t* out,
                        size_t out_cap, fo_frame_header* hdr_out)
{
    init_tables();
    fo_frame_header h;
    memset(&h, 0, sizeof(h));
    int64_t hl = parse_frame_header(data, len, si, &h);
    if (hl < 0) return hl;
    if (hdr_out) *hdr_out = h;
    uint32_t n = h.block_size;
    if (remaining && !((uint64_t)n == remaining || n > 14)) return -FO_ERR_SHORT_BLOCK; /* src/decode.rs:1405-1410 */
    if ((size_t)n * h.channels > out_cap) return -FO_ERR_IO;
    bitreader r = {data, len, (uint64_t)hl * 8, 0};
    int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * (size_t)n * 2);
    int64_t *a = tmp, *b = tmp + n;
    int rc = 0;
    uint32_t bps = h.bps;
    if (h.channel_assignment <= 7) { /* src/decode.rs:1502-1511 */
        for (uint32_t c = 0; c < h.channels && !rc; c++) {
            rc = read_subframe(&r, bps, a, n, 0);
            if (!rc)
                for (uint32_t i = 0; i < n; i++) out[(size_t)c * n + i] = (int32_t)a[i];
        }
    } else {
        int wide = bps == 32; /* 33-bit side  :1528 */
        int32_t *o0 = out, *o1 = out + n;
        if (h.channel_assignment == 8) { /* LeftSide  :1512 */
            rc = read_subframe(&r, bps, a, n, 0);
            if (!rc) rc = read_subframe(&r, bps + 1, b, n, wide);
            if (!rc)
                for (uint32_t i = 0; i < n; i++) {
                    o0[i] = (int32_t)a[i];
                    o1[i] = wide ? (int32_t)(a[i] - b[i]) : (int32_t)((uint32_t)(int32_t)a[i] - (uint32_t)(int32_t)b[i]);
                }
        } else if (h.channel_assignment == 9) { /* SideRight  :1549 */
            rc = read_subframe(&r, bps + 1, a, n, wide);
            if (!rc) rc = read_subframe(&r, bps, b, n, 0);
            if (!rc)
                for (uint32_t i = 0; i < n; i++) {
                    o0[i] = wide ? (int32_t)(a[i] + b[i]) : (int32_t)((uint32_t)(int32_t)a[i] + (uint32_t)(int32_t)b[i]);
                    o1[i] = (int32_t)b[i];
                }
        } else { /* MidSide  :1586 */
            rc = read_subframe(&r, bps, a, n, 0);
            if (!rc) rc = read_subframe(&r, bps + 1, b, n, wide);
            if (!rc)
                for (uint32_t i = 0; i < n; i++) {
                    if (wide) {
                        int64_t side = b[i];
                        int64_t sum = a[i] * 2 + (side < 0 ? (-side) % 2 : side % 2); /* :1619 */
                        o0[i] = (int32_t)((sum + side) >> 1);
                        o1[i] = (int32_t)((sum - side) >> 1);
                    } else {
                        int32_t mid = (int32_t)a[i], side = (int32_t)b[i];
                        /* i32 arithmetic, wrapping like a release build  :1599-1601 */
                        int32_t sum = (int32_t)((uint32_t)mid * 2u + (uint32_t)(side & 1));
                        o0[i] = (int32_t)((uint32_t)sum + (uint32_t)side) >> 1;
                        o1[i] = (int32_t)((uint32_t)sum - (uint32_t)side) >> 1;
                    }
                }
        }
    }
    free(tmp);
    if (rc) return -rc;
    /* byte_align + CRC-16  :1629-1630, :1429 */
    uint64_t end = (r.pos + 7) / 8;
    if (end + 2 > len) return -FO_ERR_IO;
    if (fo_crc16(data, (size_t)end + 2) != 0) return -FO_ERR_CRC16_MISMATCH;
    return (int64_t)end + 2;
}

/* BlockIterator / BlockList::read reduced to what the frame path needs  src/metadata/mod.rs:482-646 */
int fo_read_streaminfo(const uint8_t* f, size_t len, fo_streaminfo* si)
{
    if (len < 4 || memcmp(f, "fLaC", 4) != 0) return FO_ERR_MISSING_FLAC_TAG;
    size_t p = 4;
    int first = 1;
    for (;;) {
        if (p + 4 > len) return FO_ERR_IO;
        int last = f[p] >> 7;
        uint32_t type = f[p] & 0x7F;
        uint32_t blen = ((uint32_t)f[p + 1] << 16) | ((uint32_t)f[p + 2] << 8) | f[p + 3];
        p += 4;
        if (p + blen > len) return FO_ERR_IO;
        if (first) {
            if (type != 0 || blen != 34) return FO_ERR_MISSING_STREAMINFO;
            const uint8_t* b = f + p;
            si->min_block_size = (uint16_t)((b[0] << 8) | b[1]);
            si->max_block_size = (uint16_t)((b[2] << 8) | b[3]);
            si->min_frame_size = ((uint32_t)b[4] << 16) | ((uint32_t)b[5] << 8) | b[6];
            si->max_frame_size = ((uint32_t)b[7] << 16) | ((uint32_t)b[8] << 8) | b[9];
            uint64_t v = 0;
            for (int i = 0; i < 8; i++) v = (v << 8) | b[10 + i];
            si->sample_rate = (uint32_t)(v >> 44);
            si->channels = (uint8_t)(((v >> 41) & 7) + 1);
            si->bps = (uint8_t)(((v >> 36) & 31) + 1);
            si->total_samples = v & 0xFFFFFFFFFull;
            memcpy(si->md5, b + 18, 16);
            first = 0;
        } else if (type == 127) {
            return FO_ERR_INVALID_METADATA_BLOCK;
        }
        p += blen;
        if (last) break;
    }
    si->frames_start = p;
    return 0;
}

int64_t fo_decode_stream(const uint8_t* flac, size_t len, int32_t* out, size_t out_cap, fo_streaminfo* si_out, uint8_t md5_out[16])
{
    return fo_decode_stream_ex(flac, len, out, out_cap, si_out, md5_out, NULL, NULL);
}

/* as fo_decode_stream; also reports how far the serial reader got: frames delivered before the error (= index of the failing
 * frame) and interleaved samples written -- Decoder::read_frame hands every good frame out before it fails (src/decode.rs:1388) */
int64_t fo_decode_stream_ex(const uint8_t* flac, size_t len, int32_t* out, size_t out_cap, fo_streaminfo* si_out, uint8_t md5_out[16],
                            uint64_t* frames_done, uint64_t* samples_done)
{
    uint64_t nframes = 0;
    if (frames_done) *frames_done = 0;
    if (samples_done) *samples_done = 0;
    fo_streaminfo si;
    memset(&si, 0, sizeof(si));
    int rc = fo_read_streaminfo(flac, len, &si);
    if (rc) return -rc;
    if (si_out) *si_out = si;
    size_t p = (size_t)si.frames_start;
    uint64_t current = 0;
    size_t written = 0;
    int32_t* planar = (int32_t*)malloc(sizeof(int32_t) * 65536 * 8);
    int64_t result = 0;
    for (;;) {
        uint64_t remaining = 0;
        if (si.total_samples) {
            remaining = si.total_samples - current;
            if (remaining == 0) break; /* src/decode.rs:1402 */
        } else if (p >= len) {
            break; /* EOF at a frame boundary ends the stream  :1416 */
        }
        fo_frame_header h;
        int64_t used = fo_decode_frame(flac + p, len - p, &si, remaining, planar, 65536 * 8, &h);
        if (used == -FO_ERR_IO && !si.total_samples && len - p < 16) break; /* EOF inside a header ends an unsized stream  :1416 */
        if (used < 0) {
            result = used;
            break;
        }
        p += (size_t)used;
        size_t cnt = (size_t)h.block_size * h.channels;
        if (written + cnt > out_cap) {
            result = -FO_ERR_IO;
            break;
        }
        for (uint32_t i = 0; i < h.block_size; i++) /* Frame::iter interleave  src/audio.rs:94 */
            for (uint32_t c = 0; c < h.channels; c++) out[written + (size_t)i * h.channels + c] = planar[(size_t)c * h.block_size + i];
        written += cnt;
        current += h.block_size;
        nframes++;
    }
    free(planar);
    if (frames_done) *frames_done = nframes;
    if (samples_done) *samples_done = written;
    if (result < 0) return result;
    if (md5_out) {
        uint32_t bytes_per_sample = ((uint32_t)si.bps + 7) / 8;
        md5_ctx c;
        md5_init(&c);
        uint8_t buf[4096 * 4];
        size_t done = 0;
        while (done < written) {
            size_t m = written - done < 4096 ? written - done : 4096;
            fo_samples_to_bytes(out + done, m, bytes_per_sample, 0, buf);
            md5_update(&c, buf, m * bytes_per_sample);
            done += m;
        }
        md5_final(&c, md5_out);
    }
    return (int64_t)written;
}

int64_t fo_decode_frames_mt(const uint8_t* frames, const uint64_t* offsets, uint64_t n_frames, const fo_streaminfo* si,
                            int nthreads, int32_t* out, size_t out_cap)
{
    /* offsets has n_frames + 1 entries; every frame but the last has max_block_size samples */
    int64_t bad = 0;
    uint32_t bs = si->max_block_size, ch = si->channels;
    if (nthreads < 1) nthreads = 1;
    uint64_t total = 0;
#pragma omp parallel num_threads(nthreads)
    {
        int32_t* planar = (int32_t*)malloc(sizeof(int32_t) * 65536 * 8);
#pragma omp for schedule(dynamic, 4) reduction(+ : total)
        for (int64_t f = 0; f < (int64_t)n_frames; f++) {
            fo_frame_header h;
            int64_t used = fo_decode_frame(frames + offsets[f], (size_t)(offsets[f + 1] - offsets[f]), si, 0, planar, 65536 * 8, &h);
            if (used < 0) {
#pragma omp critical
                bad = used;
                continue;
            }
            size_t base = (size_t)f * bs * ch;
            if (base + (size_t)h.block_size * ch > out_cap) {
#pragma omp critical
                bad = -FO_ERR_IO;
                continue;
            }
            for (uint32_t i = 0; i < h.block_size; i++)
                for (uint32_t c = 0; c < ch; c++) out[base + (size_t)i * ch + c] = planar[(size_t)c * h.block_size + i];
            total += (uint64_t)h.block_size * ch;
        }
        free(planar);
    }
    if (bad < 0) return bad;
    return (int64_t)total;
}
