// glibc_log.cuh -- glibc 2.39's log() and log2() for doubles, restated operation by operation.
//
// The reference evaluates f64::ln in the LPC order estimate (src/encode.rs:3674-3675) and f64::log2 in the coefficient
// shift (:3360); Rust's std forwards both to the platform libm, i.e. glibc's `log` / `log2` on Linux.  Those are the
// ARM "optimized routines" implementations (sysdeps/ieee754/dbl-64/e_log.c, e_log2.c): table + polynomial, worst-case
// error 0.52 / 0.55 ulp -- not correctly rounded, so any other libm (CUDA's libdevice is specified to 1 ulp) can return
// the neighbouring double for some inputs, and a neighbouring double can flip a first-minimum comparison between two LPC
// orders or floor(log2(l)) next to a power of two.  To make the encoder's choices identical BY CONSTRUCTION this file
// follows the FMA builds of both functions (__log_fma, __log2_fma -- what glibc's ifunc resolver picks on every x86-64
// CPU with FMA) instruction for instruction: which products are fused and which are rounded separately was read from the
// disassembly of this image's libm.so.6 (the compiler chose the contractions, the C source does not fix them).  Tables:
// glibc_log_tables.inc (tools/gen_glibc_log_tables.py).
//
// Checked bit-for-bit against the C library: tests/test_glibc_log_port.py (host build of this header, CPU) and
// tests/test_gpu_libm.py (device build, >= 1e8 inputs over the encoder's domain).
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define GL_MUL(a, b) __dmul_rn((a), (b))
#define GL_ADD(a, b) __dadd_rn((a), (b))
#define GL_SUB(a, b) __dsub_rn((a), (b))
#define GL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GL_BITS(x) ((uint64_t)__double_as_longlong(x))
#define GL_DBL(u) __longlong_as_double((long long)(u))
#else
// host build (tests only; compiled with -ffp-contract=off): plain IEEE operations and the C library's exact fma
#include <cmath>
static inline uint64_t gl_bits_host(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double gl_dbl_host(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define GL_MUL(a, b) ((a) * (b))
#define GL_ADD(a, b) ((a) + (b))
#define GL_SUB(a, b) ((a) - (b))
#define GL_FMA(a, b, c) fma((a), (b), (c))
#define GL_BITS(x) gl_bits_host(x)
#define GL_DBL(u) gl_dbl_host(u)
#endif

#ifdef __CUDACC__
#define GL_FN __device__ inline
#else
#define __device__
#define GL_FN static inline
#endif

namespace flacb200 {

#include "glibc_log_tables.inc"

// log(x) for finite x > 0 (the encoder never passes anything else: take_while(err > 0) precedes the call; zero,
// negative, infinite and NaN arguments return what IEEE arithmetic makes of them below, not glibc's errno paths)
GL_FN double glibc_log(double x)
{
    uint64_t ix = GL_BITS(x);
    const uint64_t LO = 0x3fee000000000000ull;              // asuint64(1.0 - 0x1p-4)
    if (ix - LO <= 0x308ffffffffffull) {                    // < asuint64(1.0 + 0x1.09p-4) - LO: close to 1.0
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double* B = glog_B;
        const double r = GL_SUB(x, 1.0);
        const double r2 = GL_MUL(r, r);
        const double r3 = GL_MUL(r, r2);
        const double pa = GL_FMA(r2, B[3], GL_FMA(r, B[2], B[1]));
        const double pb = GL_FMA(r2, B[6], GL_FMA(r, B[5], B[4]));
        double pc = GL_FMA(r2, B[9], GL_FMA(r, B[8], B[7]));
        pc = GL_FMA(r3, B[10], pc);
        double p = GL_FMA(pc, r3, pb);
        p = GL_FMA(p, r3, pa);                              // y = r3 * p is fused into the final sum below
        const double rhi = GL_FMA(-0x1p27, r, GL_FMA(r, 0x1p27, r));   // r + w - w, w = r * 2^27
        const double rlo = GL_SUB(r, rhi);
        const double rhi2 = GL_MUL(rhi, rhi);
        const double hi = GL_FMA(rhi2, B[0], r);            // r + rhi * rhi * B[0]
        double lo = GL_FMA(rhi2, B[0], GL_SUB(r, hi));      // r - hi + w
        lo = GL_FMA(GL_MUL(B[0], rlo), GL_ADD(r, rhi), lo);
        return GL_ADD(hi, GL_FMA(p, r3, lo));
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if (ix * 2 == 0) return GL_DBL(0xfff0000000000000ull);   // -inf
        if (ix == 0x7ff0000000000000ull) return x;
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return GL_DBL(0x7ff8000000000000ull);   // NaN (glibc: __math_invalid)
        ix = GL_BITS(GL_MUL(x, 0x1p52));                    // subnormal: normalise
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
    const double kd = (double)(int32_t)((int64_t)tmp >> 52);
    const double z = GL_DBL(ix - (tmp & (0xfffull << 52)));
    const double invc = glog_T[2 * i], logc = glog_T[2 * i + 1];
    const double* A = glog_A;
    const double w = GL_FMA(kd, GLOG_LN2HI, logc);
    const double r = GL_FMA(z, invc, -1.0);
    const double p12 = GL_FMA(r, A[2], A[1]);
    const double hi = GL_ADD(r, w);
    const double r2 = GL_MUL(r, r);
    double lo = GL_ADD(GL_SUB(w, hi), r);
    lo = GL_FMA(kd, GLOG_LN2LO, lo);
    const double r3 = GL_MUL(r, r2);
    const double p34 = GL_FMA(r, A[4], A[3]);
    lo = GL_FMA(r2, A[0], lo);
    const double p = GL_FMA(p34, r2, p12);
    return GL_ADD(GL_FMA(r3, p, lo), hi);
}

// log2(x), same contract
GL_FN double glibc_log2(double x)
{
    uint64_t ix = GL_BITS(x);
    const uint64_t LO = 0x3feea4af00000000ull;              // asuint64(1.0 - 0x1.5b51p-5)
    if (ix - LO <= 0x210a9ffffffffull) {                    // < asuint64(1.0 + 0x1.6ab2p-5) - LO
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double* B = glog2_B;
        const double r = GL_SUB(x, 1.0);
        const double hi = GL_MUL(GLOG2_INVLN2HI, r);
        const double r2 = GL_MUL(r, r);
        const double e = GL_FMA(GLOG2_INVLN2HI, r, -hi);
        const double r4 = GL_MUL(r2, r2);
        const double b01 = GL_FMA(r, B[1], B[0]);
        double lo = GL_FMA(r, GLOG2_INVLN2LO, e);
        const double y = GL_FMA(b01, r2, hi);               // hi + p, p = r2 * (B[0] + r * B[1])
        const double t = GL_FMA(b01, r2, GL_SUB(hi, y));    // hi - y + p
        lo = GL_ADD(t, lo);
        const double b23 = GL_FMA(r, B[3], B[2]);
        const double b45 = GL_FMA(r, B[5], B[4]);
        const double q0 = GL_FMA(b45, r2, b23);
        const double b67 = GL_FMA(r, B[7], B[6]);
        const double b89 = GL_FMA(r, B[9], B[8]);
        double q = GL_FMA(b89, r2, b67);
        q = GL_FMA(q, r4, q0);
        q = GL_FMA(q, r4, lo);
        return GL_ADD(y, q);
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if (ix * 2 == 0) return GL_DBL(0xfff0000000000000ull);   // -inf
        if (ix == 0x7ff0000000000000ull) return x;
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return GL_DBL(0x7ff8000000000000ull);
        ix = GL_BITS(GL_MUL(x, 0x1p52));
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const uint32_t i = (uint32_t)(tmp >> 46) & 63u;
    const double kd = (double)(int32_t)((int64_t)tmp >> 52);
    const double invc = glog2_T[2 * i], logc = glog2_T[2 * i + 1];
    const double* A = glog2_A;
    const double t3 = GL_ADD(kd, logc);
    const double z = GL_DBL(ix - (tmp & (0xfffull << 52)));
    const double r = GL_FMA(z, invc, -1.0);
    const double a01 = GL_FMA(r, A[1], A[0]);
    const double t1 = GL_MUL(GLOG2_INVLN2HI, r);
    const double e = GL_FMA(GLOG2_INVLN2HI, r, -t1);
    const double hi = GL_ADD(t1, t3);
    const double t2 = GL_FMA(r, GLOG2_INVLN2LO, e);
    const double r2 = GL_MUL(r, r);
    double lo = GL_ADD(GL_SUB(t3, hi), t1);
    lo = GL_ADD(lo, t2);
    const double a23 = GL_FMA(r, A[3], A[2]);
    const double r4 = GL_MUL(r2, r2);
    const double a45 = GL_FMA(r, A[5], A[4]);
    double p = GL_FMA(a23, r2, a01);
    p = GL_FMA(a45, r4, p);
    return GL_ADD(GL_FMA(r2, p, lo), hi);
}

}   // namespace flacb200
