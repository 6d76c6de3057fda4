// strict_math.h -- bit-exact twins of the libm calls on the reference's CPU sweep.
//
// The reference's CPU solver (libepic/src/harmonic/harmonic_cpu.cpp:65-70, :117-124) evaluates
//     u = maxVal + std::log(std::exp(a - maxVal) + ... ) - std::log(2.0 * n)
// with float arguments, i.e. it calls glibc's expf and logf.  Those are third-party code that is
// not under /root/reference: glibc 2.39 (Ubuntu 2.39-0ubuntu8.5), sysdeps/ieee754/flt-32/e_expf.c
// and e_logf.c, in the FMA ifunc variants every current x86-64 host selects.  To reproduce the
// reference's fields bit for bit, the CUDA sweep replays the published algorithm of those two
// functions operation by operation in IEEE double arithmetic (the same operation order and
// the same fused multiply-adds as the compiled libm, checked against its disassembly), with
// the same tables.  The argument ranges the sweep can produce are covered exactly:
//     strict_expf(x)  for every x <= 0 (and NaN)      -- the sweep only passes a - max(a, ...) <= 0
//     strict_logf(x)  for every normal x > 0          -- the sweep passes a sum in [1, 2n]
// tests/test_strict_math.py checks both against the host libm, exhaustively over those ranges on
// the GPU and on the CPU (this header also compiles as plain C++ for that purpose).
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define EPIC_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define EPIC_HD static inline
#endif

namespace epic_b200 {

// 2^(i/32) as double bits, minus (i << 47) -- the table `T` of glibc's __exp2f_data (N = 32).
#define EPIC_EXP2F_TABLE                                                                       \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull

// {invc, logc} pairs of glibc's __logf_data (LOGF_TABLE_BITS = 4).
#define EPIC_LOGF_TABLE                                         \
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2,                \
    0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,                \
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2,                \
    0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,                \
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3,                \
    0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,                \
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4,                \
    0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,                \
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5,                \
    0x1.0000000000000p+0, 0x0.0p+0,                             \
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,                 \
    0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,                 \
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,                 \
    0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,                 \
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,                 \
    0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2

constexpr double kExpShift = 0x1.8p52;                    // __exp2f_data.shift
constexpr double kExpInvLn2N = 0x1.71547652b82fep+5;      // N / ln 2, N = 32
constexpr double kExpC0 = 0x1.c6af84b912394p-20;
constexpr double kExpC1 = 0x1.ebfce50fac4f3p-13;
constexpr double kExpC2 = 0x1.62e42ff0c52d6p-6;
constexpr double kLogLn2 = 0x1.62e42fefa39efp-1;
constexpr double kLogA0 = -0x1.00ea348b88334p-2;
constexpr double kLogA1 = 0x1.5575b0be00b6ap-2;
constexpr double kLogA2 = -0x1.ffffef20a4123p-2;
constexpr double kLog4 = 0x1.62e42fefa39efp+0;            // glibc log(2.0 * 2)
constexpr double kLog6 = 0x1.cab0bfa2a2002p+0;            // glibc log(2.0 * 3)

EPIC_HD double strict_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
EPIC_HD double strict_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
EPIC_HD double strict_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
EPIC_HD uint64_t strict_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
EPIC_HD double strict_from_bits(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
EPIC_HD uint32_t strict_fbits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
    return b;
#endif
}
EPIC_HD float strict_from_fbits(uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// glibc expf for x <= 0.  `tab` = EPIC_EXP2F_TABLE (32 x uint64).
//
// e_expf.c returns +0 below log(2^-150) and the smallest denormal below log(2^-149) from its
// |x| >= 88 branch; everywhere else it evaluates, in double,
//     kd = round(x * N/ln2), r = x * N/ln2 - kd, s = 2^(kd/N) (table), y = (C0*r + C1)*r^2 + (C2*r + 1), y*s
// and narrows to float.  What must be reproduced is the FLOAT result, and the argument domain is small
// enough (2^31 floats) to check a cheaper evaluation exhaustively (tests/native/strict_math_check.cpp,
// run by tests/test_strict_math.py on the CPU and epic_b200_selftest_math on the GPU):
//   * clamping x at -104.5 replaces both underflow branches (the main path narrows to the same +0 /
//     2^-149 results), and
//   * the polynomial in Horner form with s folded in, s + (s*r) * ((C0*r + C1)*r + C2), needs 7 double
//     operations instead of 8
//   * the float -> double widening of x done on the bit pattern (exponent rebias + shift; exact for
//     every normal x, while +-0 and denormals turn into negative numbers below 2^-126 in magnitude, for
//     which the result is the same 1.0f) -- on the GPU this moves the conversion off the quarter-rate
//     conversion pipe (F2F.F64.F32), the busiest pipe of the strict sweep,
// give the same float as glibc for every x <= 0.  NaN is not reproduced bit for bit (payload).
EPIC_HD double strict_widen_nonpos(float xc)
{
    const uint32_t b = strict_fbits(xc);
    // x <= 0 has the sign bit set: (b >> 3) carries it into bit 28, and 0x10000000 + 0xA8000000 = 0xB8000000
    // = sign | (1023 - 127) << 20.  (x == +0 becomes -2^-383: same result, 1.0f.)
    const uint32_t hi = (b >> 3) + 0xA8000000u;
    const uint32_t lo = b << 29;
    return strict_from_bits(((uint64_t)hi << 32) | lo);
}

template <typename Table>
EPIC_HD float strict_expf_nonpos(float x, const Table &tab)
{
    const float xc = (x < -104.5f) ? -104.5f : x;
    const double xd = strict_widen_nonpos(xc);
    const double kdp = strict_fma(kExpInvLn2N, xd, kExpShift);   // z + SHIFT, contracted
    const uint64_t ki = strict_bits(kdp);
    const double kd = strict_add(kdp, -kExpShift);
    const double r = strict_fma(kExpInvLn2N, xd, -kd);           // z - kd, contracted
    const uint64_t t = tab[ki & 31] + (ki << 47);
    const double s = strict_from_bits(t);
    double y = strict_fma(kExpC0, r, kExpC1);
    y = strict_fma(y, r, kExpC2);
    const double sr = strict_mul(s, r);
    return (float)strict_fma(y, sr, s);
}

// glibc logf for normal x > 0.  `tab` = EPIC_LOGF_TABLE (16 x {invc, logc} doubles).
// e_logf.c evaluates y0 = k*Ln2 + logc, r = z*invc - 1, (A1*r + A2 + A0*r^2)*r^2 + (y0 + r) in double;
// the Horner form (((A0*r + A1)*r + A2)*r + 1)*r + y0 (one operation fewer) narrows to the same float
// for every x in [1/8, 16), checked exhaustively like expf above.  (x == 1 gives +0 through this path.)
template <typename Table>
EPIC_HD float strict_logf_normal(float x, const Table &tab)
{
    const uint32_t ix = strict_fbits(x);
    const uint32_t tmp = ix - 0x3f330000u;
    const uint32_t i = (tmp >> 19) & 15u;
    const int32_t k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = tab[2 * i], logc = tab[2 * i + 1];
    // z is a normal positive float: widening on the bit pattern is exact
    const double z = strict_from_bits(((uint64_t)((iz >> 3) + 0x38000000u) << 32) | (uint32_t)(iz << 29));
    const double r = strict_fma(z, invc, -1.0);
    const double y0 = strict_fma((double)k, kLogLn2, logc);
    double y = strict_fma(kLogA0, r, kLogA1);
    y = strict_fma(y, r, kLogA2);
    y = strict_fma(y, r, 1.0);
    return (float)strict_fma(y, r, y0);
}

// The sweep's logf argument is a sum of 2n <= 6 terms, each <= 1 and one of them == 1: x in [1, 6], where
// glibc's exponent k is 0..3.  For that range the two table operands can be indexed by (k, i) together:
//   y0[k*16 + i]   = fma(k, Ln2, logc[i])   (the identical double glibc computes first), and
//   invc[i] * 2^-k (exact), so that r = fma(x, invc*2^-k, -1) is glibc's fma(z, invc, -1) with z = x*2^-k
// without forming z.  This is the form the device code uses (math_policies.cuh, StrictMath::log_sum);
// checked exhaustively over [1, 8) like the functions above.
template <typename Table>
EPIC_HD float strict_logf_sum(float x, const Table &tab)
{
    const uint32_t ix = strict_fbits(x);
    const uint32_t tmp = ix - 0x3f330000u;
    const uint32_t ki = (tmp >> 19) & 63u;   // k * 16 + i
    const uint32_t i = ki & 15u, k = ki >> 4;
    const double invc_k = strict_from_bits(strict_bits(tab[2 * i]) - ((uint64_t)k << 52));
    const double y0 = strict_fma((double)k, kLogLn2, tab[2 * i + 1]);
    const double xd = strict_from_bits(((uint64_t)((ix >> 3) + 0x38000000u) << 32) | (uint32_t)(ix << 29));
    const double r = strict_fma(xd, invc_k, -1.0);
    double y = strict_fma(kLogA0, r, kLogA1);
    y = strict_fma(y, r, kLogA2);
    y = strict_fma(y, r, 1.0);
    return (float)strict_fma(y, r, y0);
}

// The 2-D update as the reference writes it (harmonic_cpu.cpp:58-70): max of the four neighbours, four
// expf, three float adds left to right, logf, the float add of the maximum and the subtraction of log(4.0)
// in double.  `e` / `l` are the expf / logf to use.
template <typename Exp, typename Log>
EPIC_HD float strict_update4_reference(float a, float b, float c, float d, Exp e, Log l)
{
    float mx = (a < b) ? b : a;
    mx = (mx < c) ? c : mx;
    mx = (mx < d) ? d : mx;
    float s = e(a - mx) + e(b - mx);
    s = s + e(c - mx);
    s = s + e(d - mx);
    const float t1 = mx + l(s);
    return (float)((double)t1 - kLog4);
}

// The same value with three expf: the form of StrictMath::update4 (math_policies.cuh).  The largest
// neighbour's term is expf(0) = 1; a min/max network finds the three losers; E(a) + E(b) is commutative, so
// only the (c, d) pair needs its order restored.
template <typename Exp, typename Log>
EPIC_HD float strict_update4_network(float a, float b, float c, float d, Exp e, Log l)
{
    const float hi1 = (a < b) ? b : a, lo1 = (a < b) ? a : b;
    const float hi2 = (c < d) ? d : c, lo2 = (c < d) ? c : d;
    const float mx = (hi1 < hi2) ? hi2 : hi1, mid = (hi1 < hi2) ? hi1 : hi2;
    const float e1 = e(lo1 - mx), e2 = e(lo2 - mx), em = e(mid - mx);
    const bool p = hi1 < hi2;
    const float eh1 = p ? em : 1.0f, eh2 = p ? 1.0f : em;
    const bool q = c < d;
    const float ec = q ? e2 : eh2, ed = q ? eh2 : e2;
    float s = eh1 + e1;
    s = s + ec;
    s = s + ed;
    const float t1 = mx + l(s);
    return (float)((double)t1 - kLog4);
}

}  // namespace epic_b200
