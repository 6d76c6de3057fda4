// math_policies.cuh -- the two arithmetic modes of the log-sum-exp update.
//
//   StrictMath  replays glibc's expf/logf in double arithmetic (strict_math.h) and performs the
//               reference CPU path's float adds, its double `- log(2n)` and its narrowing store
//               in the same order (reference libepic/src/harmonic/harmonic_cpu.cpp:58-70,
//               :107-124), so the field is bit-identical to harmonic_complete_cpu's.
//   FastMath    MUFU ex2/lg2 approximations, the arithmetic class of the reference's own GPU
//               kernel (harmonic_gpu.cu:52-61: __expf, __logf, constant 1.38629436f).  Stated
//               tolerance against the CPU path at matched epsilon: same iteration count and
//               |du| <= 1e-5*|u| + 4e-7*iterations (tests/test_parity_gpu.py explains the second term).
//
// This translation unit is compiled with -fmad=false: every multiply-add that may fuse is written
// as an explicit fma intrinsic.
#pragma once

#include "strict_math.h"

namespace epic_b200 {

// Shared-memory image of the logf tables, one 16-byte entry per (exponent, interval): a single LDS.128 per
// logarithm.  The 32-entry expf table does not live in shared memory at all: lane i of every warp keeps entry i in
// two registers and a lookup is a pair of SHFL.IDX (the index needs no masking and no address arithmetic: the
// shuffle takes the source lane modulo 32).  Both replace the ten LDS.32 + index arithmetic of the round-1 layout
// (32-bit words for bank freedom): the strict sweep is bound by instruction issue, not by shared-memory bandwidth.
struct alignas(16) MathTables {
    // {invc[i] * 2^-k (exact), fma(k, Ln2, logc[i])} at index k*16 + i, k = 0..3: the log argument is a sum of
    // 2n <= 6 terms in [1, 6], for which glibc's exponent k is 0..3, so its first FMA (and the int -> double
    // conversion of k) becomes a lookup of the identical double; see strict_logf_sum in strict_math.h.
    double2 logt[64];
    uint32_t exp_hi[32];
    uint32_t exp_lo[32];
};

static __device__ __constant__ uint64_t c_exp2f_table[32] = {EPIC_EXP2F_TABLE};
static __device__ __constant__ double c_logf_table[32] = {EPIC_LOGF_TABLE};

__device__ __forceinline__ void load_math_tables(MathTables *t, int tid, int nthreads)
{
    for (int i = tid; i < 32; i += nthreads) {
        t->exp_hi[i] = (uint32_t)(c_exp2f_table[i] >> 32);
        t->exp_lo[i] = (uint32_t)c_exp2f_table[i];
    }
    for (int i = tid; i < 64; i += nthreads) {
        const uint64_t a = (uint64_t)__double_as_longlong(c_logf_table[2 * (i & 15)]) - ((uint64_t)(i >> 4) << 52);
        double2 e;
        e.x = __longlong_as_double((long long)a);
        e.y = __fma_rn((double)(i >> 4), kLogLn2, c_logf_table[2 * (i & 15) + 1]);
        t->logt[i] = e;
    }
}

struct alignas(16) StrictMath {
    // Every double constant lives in the kernel-parameter (constant) bank, where DFMA/DADD read it as
    // an operand through uniform registers; as literals they would be re-materialised with two moves per
    // use.  The struct is 16-byte aligned with the doubles first so that pairs of them load with one
    // LDCU.128: when a longer Sweep2DParams once shifted them to 8 (mod 16), ptxas fell back to per-use
    // LDC.64 into vector registers and the strict sweep lost 6 %.
    double log2n;      // glibc log(2.0 * n)
    double inv_ln2n;   // kExpInvLn2N
    double shift;      // kExpShift
    double c0, c1, c2; // kExpC0..2
    double ln2, a0, a1, a2;
    const MathTables *t;
    uint32_t thi, tlo;  // this lane's entry of the expf table (lane i holds T[i])
    static constexpr bool kUsesTables = true;

    __host__ __device__ void init(double log_2n)
    {
        t = nullptr;
        log2n = log_2n;
        inv_ln2n = kExpInvLn2N;
        shift = kExpShift;
        c0 = kExpC0;
        c1 = kExpC1;
        c2 = kExpC2;
        ln2 = kLogLn2;
        a0 = kLogA0;
        a1 = kLogA1;
        a2 = kLogA2;
    }

    // After load_math_tables() and a barrier.
    __device__ __forceinline__ void bind(const MathTables *tables)
    {
        t = tables;
        thi = tables->exp_hi[threadIdx.x & 31];
        tlo = tables->exp_lo[threadIdx.x & 31];
    }

    // strict_expf_nonpos (strict_math.h).  WARP-SYNCHRONOUS: all 32 lanes must call it together (the table
    // lookup is a shuffle); the sweeps call it after a warp vote, with every lane computing.
    __device__ __forceinline__ float exp_nonpos(float x) const
    {
        // widening on the bit pattern (strict_widen_nonpos): two integer operations on lightly loaded
        // pipes instead of one F2F.F64.F32 on the quarter-rate conversion pipe
        const uint32_t xb = __float_as_uint(fmaxf(x, -104.5f));
        const double xd = __hiloint2double((int)((xb >> 3) + 0xA8000000u), (int)(xb << 29));
        const double kdp = __fma_rn(inv_ln2n, xd, shift);
        const uint32_t ki = (uint32_t)__double2loint(kdp);
        const double kd = __dadd_rn(kdp, -shift);
        const double r = __fma_rn(inv_ln2n, xd, -kd);
        const uint32_t shi = __shfl_sync(0xffffffffu, thi, (int)ki);   // source lane = ki mod 32
        const uint32_t slo = __shfl_sync(0xffffffffu, tlo, (int)ki);
        const double s = __hiloint2double((int)(shi + (ki << 15)), (int)slo);
        double y = __fma_rn(c0, r, c1);
        y = __fma_rn(y, r, c2);
        const double sr = __dmul_rn(s, r);
        return __double2float_rn(__fma_rn(y, sr, s));
    }

    // strict_logf_sum (strict_math.h) for x in [1, 2n]: both table operands indexed by (k, i).
    __device__ __forceinline__ float log_sum(float x) const
    {
        const uint32_t ix = __float_as_uint(x);
        // byte offset of entry k * 16 + i (k <= 3): ((ix - 0x3f330000) >> 19 & 63) * 16; the constant has no bits
        // below 2^15, so the subtraction commutes with the shift and the pair becomes one LEA.HI
        const uint32_t off = ((ix >> 15) - (0x3f330000u >> 15)) & 0x3f0u;
        const double2 e = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(t->logt) + off);
        const double invc = e.x, y0 = e.y;
        const double xd = __hiloint2double((int)((ix >> 3) + 0x38000000u), (int)(ix << 29));
        const double r = __fma_rn(xd, invc, -1.0);
        double y = __fma_rn(a0, r, a1);
        y = __fma_rn(y, r, a2);
        y = __fma_rn(y, r, 1.0);
        return __double2float_rn(__fma_rn(y, r, y0));
    }

    // std::max(a, b) of the reference: b only when a < b.
    static __device__ __forceinline__ float max_ref(float a, float b) { return (a < b) ? b : a; }

    __device__ __forceinline__ float finish(float mx, float sum) const
    {
        const float t1 = __fadd_rn(mx, log_sum(sum));
        return __double2float_rn(__dadd_rn((double)t1, -log2n));
    }

    // 2-D: neighbours in the reference's order (x0-1, x0+1, x1-1, x1+1); the reference sums
    // ((E(a) + E(b)) + E(c)) + E(d) with E(v) = expf(v - max).
    //
    // The largest neighbour's term is expf(0) = 1 exactly, so three libm replays are enough: a min/max
    // network finds the three losers (lo1, lo2, mid), their exponentials are evaluated, and selects put
    // the terms back at the reference's positions.  E(a) + E(b) = E(max(a,b)) + E(min(a,b)) because a float
    // add is commutative, so only the (c, d) pair needs its order restored.  Ties are harmless: a loser
    // equal to the maximum goes through exp_nonpos(0) = 1.  (NaN, which the reference's std::max chain
    // treats asymmetrically, is not reproduced -- as in exp_nonpos.)
    __device__ __forceinline__ float update4(float a, float b, float c, float d) const
    {
        const float hi1 = fmaxf(a, b), lo1 = fminf(a, b);
        const float hi2 = fmaxf(c, d), lo2 = fminf(c, d);
        const float mx = fmaxf(hi1, hi2), mid = fminf(hi1, hi2);
        const float e1 = exp_nonpos(__fsub_rn(lo1, mx));
        const float e2 = exp_nonpos(__fsub_rn(lo2, mx));
        const float em = exp_nonpos(__fsub_rn(mid, mx));
        const bool p = hi1 < hi2;            // the maximum is in the (c, d) pair
        const float eh1 = p ? em : 1.0f;     // E(max(a, b))
        const float eh2 = p ? 1.0f : em;     // E(max(c, d))
        const bool q = c < d;
        const float ec = q ? e2 : eh2;
        const float ed = q ? eh2 : e2;
        float s = __fadd_rn(eh1, e1);
        s = __fadd_rn(s, ec);
        s = __fadd_rn(s, ed);
        return finish(mx, s);
    }

    // 3-D: (x0-1, x0+1, x1-1, x1+1, x2-1, x2+1).
    __device__ __forceinline__ float update6(float a, float b, float c, float d, float e, float f) const
    {
        float mx = max_ref(a, b);
        mx = max_ref(mx, c);
        mx = max_ref(mx, d);
        mx = max_ref(mx, e);
        mx = max_ref(mx, f);
        float s = __fadd_rn(exp_nonpos(__fsub_rn(a, mx)), exp_nonpos(__fsub_rn(b, mx)));
        s = __fadd_rn(s, exp_nonpos(__fsub_rn(c, mx)));
        s = __fadd_rn(s, exp_nonpos(__fsub_rn(d, mx)));
        s = __fadd_rn(s, exp_nonpos(__fsub_rn(e, mx)));
        s = __fadd_rn(s, exp_nonpos(__fsub_rn(f, mx)));
        return finish(mx, s);
    }
};

struct FastMath {
    float ln2n;  // (float) log(2n)
    static constexpr bool kUsesTables = false;

    __device__ __forceinline__ void bind(const MathTables *) {}

    static __device__ __forceinline__ float ex2(float x)
    {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float lg2(float x)
    {
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    // 2^((v - mx) * log2(e) + off).  The difference is taken first and exactly (FADD of nearby values):
    // folding it into the FFMA as v * log2(e) - mx * log2(e) rounds at the magnitude of |u| and leaves a
    // bias per update that accumulates with the distance from the goal (0.14 at u = -2000 on maze.png).
    static __device__ __forceinline__ float e(float v, float mx, float off)
    {
        return ex2(__fmaf_rn(__fsub_rn(v, mx), 1.4426950408889634f, off));
    }
    __device__ __forceinline__ float finish(float mx, float s) const
    {
        return __fsub_rn(__fmaf_rn(lg2(s), 0.6931471805599453f, mx), ln2n);
    }
    // The largest neighbour contributes exactly 2^0 = 1, so only the other three need an ex2: a
    // min/max network separates them (6 FMNMX) and saves one of the five MUFU operations per update --
    // the MUFU pipe (16 lanes per SM and clock) is this mode's tightest resource.
    //
    // finish() keeps the reference's two rounding points at the magnitude of |u| -- float(mx + log(sum))
    // and then "- log(2n)" -- on purpose.  Deep in a corridor u falls by a constant per cell, the
    // roundings at |u| ~ 2^11 are as large as 1.2e-4 and repeat systematically from cell to cell, so an
    // update with a different rounding sequence (e.g. folding log(2n) into the sum as sum/4, one rounding
    // less) drifts away from the reference linearly with the distance from the goal: 0.14 at u = -2048
    // on maps/maze.png, 14x the stated tolerance.  With the same rounding points the two agree to a few
    // ulps (tests/test_parity_gpu.py::test_fast_mode_within_stated_tolerance).
    __device__ __forceinline__ float update4(float a, float b, float c, float d) const
    {
        const float hi1 = fmaxf(a, b), lo1 = fminf(a, b);
        const float hi2 = fmaxf(c, d), lo2 = fminf(c, d);
        const float mx = fmaxf(hi1, hi2), mid = fminf(hi1, hi2);
        const float s = __fadd_rn(__fadd_rn(e(lo1, mx, 0.0f), e(lo2, mx, 0.0f)), __fadd_rn(e(mid, mx, 0.0f), 1.0f));
        return finish(mx, s);
    }
    __device__ __forceinline__ float update6(float a, float b, float c, float d, float g, float f) const
    {
        const float mx = fmaxf(fmaxf(fmaxf(a, b), fmaxf(c, d)), fmaxf(g, f));
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(e(a, mx, 0.0f), e(b, mx, 0.0f)), __fadd_rn(e(c, mx, 0.0f), e(d, mx, 0.0f))),
                                  __fadd_rn(e(g, mx, 0.0f), e(f, mx, 0.0f)));
        return finish(mx, s);
    }
};

}  // namespace epic_b200
