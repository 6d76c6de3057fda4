// math_policies.cuh -- the two arithmetic modes of the log-sum-exp update.
//
//   StrictMath  replays glibc's expf/logf in double arithmetic (strict_math.h) and performs the
//               reference CPU path's float adds, its double `- log(2n)` and its narrowing store
//               in the same order (reference libepic/src/harmonic/harmonic_cpu.cpp:58-70,
//               :107-124), so the field is bit-identical to harmonic_complete_cpu's.
//   FastMath    MUFU ex2/lg2 approximations, the arithmetic class of the reference's own GPU
//               kernel (harmonic_gpu.cu:52-61: __expf, __logf, constant 1.38629436f).  Stated
//               tolerance against the CPU path: |du| <= 1e-5*|u| + 1e-5 at equal iteration count
//               (tests/test_parity_gpu.py).
//
// This translation unit is compiled with -fmad=false: every multiply-add that may fuse is written
// as an explicit fma intrinsic.
#pragma once

#include "strict_math.h"

namespace epic_b200 {

// Shared-memory image of the libm tables, split into 32-bit words so that a warp's 32 random
// lookups hit 32 distinct banks (or broadcast): conflict-free LDS.32.
struct MathTables {
    uint32_t exp_hi[32];
    uint32_t exp_lo[32];
    uint32_t invc_hi[16];
    uint32_t invc_lo[16];
    uint32_t logc_hi[16];
    uint32_t logc_lo[16];
};

static __device__ __constant__ uint64_t c_exp2f_table[32] = {EPIC_EXP2F_TABLE};
static __device__ __constant__ double c_logf_table[32] = {EPIC_LOGF_TABLE};

__device__ __forceinline__ void load_math_tables(MathTables *t, int tid, int nthreads)
{
    for (int i = tid; i < 32; i += nthreads) {
        t->exp_hi[i] = (uint32_t)(c_exp2f_table[i] >> 32);
        t->exp_lo[i] = (uint32_t)c_exp2f_table[i];
    }
    for (int i = tid; i < 16; i += nthreads) {
        const uint64_t a = (uint64_t)__double_as_longlong(c_logf_table[2 * i]);
        const uint64_t b = (uint64_t)__double_as_longlong(c_logf_table[2 * i + 1]);
        t->invc_hi[i] = (uint32_t)(a >> 32);
        t->invc_lo[i] = (uint32_t)a;
        t->logc_hi[i] = (uint32_t)(b >> 32);
        t->logc_lo[i] = (uint32_t)b;
    }
}

struct StrictMath {
    const MathTables *t;
    double log2n;  // glibc log(2.0 * n)

    __device__ __forceinline__ void bind(const MathTables *tables) { t = tables; }

    __device__ __forceinline__ float exp_nonpos(float x) const
    {
        // strict_expf_nonpos with the table split into words (see strict_math.h for the algorithm)
        if (x < -0x1.9fe368p6f) {
            return 0.0f;
        }
        if (x < -0x1.9d1d9ep6f) {
            return 0x1p-149f;
        }
        if (x != x) {
            return x + x;
        }
        const double xd = (double)x;
        const double kdp = __fma_rn(kExpInvLn2N, xd, kExpShift);
        const uint32_t ki = (uint32_t)__double2loint(kdp);
        const double kd = __dadd_rn(kdp, -kExpShift);
        const double r = __fma_rn(kExpInvLn2N, xd, -kd);
        const uint32_t idx = ki & 31u;
        const double s = __hiloint2double((int)(t->exp_hi[idx] + (ki << 15)), (int)t->exp_lo[idx]);
        const double z = __fma_rn(kExpC0, r, kExpC1);
        const double r2 = __dmul_rn(r, r);
        double y = __fma_rn(kExpC2, r, 1.0);
        y = __fma_rn(z, r2, y);
        return __double2float_rn(__dmul_rn(y, s));
    }

    __device__ __forceinline__ float log_sum(float x) const
    {
        const uint32_t ix = __float_as_uint(x);
        if (ix == 0x3f800000u) {
            return 0.0f;
        }
        const uint32_t tmp = ix - 0x3f330000u;
        const uint32_t i = (tmp >> 19) & 15u;
        const int k = (int)tmp >> 23;
        const uint32_t iz = ix - (tmp & 0xff800000u);
        const double invc = __hiloint2double((int)t->invc_hi[i], (int)t->invc_lo[i]);
        const double logc = __hiloint2double((int)t->logc_hi[i], (int)t->logc_lo[i]);
        const double z = (double)__uint_as_float(iz);
        const double r = __fma_rn(z, invc, -1.0);
        const double y0 = __fma_rn((double)k, kLogLn2, logc);
        const double r2 = __dmul_rn(r, r);
        double y = __fma_rn(kLogA1, r, kLogA2);
        y = __fma_rn(kLogA0, r2, y);
        y = __fma_rn(y, r2, __dadd_rn(y0, r));
        return __double2float_rn(y);
    }

    // std::max(a, b) of the reference: b only when a < b.
    static __device__ __forceinline__ float max_ref(float a, float b) { return (a < b) ? b : a; }

    __device__ __forceinline__ float finish(float mx, float sum) const
    {
        const float t1 = __fadd_rn(mx, log_sum(sum));
        return __double2float_rn(__dadd_rn((double)t1, -log2n));
    }

    // 2-D: neighbours in the reference's order (x0-1, x0+1, x1-1, x1+1).
    __device__ __forceinline__ float update4(float a, float b, float c, float d) const
    {
        float mx = max_ref(a, b);
        mx = max_ref(mx, c);
        mx = max_ref(mx, d);
        float s = __fadd_rn(exp_nonpos(__fsub_rn(a, mx)), exp_nonpos(__fsub_rn(b, mx)));
        s = __fadd_rn(s, exp_nonpos(__fsub_rn(c, mx)));
        s = __fadd_rn(s, exp_nonpos(__fsub_rn(d, mx)));
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
    static __device__ __forceinline__ float e(float v, float mx)
    {
        return ex2(__fmul_rn(__fsub_rn(v, mx), 1.4426950408889634f));
    }
    __device__ __forceinline__ float finish(float mx, float s) const
    {
        return __fsub_rn(__fmaf_rn(lg2(s), 0.6931471805599453f, mx), ln2n);
    }
    __device__ __forceinline__ float update4(float a, float b, float c, float d) const
    {
        const float mx = fmaxf(fmaxf(a, b), fmaxf(c, d));
        const float s = __fadd_rn(__fadd_rn(e(a, mx), e(b, mx)), __fadd_rn(e(c, mx), e(d, mx)));
        return finish(mx, s);
    }
    __device__ __forceinline__ float update6(float a, float b, float c, float d, float g, float f) const
    {
        const float mx = fmaxf(fmaxf(fmaxf(a, b), fmaxf(c, d)), fmaxf(g, f));
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(e(a, mx), e(b, mx)), __fadd_rn(e(c, mx), e(d, mx))),
                                  __fadd_rn(e(g, mx), e(f, mx)));
        return finish(mx, s);
    }
};

}  // namespace epic_b200
