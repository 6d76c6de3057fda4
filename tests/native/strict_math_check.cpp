// Host-side check of epic_b200/csrc/kernels/strict_math.h against the host libm.
// usage: strict_math_check [stride]   (stride 1 = every float in the ranges the sweep can produce)
// Prints "expf mismatches M of N" / "logf mismatches M of N"; exit status 0 iff both are 0.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../epic_b200/csrc/kernels/strict_math.h"

using namespace epic_b200;
static const uint64_t kT[32] = {EPIC_EXP2F_TABLE};
static const double kL[32] = {EPIC_LOGF_TABLE};

int main(int argc, char **argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    unsigned long long bad_e = 0, n_e = 0, bad_l = 0, n_l = 0;
    // every x <= 0 from -0.0 down to -inf: bit patterns 0x80000000 .. 0xff800000
#pragma omp parallel for reduction(+ : bad_e, n_e) schedule(static)
    for (long long b = 0x80000000ll; b <= 0xff800000ll; b += stride) {
        const float x = strict_from_fbits((uint32_t)b);
        const float a = strict_expf_nonpos(x, kT), e = expf(x);
        n_e++;
        if (strict_fbits(a) != strict_fbits(e)) {
            if (bad_e < 5) printf("expf(%a): got %a want %a\n", x, a, e);
            bad_e++;
        }
    }
    if (strict_fbits(strict_expf_nonpos(0.0f, kT)) != strict_fbits(expf(0.0f))) {  // +0 (a == max neighbour)
        printf("expf(+0) differs\n");
        bad_e++;
    }
    // every normal x in [2^-3, 2^4): covers the sweep's [1, 6]
#pragma omp parallel for reduction(+ : bad_l, n_l) schedule(static)
    for (long long b = 0x3e000000ll; b < 0x41800000ll; b += stride) {
        const float x = strict_from_fbits((uint32_t)b);
        const float a = strict_logf_normal(x, kL), e = logf(x);
        n_l++;
        if (strict_fbits(a) != strict_fbits(e)) {
            if (bad_l < 5) printf("logf(%a): got %a want %a\n", x, a, e);
            bad_l++;
        }
    }
    // the sweep's own form of logf (tables indexed by exponent and interval together): every x in [1, 8)
#pragma omp parallel for reduction(+ : bad_l, n_l) schedule(static)
    for (long long b = 0x3f800000ll; b < 0x41000000ll; b += stride) {
        const float x = strict_from_fbits((uint32_t)b);
        const float a = strict_logf_sum(x, kL), e = logf(x);
        n_l++;
        if (strict_fbits(a) != strict_fbits(e)) {
            if (bad_l < 5) printf("logf_sum(%a): got %a want %a\n", x, a, e);
            bad_l++;
        }
    }
    printf("expf mismatches %llu of %llu\nlogf mismatches %llu of %llu\n", bad_e, n_e, bad_l, n_l);
    return (bad_e || bad_l) ? 1 : 0;
}
