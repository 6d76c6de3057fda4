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
    // the three-expf form of the 2-D update (StrictMath::update4) against the reference's four-expf form, both
    // on the host libm: random neighbourhoods drawn from the values a field holds (relaxed potentials, the
    // -1e6 of obstacles and unreached cells, goals at 0), with ties, signed zeros and near-ties forced in
    unsigned long long bad_u = 0, n_u = 0;
    {
        auto e = [](float x) { return expf(x); };
        auto l = [](float x) { return logf(x); };
        uint64_t st = 0x9e3779b97f4a7c15ull;
        auto rnd = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
        const float special[8] = {0.0f, -0.0f, -1e6f, -1e6f, -0x1p-149f, -104.25f, -3.5f, -969.5537f};
        for (unsigned long long it = 0; it < 40000000ull / stride; ++it) {
            float v[4];
            const float base = -(float)(rnd() % 2000) * 0.5f;
            for (int i = 0; i < 4; ++i) {
                const uint64_t r = rnd();
                switch (r & 7) {
                case 0: v[i] = special[(r >> 3) & 7]; break;
                case 1: v[i] = base; break;                                            // exact tie with the base
                case 2: v[i] = strict_from_fbits(strict_fbits(base) + (uint32_t)((r >> 3) & 3)); break;  // a few ulps off
                default: v[i] = base - (float)((r >> 8) & 0xffff) * (1.0f / 4096.0f) * (float)(1 + ((r >> 3) & 15)); break;
                }
            }
            const float want = strict_update4_reference(v[0], v[1], v[2], v[3], e, l);
            const float got = strict_update4_network(v[0], v[1], v[2], v[3], e, l);
            n_u++;
            if (strict_fbits(want) != strict_fbits(got) && !(want != want && got != got)) {
                if (bad_u < 5) printf("update4(%a, %a, %a, %a): got %a want %a\n", v[0], v[1], v[2], v[3], got, want);
                bad_u++;
            }
        }
    }
    printf("expf mismatches %llu of %llu\nlogf mismatches %llu of %llu\nupdate4 mismatches %llu of %llu\n", bad_e, n_e, bad_l,
           n_l, bad_u, n_u);
    return (bad_e || bad_l || bad_u) ? 1 : 0;
}
