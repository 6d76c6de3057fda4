/*
 * oracle/harmonic_oracle.c -- TEST INFRASTRUCTURE ONLY (see harmonic_oracle.h).
 *
 * "parity pinned": checked against the untouched reference sources compiled
 * into oracle/_ref/ and against tests/golden/ (tests/test_oracle.py).
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off [-fopenmp] -fPIC -shared (oracle/Makefile).
 * No -march / -ffast-math: the reference Makefile builds with plain -O3
 * (libepic/Makefile:2), so on x86-64 every float multiply and add below is a
 * separate IEEE operation.  The transcendental functions are glibc's libm
 * (expf, logf, log, sqrt) exactly as the reference calls them.
 */
#include "harmonic_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline float maxf_ref(float a, float b)
{
    /* std::max(a, b): returns b only if a < b (harmonic_cpu.cpp:61-63). */
    return (a < b) ? b : a;
}

/* One 2-D row.  Reference: harmonic_cpu.cpp:46-77.
 * Active columns: x1 = 1 + offset, +2, with offset = (it%2) != (x0%2). */
static float sweep_row_2d(const OracleHarmonic *h, uint64_t x0, int check)
{
    const uint64_t m1 = h->m[1];
    float *u = h->u;
    const uint32_t *locked = h->locked;
    const double log2n = log(2.0 * h->n);          /* :70, evaluated in double */
    float delta = 0.0f;
    uint64_t offset = (uint64_t)((h->iteration % 2) != (x0 % 2));

    for (uint64_t x1 = 1 + offset; x1 + 1 < m1; x1 += 2) {
        const uint64_t c = x0 * m1 + x1;
        if (locked[c]) {
            continue;
        }
        const float prev = u[c];
        const float up = u[c - m1], down = u[c + m1], left = u[c - 1], right = u[c + 1];
        float mx = maxf_ref(up, down);
        mx = maxf_ref(mx, left);
        mx = maxf_ref(mx, right);
        /* float adds left to right (:65-69), then float + float, then the
         * subtraction in double, then the store narrows to float (:65-70). */
        const float sum = expf(up - mx) + expf(down - mx) + expf(left - mx) + expf(right - mx);
        const float t = mx + logf(sum);
        u[c] = (float)((double)t - log2n);
        if (check) {
            const float d = fabsf(prev - u[c]);    /* :74 */
            delta = maxf_ref(delta, d);
        }
    }
    return delta;
}

/* One 3-D (x0, x1) pencil.  Reference: harmonic_cpu.cpp:90-131. */
static float sweep_pencil_3d(const OracleHarmonic *h, uint64_t x0, uint64_t x1, int check)
{
    const uint64_t m1 = h->m[1], m2 = h->m[2];
    const uint64_t s0 = m1 * m2, s1 = m2;
    float *u = h->u;
    const uint32_t *locked = h->locked;
    const double log2n = log(2.0 * h->n);
    float delta = 0.0f;
    uint64_t offset = (uint64_t)((h->iteration % 2) != (x0 % 2));
    if (x1 % 2 == 0) {
        offset = !offset;                          /* :97-99 */
    }
    for (uint64_t x2 = 1 + offset; x2 + 1 < m2; x2 += 2) {
        const uint64_t c = x0 * s0 + x1 * s1 + x2;
        if (locked[c]) {
            continue;
        }
        const float prev = u[c];
        const float a0 = u[c - s0], a1 = u[c + s0];
        const float b0 = u[c - s1], b1 = u[c + s1];
        const float c0 = u[c - 1], c1 = u[c + 1];
        float mx = maxf_ref(a0, a1);
        mx = maxf_ref(mx, b0);
        mx = maxf_ref(mx, b1);
        mx = maxf_ref(mx, c0);
        mx = maxf_ref(mx, c1);
        const float sum = expf(a0 - mx) + expf(a1 - mx) + expf(b0 - mx) + expf(b1 - mx) +
                          expf(c0 - mx) + expf(c1 - mx);
        const float t = mx + logf(sum);
        u[c] = (float)((double)t - log2n);
        if (check) {
            const float d = fabsf(prev - u[c]);
            delta = maxf_ref(delta, d);
        }
    }
    return delta;
}

void oracle_sweep(OracleHarmonic *h, int check)
{
    float delta = 0.0f;
    const int threads = h->threads > 1 ? h->threads : 1;
    (void)threads;

    if (h->n == 2) {
        const int64_t rows = (int64_t)h->m[0] - 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(max : delta) num_threads(threads) if (threads > 1)
#endif
        for (int64_t x0 = 1; x0 < rows; x0++) {
            const float d = sweep_row_2d(h, (uint64_t)x0, check);
            delta = maxf_ref(delta, d);
        }
    } else if (h->n == 3) {
        const int64_t planes = (int64_t)h->m[0] - 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(max : delta) num_threads(threads) if (threads > 1)
#endif
        for (int64_t x0 = 1; x0 < planes; x0++) {
            for (uint64_t x1 = 1; x1 + 1 < h->m[1]; x1++) {
                const float d = sweep_pencil_3d(h, (uint64_t)x0, x1, check);
                delta = maxf_ref(delta, d);
            }
        }
    }
    /* n == 4 is an empty branch in the reference (harmonic_cpu.cpp:193-195). */

    if (check) {
        h->delta = delta;                          /* :40-42, :74 */
    }
}

int oracle_update(OracleHarmonic *h)
{
    oracle_sweep(h, 0);
    h->iteration++;
    return ORACLE_SUCCESS;
}

int oracle_update_and_check(OracleHarmonic *h)
{
    oracle_sweep(h, 1);
    h->iteration++;
    return (h->delta < h->epsilon) ? ORACLE_SUCCESS_AND_CONVERGED : ORACLE_SUCCESS;
}

static int oracle_valid(const OracleHarmonic *h)
{
    return h != NULL && h->u != NULL && h->locked != NULL && !(h->epsilon <= 0.0) &&
           h->stagger != 0;
}

int oracle_complete(OracleHarmonic *h)
{
    if (!oracle_valid(h)) {
        return ORACLE_ERROR_INVALID_DATA;          /* harmonic_cpu.cpp:141-145 */
    }
    uint64_t m_max = 0;
    for (uint32_t i = 0; i < h->n; i++) {
        m_max = h->m[i] > m_max ? h->m[i] : m_max;
    }
    h->iteration = 0;
    h->delta = h->epsilon + 1.0f;
    int result = ORACLE_SUCCESS;
    /* :158-174 -- a plain update resets `result`, so the loop can only leave
     * right after a check sweep. */
    while (result != ORACLE_SUCCESS_AND_CONVERGED || h->iteration < m_max) {
        if (h->iteration % h->stagger == 0) {
            result = oracle_update_and_check(h);
        } else {
            result = oracle_update(h);
        }
    }
    return ORACLE_SUCCESS;
}

int oracle_run_iterations(OracleHarmonic *h, uint32_t count)
{
    if (!oracle_valid(h)) {
        return ORACLE_ERROR_INVALID_DATA;
    }
    for (uint32_t i = 0; i < count; i++) {
        if (h->iteration % h->stagger == 0) {
            oracle_update_and_check(h);
        } else {
            oracle_update(h);
        }
    }
    return ORACLE_SUCCESS;
}

int oracle_set_cells_2d(OracleHarmonic *h, uint32_t k, const uint32_t *v, const uint32_t *types)
{
    if (h == NULL || h->n == 0 || h->u == NULL || h->locked == NULL || k == 0 || v == NULL ||
        types == NULL) {
        return ORACLE_ERROR_INVALID_DATA;          /* harmonic_utilities_cpu.cpp:40-45 */
    }
    for (uint32_t i = 0; i < k; i++) {
        const uint64_t x = v[2 * i + 0], y = v[2 * i + 1];
        if (y >= h->m[0] || x >= h->m[1]) {
            continue;                              /* warn + skip, :51-56 */
        }
        const uint64_t c = y * h->m[1] + x;
        switch (types[i]) {
        case 0: h->u[c] = 0.0f;  h->locked[c] = 1; break;   /* goal */
        case 1: h->u[c] = -1e6f; h->locked[c] = 1; break;   /* obstacle */
        case 2: h->u[c] = -1e6f; h->locked[c] = 0; break;   /* free */
        default: break;                            /* warn + skip, :67-72 */
        }
    }
    return ORACLE_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* Streamlines.  Reference: libepic/src/harmonic/harmonic_path_cpu.cpp        */

static inline uint32_t f2u(float f)
{
    /* (unsigned int)f as gcc emits it on x86-64: cvttss2si to 64 bits, keep
     * the low 32.  Spelled out so that negative inputs wrap the same way. */
    return (uint32_t)(int64_t)f;
}

static int cell_is_blocked(const OracleHarmonic *h, uint32_t xc, uint32_t yc)
{
    if (xc >= h->m[1] || yc >= h->m[0]) {
        return 1;
    }
    const uint64_t c = (uint64_t)yc * h->m[1] + xc;
    return h->locked[c] == 1 && h->u[c] < 0.0f;    /* obstacle: locked and negative */
}

int oracle_potential_2d(const OracleHarmonic *h, float x, float y, float *potential)
{
    if (h == NULL || h->u == NULL || h->locked == NULL) {
        return ORACLE_ERROR_INVALID_DATA;
    }
    if (cell_is_blocked(h, f2u(x + 0.5f), f2u(y + 0.5f))) {
        return ORACLE_ERROR_INVALID_LOCATION;      /* :52-58 */
    }
    const uint64_t m1 = h->m[1];
    const uint32_t xl = f2u(x - 0.5f), xr = f2u(x + 0.5f);
    const uint32_t yt = f2u(y - 0.5f), yb = f2u(y + 0.5f);
    if (xl >= h->m[1] || xr >= h->m[1] || yt >= h->m[0] || yb >= h->m[0]) {
        /* The reference reads out of bounds here (undefined behaviour); the
         * oracle reports it instead. */
        return ORACLE_ERROR_INVALID_LOCATION;
    }
    const float alpha = x - (float)xl;             /* :72-73 */
    const float beta = y - (float)yt;
    const float one = (1.0f - alpha) * h->u[yt * m1 + xl] + alpha * h->u[yt * m1 + xr];
    const float two = (1.0f - alpha) * h->u[yb * m1 + xl] + alpha * h->u[yb * m1 + xr];
    *potential = (1.0f - beta) * one + beta * two; /* :75-79 */
    return ORACLE_SUCCESS;
}

int oracle_gradient_2d(const OracleHarmonic *h, float x, float y, float cd, float *px, float *py)
{
    if (h == NULL || h->u == NULL || h->locked == NULL) {
        return ORACLE_ERROR_INVALID_DATA;
    }
    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
    int result = oracle_potential_2d(h, x - cd, y, &v0);
    result += oracle_potential_2d(h, x + cd, y, &v1);
    result += oracle_potential_2d(h, x, y - cd, &v2);
    result += oracle_potential_2d(h, x, y + cd, &v3);
    if (result != ORACLE_SUCCESS) {
        return ORACLE_ERROR_INVALID_GRADIENT;      /* :104-108 */
    }
    float gx = (v1 - v0) / (2.0f * cd);
    float gy = (v3 - v2) / (2.0f * cd);
    /* std::pow(float, int) promotes to double; sqrt in double (:113). */
    const float denom = (float)sqrt((double)gx * (double)gx + (double)gy * (double)gy);
    gx /= denom;
    gy /= denom;
    *px = gx;
    *py = gy;
    return ORACLE_SUCCESS;
}

static int path_is_stuck(const float *p, uint64_t n, float step)
{
    /* p holds n floats = n/2 points (:121-151). */
    if (n % 2 == 1) {
        return 1;
    }
    if (n == 0) {
        return 0;
    }
    const float x = p[n - 2], y = p[n - 1];
    int64_t lo = (int64_t)n - 2 * 5 - 2;
    if (lo < 0) {
        lo = 0;
    }
    for (uint64_t i = n - 2; i > (uint64_t)lo; i -= 2) {
        const float dx = x - p[i - 2], dy = y - p[i - 1];
        const float dist = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
        if (dist < step / 2.0f) {
            return 1;
        }
    }
    return 0;
}

int oracle_path_2d(const OracleHarmonic *h, float x, float y, float step, float cd,
                   uint32_t max_length, uint32_t *k, float **path)
{
    if (h == NULL || h->u == NULL || h->locked == NULL || k == NULL || path == NULL ||
        *path != NULL) {
        return ORACLE_ERROR_INVALID_DATA;
    }
    uint32_t xc = f2u(x + 0.5f), yc = f2u(y + 0.5f);
    if (cell_is_blocked(h, xc, yc)) {
        return ORACLE_ERROR_INVALID_LOCATION;      /* :168-175 */
    }
    uint64_t cap = 1024, n = 0;
    float *p = (float *)malloc(cap * sizeof(float));
    if (p == NULL) {
        return ORACLE_ERROR_INVALID_DATA;
    }
    p[n++] = x;
    p[n++] = y;
    while (h->locked[(uint64_t)yc * h->m[1] + xc] != 1 && !path_is_stuck(p, n, step) &&
           n < 2ull * max_length) {
        float gx = 0.0f, gy = 0.0f;
        if (oracle_gradient_2d(h, x, y, cd, &gx, &gy) != ORACLE_SUCCESS) {
            free(p);
            return ORACLE_ERROR_INVALID_GRADIENT;  /* :190-193 */
        }
        x += gx * step;                            /* ascent on the log-potential */
        y += gy * step;
        if (n + 2 > cap) {
            cap *= 2;
            float *q = (float *)realloc(p, cap * sizeof(float));
            if (q == NULL) {
                free(p);
                return ORACLE_ERROR_INVALID_DATA;
            }
            p = q;
        }
        p[n++] = x;
        p[n++] = y;
        xc = f2u(x + 0.5f);
        yc = f2u(y + 0.5f);
        if (xc >= h->m[1] || yc >= h->m[0]) {
            break;                                 /* reference would read out of bounds */
        }
    }
    if (n / 2 <= 2) {
        free(p);
        return ORACLE_ERROR_INVALID_PATH;          /* :207-210 */
    }
    *k = (uint32_t)(n / 2);
    *path = p;
    return ORACLE_SUCCESS;
}

void oracle_free_path(float *path)
{
    free(path);
}
