/*
 * oracle/harmonic_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of libepic's log-space harmonic relaxation and
 * streamline extraction.  It exists so that the CUDA product path can be
 * checked against an independent implementation of the reference algorithm.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call anything in oracle/.  The
 * product (epic_b200/, libepic.so) never includes, links or calls it.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned (tests/test_oracle.py)
 *   (a) against the reference's own sources compiled untouched into
 *       oracle/_ref/libepic_ref_cpu.so (oracle/Makefile, container only), and
 *   (b) against the JSON files in tests/golden, vectors produced by that reference build
 *       with tools/make_golden.py.
 *
 * Differences from the reference, on purpose: 64-bit cell counts and indices
 * (the reference overflows `unsigned int` at 65536^2 cells), and an optional
 * OpenMP row-parallel sweep (red-black makes the result independent of the
 * visiting order, so threading cannot change a single bit).
 */
#ifndef EPIC_B200_HARMONIC_ORACLE_H
#define EPIC_B200_HARMONIC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Return codes: reference libepic/include/epic/error_codes.h:31-46 */
enum {
    ORACLE_SUCCESS = 0,
    ORACLE_SUCCESS_AND_CONVERGED = 1,
    ORACLE_ERROR_INVALID_DATA = 2,
    ORACLE_ERROR_INVALID_LOCATION = 10,
    ORACLE_ERROR_INVALID_CELL_TYPE = 11,
    ORACLE_ERROR_INVALID_GRADIENT = 12,
    ORACLE_ERROR_INVALID_PATH = 13
};

/* Mirrors the user-visible half of `struct Harmonic`
 * (reference libepic/include/epic/harmonic/harmonic.h:44-64). */
typedef struct OracleHarmonic {
    uint32_t n;              /* 2 or 3 */
    uint64_t m[4];           /* size of each dimension, m[n-1] fastest */
    float *u;                /* log-potentials, row-major */
    uint32_t *locked;        /* 0 = free, non-zero = locked */
    float epsilon;
    float delta;
    uint32_t stagger;        /* numIterationsToStaggerCheck */
    uint32_t iteration;      /* currentIteration */
    int threads;             /* 0/1 = serial; >1 = OpenMP threads for the sweep */
} OracleHarmonic;

/* One red-black half-sweep (harmonic_cpu.cpp:38-78 for n=2, :81-133 for n=3).
 * If check != 0, h->delta is reset and receives max |u_prev - u_new| over the
 * cells this half-sweep touched.  Does NOT advance h->iteration. */
void oracle_sweep(OracleHarmonic *h, int check);

/* harmonic_update_cpu (harmonic_cpu.cpp:187-200) */
int oracle_update(OracleHarmonic *h);
/* harmonic_update_and_check_cpu (harmonic_cpu.cpp:203-220) */
int oracle_update_and_check(OracleHarmonic *h);
/* harmonic_complete_cpu (harmonic_cpu.cpp:136-184) */
int oracle_complete(OracleHarmonic *h);
/* Exactly `count` iterations with the complete() schedule (check sweeps on
 * iteration % stagger == 0), no termination test.  Used for fixed-K parity. */
int oracle_run_iterations(OracleHarmonic *h, uint32_t count);

/* harmonic_utilities_set_cells_2d_cpu (harmonic_utilities_cpu.cpp:38-76) */
int oracle_set_cells_2d(OracleHarmonic *h, uint32_t k, const uint32_t *v, const uint32_t *types);

/* harmonic_compute_potential_2d_cpu (harmonic_path_cpu.cpp:41-82) */
int oracle_potential_2d(const OracleHarmonic *h, float x, float y, float *potential);
/* harmonic_compute_gradient_2d_cpu (harmonic_path_cpu.cpp:85-118) */
int oracle_gradient_2d(const OracleHarmonic *h, float x, float y, float cd, float *px, float *py);
/* harmonic_compute_path_2d_cpu (harmonic_path_cpu.cpp:154-221).  *path is
 * malloc'ed ([x0,y0,x1,y1,...], *k points); release with oracle_free_path. */
int oracle_path_2d(const OracleHarmonic *h, float x, float y, float step, float cd,
                   uint32_t max_length, uint32_t *k, float **path);
void oracle_free_path(float *path);

#ifdef __cplusplus
}
#endif
#endif
