/*
 * epic_b200.h -- slab-level C ABI of libepic.so (extension; the reference has no multi-GPU API).
 *
 * libepic's own ABI (epic/libepic.h) describes ONE grid held by ONE device.  To shard a grid over
 * several B200s (one process per GPU) each rank holds a *slab*: the x0-range [row0, row0+rows) of
 * the global grid plus `ghost` x0-layers on each side that mirror the neighbouring ranks' edge
 * layers.  This header exposes the slab object that libepic's entry points are themselves built on
 * (the reference functions it generalises: libepic/src/harmonic/harmonic_model_gpu.cu:34-204 for
 * residency, harmonic_gpu.cu:327-434 for update / update_and_check / get_potential_values,
 * harmonic_utilities_gpu.cu:66-138 for set_cells).  The max-reduction of delta over the ranks is the
 * caller's (epic_b200/sharded.py: one all-reduce per check sweep); ghost layers are refreshed either by the
 * caller (NCCL send/recv on the pointers epic_b200_field_layer_ptr returns) or by the library itself once the
 * neighbouring slabs have been introduced with the peer calls below (NVLink peer stores from inside the sweep
 * kernel).
 *
 * Plain C: pointers and sizes only.  Return values are libepic's error codes (0 = success).
 */
#ifndef EPIC_B200_H
#define EPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct epic_b200_field epic_b200_field;

enum { EPIC_B200_MATH_STRICT = 0, EPIC_B200_MATH_FAST = 1, EPIC_B200_MATH_ENV = -1 };

typedef struct epic_b200_info {
    uint64_t pitch;          /* floats per innermost row in device memory */
    uint64_t layer_floats;   /* floats per x0-layer in device memory */
    uint64_t launches;       /* kernels launched by this field so far */
    uint64_t device_bytes;   /* device memory held */
    uint32_t sweeps_per_pass;/* T: half-sweeps fused into one kernel launch */
    uint32_t tile_rows;      /* rows of the shared-memory tile (2-D) */
    uint32_t math;           /* EPIC_B200_MATH_* in effect */
    int32_t device;
    uint64_t skipped_tiles;  /* tiles the solves so far skipped as static (bit-identical results either way) */
} epic_b200_info;

/* n = 2 or 3; m[n] = GLOBAL dimensions; the slab owns x0 in [row0, row0+rows) and keeps `ghost`
 * extra layers per side (0 for a whole grid).  math: EPIC_B200_MATH_*; device: ordinal or -1 for the
 * current one; stream: a cudaStream_t to run on when use_stream != 0 (else a private stream). */
int epic_b200_field_create(epic_b200_field **out, unsigned int n, const uint64_t *m, uint64_t row0, uint64_t rows,
                           unsigned int ghost, int math, int device, void *stream, int use_stream);
void epic_b200_field_destroy(epic_b200_field *f);
int epic_b200_field_info(epic_b200_field *f, epic_b200_info *info);

/* Dense host arrays covering `layers` x0-layers from global layer `first` (owned or ghost). */
int epic_b200_field_upload_u(epic_b200_field *f, const float *host, uint64_t first, uint64_t layers);
int epic_b200_field_upload_locked(epic_b200_field *f, const uint32_t *host, uint64_t first, uint64_t layers);
int epic_b200_field_download_u(epic_b200_field *f, float *host, uint64_t first, uint64_t layers);
int epic_b200_field_download_locked(epic_b200_field *f, uint32_t *host, uint64_t first, uint64_t layers);

/* Enqueue `count` half-sweeps starting at iteration it0 (asynchronous).  With check_last the last
 * sweep accumulates max|du| over the owned cells of its colour; read_delta fetches and resets it
 * (synchronises the stream). */
int epic_b200_field_run(epic_b200_field *f, uint32_t it0, uint32_t count, int check_last);
int epic_b200_field_read_delta(epic_b200_field *f, float *delta);
/* The complete reference loop on a whole-grid field (ghost == 0, rows == m[0]). */
int epic_b200_field_solve(epic_b200_field *f, float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iterations,
                          float *delta);
int epic_b200_field_sync(epic_b200_field *f);
/* Static-tile skipping for the passes epic_b200_field_run issues from now on (2-D; field_solve uses it on its
 * own): a tile whose 3 x 3 neighbourhood saw no update change a value in the previous pass returns at once --
 * bit-identical results, less work on fields that are unreached or converged in places.  Tiles that read ghost
 * layers always run.  Leave it off for throughput measurements. */
int epic_b200_field_set_tracking(epic_b200_field *f, int on);

/* Device address of global layer `layer` in the buffer that currently holds the field (it changes
 * with every pass: query after each run).  Null when the layer is not held by this slab. */
void *epic_b200_field_layer_ptr(epic_b200_field *f, int64_t layer);

/* Peer-to-peer halos over NVLink (one process per GPU on one node).  peer_export fills an opaque blob
 * (EPIC_B200_PEER_BLOB_BYTES) holding CUDA IPC handles of this slab's buffers and flag words; the ranks
 * exchange blobs (any transport) and hand the neighbours' blobs to set_peer_ipc: dir 0 = the slab above
 * (lower x0), 1 = the slab below.  From then on every pass stores its edge layers straight into the
 * neighbours' ghost layers from inside the sweep kernel and signals completion with a stream-ordered flag
 * write; the next pass waits on the neighbours' flags.  No collective, no host involvement per pass.  All
 * slabs must issue the same sequence of passes, and callers must synchronise all ranks after uploads.
 * set_peer_local is the same for two slabs that live in one process. */
#define EPIC_B200_PEER_BLOB_BYTES 256
int epic_b200_field_peer_export(epic_b200_field *f, void *blob, uint64_t blob_bytes);
int epic_b200_field_set_peer_ipc(epic_b200_field *f, int dir, const void *blob, uint64_t blob_bytes);
int epic_b200_field_set_peer_local(epic_b200_field *f, int dir, epic_b200_field *other);

int epic_b200_field_set_cells_2d(epic_b200_field *f, uint32_t k, const uint32_t *v, const uint32_t *types);
int epic_b200_field_potential_2d(epic_b200_field *f, float x, float y, float *value);
int epic_b200_field_gradient_2d(epic_b200_field *f, float x, float y, float cd, float *px, float *py);
/* paths[i] is allocated with new float[2*k[i]]; release with epic_b200_free_path. */
int epic_b200_field_paths_2d(epic_b200_field *f, uint32_t count, const float *starts, float step, float cd,
                             uint32_t max_length, int *results, uint32_t *k, float **paths);
void epic_b200_free_path(float *path);

/* What the grid behind a libepic `Harmonic` struct (its d_* handles) looks like and did last: how many slabs the
 * library sharded it into (EPIC_DEVICES), kernels launched, the duration / iteration count / delta of the most recent
 * harmonic_execute_gpu solve (device time of the loop, without the copies around it), and how many tile passes each
 * slab skipped as static -- the load balance of a sharded solve.  `harmonic` is a `const epic::Harmonic *`. */
typedef struct epic_b200_stats {
    uint32_t slabs;
    uint32_t last_solve_iterations;
    float last_solve_delta;
    uint32_t reserved;
    double last_solve_seconds;
    uint64_t launches;
    uint64_t skipped_tiles[16];
} epic_b200_stats;
int epic_b200_harmonic_stats(const void *harmonic, epic_b200_stats *out);

/* Self-test: the strict-math device functions (bit-exact twins of glibc expf / logf, see
 * epic_b200/csrc/kernels/strict_math.h) against THIS host's libm, over every `stride`-th float of the
 * argument ranges the sweep can produce (x <= 0 for expf, [1, 8] for logf). */
int epic_b200_selftest_math(uint32_t stride, uint64_t *exp_checked, uint64_t *exp_mismatches, uint64_t *log_checked,
                            uint64_t *log_mismatches);

/* Library self-description: "epic_b200 <version> sm_100a". */
const char *epic_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EPIC_B200_H */
