// grid.h -- one harmonic grid held by ONE process on one or several B200s.
//
// This is what the libepic C ABI's four device handles (Harmonic::d_m / d_u / d_locked / d_delta) stand for.
// The reference keeps a grid on device 0 only (reference libepic/src/harmonic/harmonic_model_gpu.cu:34-204,
// harmonic_gpu.cu:168-434: no cudaSetDevice anywhere) and its callers are single-process programs (the
// nav_core plugin, src/epic_nav_core_plugin.cpp:256; the anytime node,
// src/epic_navigation_node_harmonic.cpp:178-201; the Python wrapper), so the way for them to use more than one
// GPU *unchanged* is a library that shards behind the ABI:
//
//   EPIC_DEVICES=0,1,2,3 | all | <count>      (unset: one slab on the current / EPIC_DEVICE device)
//
// A Grid cuts the grid into row slabs (2-D) / x0-slabs (3-D), one `Field` per listed device (a device may be
// listed more than once), ghost depth = sweeps per pass, and wires neighbouring slabs with
// Field::set_peer_local: halos are NVLink peer stores from inside the sweep kernel, passes are ordered between
// GPUs by in-kernel acquire / release flags (2-D) or stream memory operations (3-D).  Nothing crosses the host
// per pass.  The convergence check's max over slabs is an all-reduce done by the slabs' decision kernels over
// peer memory (decide_all_kernel), so solve() queues whole solver periods ahead of the devices exactly as the
// single-slab solve does; each slab is driven by its own host thread.
//
// Red-black ordering makes the result independent of the partition: fields, deltas and iteration counts are
// bit-identical to the single-GPU (and the reference CPU) ones.
//
// Return codes are the reference's (libepic/include/epic/error_codes.h:31-46).
#pragma once

#include <stdint.h>

#include <vector>

#include "field.h"

namespace epic_b200 {

struct GridStats {           // EPIC_VERBOSE / epic_b200 introspection
    uint32_t slabs = 0;
    uint64_t launches = 0;
    uint64_t skipped_tiles = 0;
    uint64_t skipped_by_slab[kMaxSlabs] = {0};
    double last_solve_seconds = 0.0;
    uint32_t last_solve_iterations = 0;
    float last_solve_delta = 0.0f;
};

class Grid {
public:
    // n = 2 or 3; gm = dimensions.  cfg.devices / cfg.ndevices select the slabs (see above).
    static int create(Grid **out, unsigned n, const uint64_t *gm, const FieldConfig &cfg);
    ~Grid();

    int slabs() const { return (int)slabs_.size(); }
    Field *slab(int i) { return slabs_[(size_t)i]; }
    int sweeps_per_pass() const { return slabs_[0]->sweeps_per_pass(); }

    // Whole-grid dense host arrays (the reference's layout).  Complete on return.
    int upload_u(const float *host);
    int upload_locked(const uint32_t *host);
    int download_u(float *host);

    // `count` half-sweeps from iteration it0 on every slab (asynchronous); read_delta = max over the slabs of
    // the last check sweep's delta (synchronises).
    int run(uint32_t it0, uint32_t count, bool check_last);
    int read_delta(float *delta);
    // harmonic_execute_gpu's loop (reference harmonic_gpu.cu:266-290).
    int solve(float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iteration, float *delta);

    int set_cells_2d(uint32_t k, const uint32_t *v, const uint32_t *types);
    int ingest_occupancy_2d(const signed char *host, int threshold, int no_change);
    int reset_free_cells_2d();

    int potential_2d(float x, float y, float *value);
    int gradient_2d(float x, float y, float cd, float *px, float *py);
    int paths_2d(uint32_t count, const float *starts, float step, float cd, uint32_t max_length, int *ret,
                 uint32_t *k, float **paths);

    int sync();
    GridStats stats() const;

private:
    Grid() {}
    int solve_sharded(float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iteration, float *delta);
    int decide_all(size_t slab, uint32_t tag, uint32_t count, bool rule);
    // held x0-range of slab i (owned + ghost, clipped to the grid)
    void held(size_t i, uint64_t *first, uint64_t *layers) const;

    unsigned n_ = 0;
    uint64_t gm_[3] = {1, 1, 1};
    std::vector<Field *> slabs_;
    std::vector<cudaStream_t> streams_;        // one per distinct device, shared by the slabs on it
    std::vector<int> stream_device_;
    std::vector<unsigned long long *> inbox_;  // per slab, on its device: 2 x kDecideSlots words
    uint32_t tag_ = 0;                         // decision periods issued so far
    bool verbose_ = false;
    GridStats stats_;
};

}  // namespace epic_b200
