// field.h -- device residency and sweep scheduling for one slab of a harmonic grid.
//
// A `Field` owns everything the reference keeps behind `Harmonic::d_m/d_u/d_locked/d_delta`
// (reference libepic/src/harmonic/harmonic_model_gpu.cu:34-204, harmonic_gpu.cu:205-223) for
// the x0-range [row0, row0+rows) of a grid, laid out for B200:
//
//   u[2]      two ping-pong buffers of (rows + 2*ghost) x0-layers, each innermost row padded to a
//             128-byte multiple ("pitch"); a pass reads one and writes the other, so tiles can
//             re-read their halos while their neighbours are being written.
//   freemask  1 bit per cell, 1 = the cell is not locked.  1/32 of the traffic of the reference's
//             uint32 `locked` array.  (The sweep additionally never touches the global border.)
//   ctrl      a few words of device state: the max-delta accumulator, the convergence flag that
//             lets already-queued passes retire as no-ops, the final iteration count.
//
// A "pass" is one kernel launch that performs up to T consecutive red-black half-sweeps on every
// tile in shared memory (temporal blocking).  `run()` cuts any iteration range into passes so that
// a convergence-check sweep is always the last sweep of its pass; `solve()` is the reference's
// harmonic_execute_gpu loop (harmonic_gpu.cu:226-305) with the checks decided on the device and
// read back asynchronously.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace epic_b200 {

enum MathMode { MATH_STRICT = 0, MATH_FAST = 1 };

struct Ctrl;
struct FieldView2D;

constexpr int kMaxSlabs = 16;   // slabs of one grid held by one process (Grid, grid.h)

struct FieldConfig {
    int device = -1;          // -1 = the current device
    int ndevices = 0;         // EPIC_DEVICES: the devices a whole grid is sharded over by the libepic ABI (Grid);
    int devices[kMaxSlabs] = {0};   // 0 = not set (one slab on `device`).  An ordinal may repeat (several slabs on one GPU).
    MathMode math = MATH_STRICT;
    int sweeps_per_pass = 0;  // T; 0 = default
    int tile_rows = 0;        // 2-D: rows of the shared-memory tile (incl. halo); 0 = by grid size
    int threads = 0;          // 2-D: threads per CTA (256 or 512); 0 = by tile size
    cudaStream_t stream = nullptr;  // run on this stream instead of a private one
    bool use_stream = false;
};

// What a neighbouring rank needs to store its edge layers straight into this slab's ghost layers and to
// signal "pass p done": CUDA IPC handles of the two buffers and of the flag words, plus the geometry.
struct PeerInfo {
    cudaIpcMemHandle_t u[2];
    cudaIpcMemHandle_t flags;
    uint64_t own_lo, own_hi;     // buffer layers owned by that slab
    uint64_t layer_floats;
    uint64_t buf_layers;
    uint32_t ghost;
    int32_t device;
};

// How the slabs of one Grid find each other's inboxes for the all-reduce of the convergence check
// (decide_all_kernel, aux_kernels.cuh), and what the host reads back of a slab's control block.
struct DecideWiring {
    unsigned long long *inbox[kMaxSlabs];
    uint32_t nslabs = 0, me = 0;
};
struct SolveSnapshot {
    uint32_t done = 0, final_iteration = 0, final_buffer = 0, skipped = 0, failed = 0;
    float last_delta = 0.0f;
};

// Return codes are the reference's (libepic/include/epic/error_codes.h:31-46).
class Grid;

class Field {
    friend class Grid;
public:
    // n = 2 or 3; gm = global dimensions; this slab owns x0 in [row0, row0+rows).
    // `ghost` x0-layers are kept on each side for neighbouring slabs (0 for a whole grid).
    static int create(Field **out, unsigned n, const uint64_t *gm, uint64_t row0, uint64_t rows,
                      unsigned ghost, const FieldConfig &cfg);
    ~Field();

    // Host <-> device.  Host arrays are dense; `layers` x0-layers starting at global layer `first`
    // (must lie inside the owned + ghost range).  All four are complete on return.
    int upload_u(const float *host, uint64_t first, uint64_t layers);
    int upload_locked(const uint32_t *host, uint64_t first, uint64_t layers);
    int download_u(float *host, uint64_t first, uint64_t layers);
    int download_locked(uint32_t *host, uint64_t first, uint64_t layers);

    // `count` half-sweeps starting at iteration `it0`, enqueued on the stream (no sync).  If
    // `check_last`, the last one accumulates delta; fetch it with read_delta().
    int run(uint32_t it0, uint32_t count, bool check_last);
    // Fetch (and reset) the delta accumulated by the last check sweep; syncs the stream.
    int read_delta(float *delta);

    // The reference's execute loop: iterate from iteration 0 until a check sweep sees
    // delta < epsilon with currentIteration >= m_max.  Outputs the final iteration and delta.
    int solve(float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iteration, float *delta);

    // k sparse edits (x = column, y = global row), types 0 goal / 1 obstacle / 2 free.
    int set_cells_2d(uint32_t k, const uint32_t *v, const uint32_t *types);

    // Dense map ingest (2-D): an occupancy grid of one byte per cell covering global rows
    // [first, first + layers), classified on the device; and "reset free cells".  See reclassify_2d_kernel.
    int ingest_occupancy_2d(const signed char *host, uint64_t first, uint64_t layers, int threshold, int no_change);
    int reset_free_cells_2d();

    // Streamlines on the device-resident field (2-D fields holding the rows the path visits).
    // Same arithmetic, operation for operation, as the reference's CPU functions.
    int potential_2d(float x, float y, float *value);
    int gradient_2d(float x, float y, float cd, float *px, float *py);
    // `count` paths in one go; ret[i], k[i] per path; paths[i] = new float[2*k[i]] (or nullptr).
    int paths_2d(uint32_t count, const float *starts, float step, float cd, uint32_t max_length,
                 int *ret, uint32_t *k, float **paths);

    // Ghost-layer plumbing for sharded runs: device pointer into the CURRENT buffer at global
    // layer `layer` (owned or ghost).
    float *layer_ptr(int64_t layer);

    // Peer-to-peer halos.  dir 0 = the slab above (lower x0), 1 = the slab below.  Once a peer is set,
    // every pass (a) waits until that neighbour has signalled the completion of as many passes as this
    // slab has issued, (b) stores its edge layers into the neighbour's ghost layers from inside the
    // sweep kernel (NVLink peer stores), and (c) signals its own completion to the neighbour with a
    // stream-ordered flag write.  All slabs of a grid must issue the same sequence of passes.
    int peer_export(PeerInfo *out);
    int set_peer_ipc(int dir, const PeerInfo *info);      // neighbour lives in another process
    int set_peer_local(int dir, Field *other);            // neighbour lives in this process
    bool has_peers() const { return peer_[0].on || peer_[1].on; }

    // Static-tile skipping for the passes run() issues from now on (solve() switches it on by itself).  The
    // caller promises that nothing but these passes and the neighbours' halo stores touches the field in
    // between; uploads and edits reset the bookkeeping on their own.  Meant for sharded solves to epsilon;
    // throughput measurements leave it off so that every tile is swept.
    void set_tracking(bool on) { tracking_ = on; }

    int sync();
    cudaStream_t stream() const { return stream_; }
    int sweeps_per_pass() const { return T_; }
    uint64_t pitch() const { return pitch_; }
    uint64_t layer_floats() const { return layer_floats_; }
    unsigned dims() const { return n_; }
    const uint64_t *global_dims() const { return gm_; }
    uint64_t rows() const { return rows_; }
    uint64_t row0() const { return row0_; }
    unsigned ghost() const { return ghost_; }
    MathMode math() const { return cfg_.math; }
    int device() const { return cfg_.device; }
    uint64_t launches() const { return launches_; }
    int tile_rows() const { return TH_; }
    size_t device_bytes() const { return device_bytes_; }
    uint64_t skipped_tiles() const { return skipped_tiles_; }   // by static-tile skipping, over all solves

private:
    Field() {}
    // The streamline entry points on an explicit view of the field: a Grid passes the slabs of all its devices
    // (add_to_view appends this slab), the kernels run on this slab's device and read the others over NVLink.
    void add_to_view(FieldView2D *view) const;
    int potential_view(const FieldView2D &view, float x, float y, float *value);
    int gradient_view(const FieldView2D &view, float x, float y, float cd, float *px, float *py);
    int paths_view(const FieldView2D &view, uint32_t count, const float *starts, float step, float cd,
                   uint32_t max_length, int *ret, uint32_t *k, float **paths);
    // The same on the slabs of a whole Grid (this = the slab whose device runs the kernels).
    int potential_grid(const std::vector<Field *> &slabs, float x, float y, float *value);
    int gradient_grid(const std::vector<Field *> &slabs, float x, float y, float cd, float *px, float *py);
    int paths_grid(const std::vector<Field *> &slabs, uint32_t count, const float *starts, float step, float cd,
                   uint32_t max_length, int *ret, uint32_t *k, float **paths);
    // Pieces of the solve loop a Grid drives on each of its slabs (grid.cu): arm the device-side termination
    // rule; after a check sweep publish this slab's delta to every slab and decide on the maximum; copy the
    // control block to pinned slot `slot` / wait for that copy; adopt the final state.
    int solve_begin(float epsilon, uint32_t m_max);
    int publish_delta(const DecideWiring &w, uint32_t tag);
    int decide_all(const DecideWiring &w, uint32_t tag, uint32_t count, bool rule);
    int snapshot(int slot);
    int wait_snapshot(int slot, SolveSnapshot *out);
    int solve_end(const SolveSnapshot &fin);
    static int default_sweeps_per_pass(unsigned n, const FieldConfig &cfg);
    // Layers allocated per buffer: padded so that a TMA box never exceeds the tensor it reads from.
    uint64_t alloc_layers() const
    {
        if (n_ == 2) {
            return buf_layers_ > 256 ? buf_layers_ : 256;
        }
        const uint64_t need = (64 + gm_[1] - 1) / gm_[1];   // at least 64 rows of pitch floats
        return buf_layers_ > need ? buf_layers_ : need;
    }
    int build_tensor_maps();
    int launch_pass(uint32_t it0, uint32_t count, bool check_last);
    int launch_pass_2d(uint32_t it0, uint32_t count, bool check_last);
    int launch_pass_3d(uint32_t it0, uint32_t count, bool check_last);
    // One solver period (`count` half-sweeps ending in a check sweep, then the termination decision) as a
    // CUDA graph: captured once per (starting buffer, colour phase, count) and replayed, so a period costs one
    // launch call instead of count / T + 1 -- the small maps are bound by the host's launch rate otherwise.
    int run_period(uint32_t it0, uint32_t count);
    struct PeriodGraph {
        uint32_t key = 0, count = 0;
        cudaGraphExec_t exec = nullptr;
        int cur_after = 0;
        uint64_t launches = 0;
        uint32_t passes = 0;
    };
    PeriodGraph graphs_[4];
    bool graphs_off_ = false;
    // Static-tile skipping (2-D whole-grid solves, see Sweep2DParams::chg_prev): two flag arrays that swap
    // with the ping-pong buffers.  `tracking_` is on inside solve(); any other writer of the field (plain
    // run(), uploads, edits) leaves the flags stale, and the next tracked pass resets them to "changed".
    uint8_t *chg_[2] = {nullptr, nullptr};
    size_t chg_bytes_ = 0;
    bool tracking_ = false;
    bool chg_stale_ = true;
    bool skip_static_ = true;    // EPIC_SKIP_STATIC=0 turns the feature off
    bool track_runs_ = false;    // EPIC_SKIP_STATIC=all: also for plain run() passes (libepic update calls)
    uint64_t skipped_tiles_ = 0;

    FieldConfig cfg_;
    unsigned n_ = 0;
    uint64_t gm_[3] = {1, 1, 1};
    uint64_t row0_ = 0, rows_ = 0;
    unsigned ghost_ = 0;
    int64_t grow0_ = 0;          // global layer of buffer layer 0 (= row0 - ghost, may be negative)
    uint64_t pitch_ = 0;         // floats per innermost row
    uint64_t layer_floats_ = 0;  // floats per x0-layer (2-D: pitch; 3-D: m1 * pitch)
    uint64_t buf_layers_ = 0;    // owned + 2*ghost
    uint64_t own_lo_ = 0, own_hi_ = 0;  // buffer layers written by this slab
    float *u_[2] = {nullptr, nullptr};
    int cur_ = 0;
    uint32_t *freemask_ = nullptr;
    uint64_t mask_wpr_ = 0;      // mask words per innermost row
    Ctrl *ctrl_ = nullptr;       // device
    Ctrl *ctrl_host_ = nullptr;  // pinned ring of slots
    static const int kSlots = 4;
    cudaEvent_t events_[kSlots];
    bool events_ok_ = false;
    cudaStream_t stream_ = nullptr;
    bool own_stream_ = false;
    CUtensorMap tmap_[2];
    int T_ = 4, TH_ = 96, NT_ = 256;
    int sms_ = 148;
    int ctas_per_sm_ = 2;        // 3-D: resident CTAs per SM of the sweep kernel (occupancy query)
    uint32_t zchunk_ = 0;        // 3-D: owned layers per CTA (the split along x0, fixed at creation)
    uint64_t launches_ = 0;
    bool attr_done_ = false;
    size_t device_bytes_ = 0;
    void *staging_ = nullptr;    // device scratch for upload_locked / download_locked
    size_t staging_bytes_ = 0;

    struct Peer {
        bool on = false;
        bool ipc = false;
        float *u[2] = {nullptr, nullptr};  // the neighbour's two buffers, mapped here
        uint32_t *flags = nullptr;         // the neighbour's flag words {from_up, from_down}
        uint64_t own_lo = 0, own_hi = 0;
    };
    Peer peer_[2];
    uint32_t *flags_ = nullptr;  // device: {from_up, from_down} = passes completed by the neighbours
    uint32_t pass_count_ = 0;    // passes issued by this slab
    bool kernel_sync_ = true;    // 2-D: order passes between GPUs inside the sweep kernel (EPIC_P2P_SYNC=stream: by
                                 // stream memory operations around it, as the 3-D path does)
    int wait_peers();            // enqueue: flags_[d] >= pass_count_ for every peer
    int signal_peers();          // enqueue: neighbour flags <- pass_count_
    void close_peers();
};

// Parses EPIC_MATH (strict|fast), EPIC_SWEEPS_PER_PASS, EPIC_TILE_ROWS, EPIC_THREADS, EPIC_DEVICE and
// EPIC_DEVICES ("0,1,2,3", "all", or a count "4" = the first four devices).
FieldConfig config_from_env();

}  // namespace epic_b200
