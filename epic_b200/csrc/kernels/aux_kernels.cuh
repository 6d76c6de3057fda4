// aux_kernels.cuh -- the small kernels around the sweep: locked-array packing, sparse cell edits,
// the device-side termination rule, and streamline extraction on the device-resident field.
//
// Reference behaviour replaced here:
//   pack_locked_kernel   the 4-byte-per-cell d_locked array of harmonic_model_gpu.cu:124-160 becomes
//                        1 bit per cell ("free" = locked word is zero)
//   set_cells_2d_kernel  harmonic_utilities_set_cells_2d_gpu (harmonic_utilities_gpu.cu:38-63)
//   decide_kernel        the host-side `delta < epsilon` test and loop condition of
//                        harmonic_execute_gpu (harmonic_gpu.cu:266-290, :409-413), moved onto the device
//                        so that queued passes retire as no-ops once the rule is met
//   path kernels         harmonic_compute_{potential,gradient,path}_2d_cpu
//                        (harmonic_path_cpu.cpp:41-221), same float operations in the same order; this
//                        translation unit is compiled with -fmad=false so no multiply-add is fused.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace epic_b200 {

struct Ctrl {                      // device-resident control block
    uint32_t delta_bits;           // running max |du| of the current check sweep (float bits, >= 0)
    uint32_t done;                 // 1 once a check sweep met the termination rule
    uint32_t final_iteration;      // currentIteration when `done` was raised
    uint32_t final_buffer;         // which ping-pong buffer holds the field of that moment
    float last_delta;              // delta of the most recent check sweep
    uint32_t last_check_iteration; // currentIteration right after that sweep
    uint32_t checks;               // number of check sweeps decided so far
    uint32_t iteration;            // currentIteration as the device sees it (replayed solver periods advance it)
    float epsilon;                 // termination rule of the solve in progress: delta < epsilon ...
    uint32_t m_max;                // ... and currentIteration >= m_max
    uint32_t skipped;              // tiles that returned early as static since the counter was last taken
    uint32_t taken_skipped;        // take_delta_kernel moves `skipped` here for the host to read
    uint32_t failed;               // decide_all_kernel gave up waiting for another slab of the grid
};

// One warp packs 32 consecutive cells of one row into one mask word (ballot), coalesced reads.
// `locked` holds `rows` dense rows of m1 words; mask rows start at `mask` (already offset).
__global__ void pack_locked_kernel(const uint32_t *__restrict__ locked, uint32_t *__restrict__ mask,
                                   uint64_t rows, uint32_t m1, uint32_t mask_wpr)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t total = rows * mask_wpr;
    if (warp >= total) {
        return;
    }
    const uint64_t row = warp / mask_wpr;
    const uint32_t w = (uint32_t)(warp % mask_wpr);
    const uint32_t x = w * 32u + lane;
    const bool is_free = (x < m1) && (__ldg(locked + row * m1 + x) == 0u);
    const uint32_t bits = __ballot_sync(0xffffffffu, is_free);
    if (lane == 0) {
        mask[row * mask_wpr + w] = bits;
    }
}

// Inverse of the above, for download_locked (tests, checkpointing): 1 = locked.
__global__ void unpack_locked_kernel(const uint32_t *__restrict__ mask, uint32_t *__restrict__ locked,
                                     uint64_t rows, uint32_t m1, uint32_t mask_wpr)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * m1) {
        return;
    }
    const uint64_t row = i / m1;
    const uint32_t x = (uint32_t)(i % m1);
    const uint32_t w = __ldg(mask + row * mask_wpr + (x >> 5));
    locked[i] = ((w >> (x & 31u)) & 1u) ? 0u : 1u;
}

// k edits; v = [x0, y0, x1, y1, ...] with x = column, y = global row.  Rows outside
// [grow_lo, grow_hi) belong to another slab; out-of-range cells and unknown types are skipped.
// Edits are applied one after the other by a single thread per *cell chain*: the reference's CPU
// twin applies them in order (harmonic_utilities_cpu.cpp:47-73), so when the same cell appears twice
// the last edit must win.  Thread i therefore skips its edit when a later edit targets the same cell.
__global__ void set_cells_2d_kernel(float *__restrict__ u, uint32_t *__restrict__ mask, uint64_t pitch,
                                    uint32_t mask_wpr, uint32_t m0, uint32_t m1, int64_t grow0,
                                    uint32_t buf_rows, uint32_t k, const uint32_t *__restrict__ v,
                                    const uint32_t *__restrict__ types, const uint32_t *__restrict__ last_writer)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) {
        return;
    }
    if (last_writer[i] != i) {
        return;
    }
    const uint32_t x = v[2 * i], y = v[2 * i + 1], type = types[i];
    if (x >= m1 || y >= m0 || type > 2u) {
        return;
    }
    const int64_t b = (int64_t)y - grow0;
    if (b < 0 || b >= (int64_t)buf_rows) {
        return;
    }
    u[(uint64_t)b * pitch + x] = (type == 0u) ? 0.0f : -1e6f;
    uint32_t *word = mask + (uint64_t)b * mask_wpr + (x >> 5);
    const uint32_t bit = 1u << (x & 31u);
    if (type == 2u) {
        atomicOr(word, bit);
    } else {
        atomicAnd(word, ~bit);
    }
}

// Dense reclassification of the interior cells, one warp per mask word (32 cells of a row).  Replaces the
// k = N scatter lists the reference's node builds for every /map message and for reset-free-cells
// (src/epic_navigation_node_harmonic.cpp:383-426, :582-611; 12 bytes per cell over PCIe, one thread per
// edit) with one byte per cell (mode 0) or nothing at all (mode 1):
//   mode 0  occupancy grid: cells whose value is `no_change` and goal cells (locked with u == 0, the
//           node's isCellGoal) keep their state; value >= threshold -> obstacle (u = -1e6, locked);
//           anything else -> free (u = -1e6, unlocked) -- harmonic_utilities_set_cells_2d's type table.
//   mode 1  every unlocked interior cell back to u = -1e6.
// `occ` is dense over the rows this slab holds: occ[(row - occ_row0) * m1 + x].
__global__ void reclassify_2d_kernel(float *__restrict__ u, uint32_t *__restrict__ mask, uint64_t pitch,
                                     uint32_t mask_wpr, uint32_t m0, uint32_t m1, int64_t grow0, uint32_t buf_rows,
                                     const signed char *__restrict__ occ, int64_t occ_row0, int threshold,
                                     int no_change, int mode)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t words = (m1 + 31u) / 32u;
    if (warp >= (uint64_t)buf_rows * words) {
        return;
    }
    const uint32_t b = (uint32_t)(warp / words), w = (uint32_t)(warp % words);
    const int64_t y = grow0 + b;
    if (y <= 0 || y >= (int64_t)m0 - 1) {
        return;   // warp-uniform: rows outside the grid or on its border are never edited
    }
    const uint32_t x = w * 32u + lane;
    const bool interior = x > 0u && x < m1 - 1u;
    uint32_t *word = mask + (uint64_t)b * mask_wpr + w;
    const uint32_t old = *word;
    bool is_free = (old >> lane) & 1u;
    if (interior) {
        float *cell = u + (uint64_t)b * pitch + x;
        if (mode == 1) {
            if (is_free) {
                *cell = -1e6f;
            }
        } else {
            const int o = occ[(uint64_t)(y - occ_row0) * m1 + x];
            const bool goal = !is_free && *cell == 0.0f;
            if (o != no_change && !goal) {
                *cell = -1e6f;
                is_free = o < threshold;
            }
        }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, is_free);
    if (mode == 0 && lane == 0 && bits != old) {
        *word = bits;
    }
}

// The termination rule of harmonic_execute_gpu, evaluated right after a check sweep.
// `it_after` = currentIteration after that sweep; `buffer` = ping-pong index that now holds the field.
__global__ void decide_kernel(Ctrl *ctrl, float epsilon, uint32_t it_after, uint32_t m_max, uint32_t buffer)
{
    if (ctrl->done) {
        return;
    }
    const float delta = __uint_as_float(ctrl->delta_bits);
    ctrl->delta_bits = 0u;
    ctrl->last_delta = delta;
    ctrl->last_check_iteration = it_after;
    ctrl->iteration = it_after;
    ctrl->checks += 1u;
    if (delta < epsilon && it_after >= m_max) {
        ctrl->final_iteration = it_after;
        ctrl->final_buffer = buffer;
        __threadfence();
        ctrl->done = 1u;
    }
}

// The same decision for a solver period replayed from a CUDA graph: nothing that changes from period to
// period may be a kernel argument, so the iteration counter and the rule's constants live in `ctrl`
// (set_rule_kernel).  `count` = half-sweeps of the period, `buffer` = the ping-pong buffer it ends in.
__global__ void set_rule_kernel(Ctrl *ctrl, float epsilon, uint32_t m_max)
{
    ctrl->epsilon = epsilon;
    ctrl->m_max = m_max;
}

__global__ void decide_period_kernel(Ctrl *ctrl, uint32_t count, uint32_t buffer)
{
    if (ctrl->done) {
        return;
    }
    const float delta = __uint_as_float(ctrl->delta_bits);
    const uint32_t it_after = ctrl->iteration + count;
    ctrl->delta_bits = 0u;
    ctrl->last_delta = delta;
    ctrl->last_check_iteration = it_after;
    ctrl->iteration = it_after;
    ctrl->checks += 1u;
    if (delta < ctrl->epsilon && it_after >= ctrl->m_max) {
        ctrl->final_iteration = it_after;
        ctrl->final_buffer = buffer;
        __threadfence();
        ctrl->done = 1u;
    }
}

// Fetch-and-reset of the delta accumulator for single update_and_check calls.
__global__ void take_delta_kernel(Ctrl *ctrl, uint32_t it_after)
{
    ctrl->last_delta = __uint_as_float(ctrl->delta_bits);
    ctrl->delta_bits = 0u;
    ctrl->last_check_iteration = it_after;
    ctrl->taken_skipped = ctrl->skipped;
    ctrl->skipped = 0u;
}

// The termination rule for a grid that is sharded over several slabs of ONE process (engine/grid.cu): an
// all-reduce(max) of the slabs' deltas done by the decision kernels themselves over NVLink peer memory, so that
// the host never has to gather deltas and can keep queueing solver periods ahead of the device (as it does for a
// single slab).  Every slab has an inbox of 2 x kDecideSlots 64-bit words; slab `me` stores {period tag, delta
// bits} into word [tag & 1][me] of EVERY slab's inbox (one atomic 8-byte store each, system-scope release), then
// waits until its own inbox holds this period's tag from every slab and applies the rule to the maximum.  All
// slabs see the same deltas, so all reach the same decision for the same period.  Two parities suffice: a slab
// can publish period k + 2 only after every slab published k + 1, i.e. after every slab finished reading k.
constexpr int kDecideSlots = 16;

struct DecideAllParams {
    Ctrl *ctrl;
    unsigned long long *inbox;                      // this slab's inbox
    unsigned long long *peer_inbox[kDecideSlots];   // every slab's inbox (this slab's own included)
    uint32_t nslabs, me;
    uint32_t tag;                                   // 1-based period index over the grid's lifetime
    uint32_t count;                                 // half-sweeps in this period
    uint32_t buffer;                                // ping-pong buffer the period ends in
    uint32_t use_ctrl_rule;                         // 1: epsilon / m_max / iteration from ctrl (solve); 0: only publish
                                                    //    the global delta in ctrl->last_delta (update_and_check)
};

// First half: store {tag, my delta} into every slab's inbox.  A separate kernel from the wait below so that slabs
// sharing one stream (several slabs on one device) can all publish before the first of them waits.
__global__ void publish_delta_kernel(DecideAllParams p)
{
    const uint32_t lane = threadIdx.x;
    if (p.use_ctrl_rule && p.ctrl->done) {
        return;     // every slab decided "done" in the same period: nobody publishes, nobody waits
    }
    const uint32_t mine = p.ctrl->delta_bits;
    const uint32_t slot = (p.tag & 1u) * kDecideSlots;
    if (lane < p.nslabs) {
        const unsigned long long word = ((unsigned long long)p.tag << 32) | mine;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_inbox[lane] + slot + p.me), "l"(word) : "memory");
    }
}

__global__ void decide_all_kernel(DecideAllParams p)
{
    const uint32_t lane = threadIdx.x;
    if (p.use_ctrl_rule && p.ctrl->done) {
        return;
    }
    const uint32_t slot = (p.tag & 1u) * kDecideSlots;
    uint32_t bits = 0u;
    bool late = false;
    if (lane < p.nslabs) {
        unsigned long long w, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p.inbox + slot + lane) : "memory");
            if ((uint32_t)(w >> 32) != p.tag) {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                late = (t1 - t0) > 30000000000ull;     // 30 s: a slab of this grid died; give up instead of hanging
            }
        } while ((uint32_t)(w >> 32) != p.tag && !late);
        bits = (uint32_t)w;
    }
    const bool failed = __any_sync(0xffffffffu, late);
    for (int o = 16; o > 0; o >>= 1) {       // deltas are non-negative floats: their bit patterns order like integers
        bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
    }
    if (lane != 0) {
        return;
    }
    Ctrl *ctrl = p.ctrl;
    const float delta = __uint_as_float(bits);
    ctrl->delta_bits = 0u;
    ctrl->last_delta = delta;
    if (failed) {
        ctrl->failed = 1u;
        ctrl->final_buffer = p.buffer;
        __threadfence();
        ctrl->done = 1u;
        return;
    }
    if (!p.use_ctrl_rule) {
        ctrl->taken_skipped = ctrl->skipped;
        ctrl->skipped = 0u;
        return;
    }
    const uint32_t it_after = ctrl->iteration + p.count;
    ctrl->last_check_iteration = it_after;
    ctrl->iteration = it_after;
    ctrl->checks += 1u;
    if (delta < ctrl->epsilon && it_after >= ctrl->m_max) {
        ctrl->final_iteration = it_after;
        ctrl->final_buffer = p.buffer;
        __threadfence();
        ctrl->done = 1u;
    }
}

// ------------------------------------------------------------------------------------------------
// Streamlines

// The device-resident field as the streamline kernels see it: one entry per slab (one for a whole-grid field).
// Slab i holds the authoritative copy of global rows [row_end[i-1], row_end[i]); slabs of a sharded grid may
// live on other GPUs of the node, in which case their rows are read over NVLink (peer access).
constexpr int kViewSlabs = 16;

struct FieldView2D {
    const float *u[kViewSlabs];        // current buffer of slab i, buffer-row 0
    const uint32_t *mask[kViewSlabs];  // free bits, buffer layout
    int64_t grow0[kViewSlabs];         // global row of buffer row 0 (0 for a whole-grid field)
    uint32_t row_end[kViewSlabs];      // first global row NOT owned by slab i
    uint32_t nslabs;
    uint64_t pitch;
    uint32_t mask_wpr;
    uint32_t m0, m1;                   // global dimensions
};

enum { kPathOk = 0, kPathInvalidLocation = 10, kPathInvalidGradient = 12, kPathInvalidPath = 13 };

// (unsigned int)f as x86-64 gcc evaluates it: truncate to a signed 64-bit integer, keep the low word.
__device__ __forceinline__ uint32_t f2u_x86(float f)
{
    return (uint32_t)(uint64_t)__float2ll_rz(f);
}

// Row yc (inside the grid) of the field: where its potentials and its free bits start.  MULTI = false is the
// whole-grid field (one slab, everything about it a kernel-parameter constant); MULTI = true finds the slab that
// owns the row first.  A streamline is a chain of dependent instructions issued by one warp, so every instruction
// of the address arithmetic is latency on the critical path: the taps of a bilinear lookup share their two rows.
struct RowRef {
    const float *u;
    const uint32_t *mask;
};

template <bool MULTI>
__device__ __forceinline__ RowRef row_ref(const FieldView2D &f, uint32_t yc)
{
    RowRef r;
    if (MULTI) {
        uint32_t i = 0;
        while (i + 1u < f.nslabs && yc >= f.row_end[i]) {
            ++i;
        }
        const uint64_t b = (uint64_t)((int64_t)yc - f.grow0[i]);
        r.u = f.u[i] + b * f.pitch;
        r.mask = f.mask[i] + b * f.mask_wpr;
    } else {
        r.u = f.u[0] + (uint64_t)yc * f.pitch;            // a whole-grid field starts at global row 0
        r.mask = f.mask[0] + (uint64_t)yc * f.mask_wpr;
    }
    return r;
}

// Everything one bilinear potential lookup (harmonic_path_cpu.cpp:41-82) reads, fetched with INDEPENDENT loads
// issued back to back (out-of-range cells are not read): a streamline is a chain of dependent steps, so what
// matters is the number of memory round trips per step, not the number of loads.
struct PotentialTaps {
    uint32_t xl, xr, yt, yb;       // the four cell indices
    bool centre_in, taps_in;
    uint32_t centre_word;          // mask word of the cell the point rounds to
    float centre_u;
    float tl, tr, bl, br;
    float x, y;
};

template <bool MULTI>
__device__ __forceinline__ void potential_fetch(const FieldView2D &f, float x, float y, PotentialTaps &t)
{
    t.x = x;
    t.y = y;
    t.xl = f2u_x86(__fsub_rn(x, 0.5f));
    t.xr = f2u_x86(__fadd_rn(x, 0.5f));
    t.yt = f2u_x86(__fsub_rn(y, 0.5f));
    t.yb = f2u_x86(__fadd_rn(y, 0.5f));
    t.centre_in = t.xr < f.m1 && t.yb < f.m0;                 // the centre cell is (xr, yb): (unsigned)(x + 0.5f)
    t.taps_in = t.centre_in && t.xl < f.m1 && t.yt < f.m0;
    t.centre_word = 0u;
    t.centre_u = 0.0f;
    t.tl = t.tr = t.bl = t.br = 0.0f;
    if (t.centre_in) {
        const RowRef rb = row_ref<MULTI>(f, t.yb);
        t.centre_word = __ldg(rb.mask + (t.xr >> 5));
        t.centre_u = __ldg(rb.u + t.xr);
        if (t.taps_in) {
            const RowRef rt = row_ref<MULTI>(f, t.yt);
            t.tl = __ldg(rt.u + t.xl);
            t.tr = __ldg(rt.u + t.xr);
            t.bl = __ldg(rb.u + t.xl);
            t.br = t.centre_u;
        }
    }
}

// harmonic_path_cpu.cpp:52-58: outside the grid, or an obstacle (locked and negative) -> invalid location;
// otherwise the bilinear blend with separate multiplies and adds (:60-79).
__device__ __forceinline__ int potential_eval(const PotentialTaps &t, float *out)
{
    if (!t.centre_in) {
        return kPathInvalidLocation;
    }
    const bool locked = ((t.centre_word >> (t.xr & 31u)) & 1u) == 0u;
    if (locked && t.centre_u < 0.0f) {
        return kPathInvalidLocation;
    }
    if (!t.taps_in) {
        return kPathInvalidLocation;  // the reference would read outside the arrays here
    }
    const float alpha = __fsub_rn(t.x, (float)t.xl);
    const float beta = __fsub_rn(t.y, (float)t.yt);
    const float na = __fsub_rn(1.0f, alpha), nb = __fsub_rn(1.0f, beta);
    const float one = __fadd_rn(__fmul_rn(na, t.tl), __fmul_rn(alpha, t.tr));
    const float two = __fadd_rn(__fmul_rn(na, t.bl), __fmul_rn(alpha, t.br));
    *out = __fadd_rn(__fmul_rn(nb, one), __fmul_rn(beta, two));
    return kPathOk;
}

template <bool MULTI>
__device__ int potential_2d(const FieldView2D &f, float x, float y, float *out)
{
    PotentialTaps t;
    potential_fetch<MULTI>(f, x, y, t);
    return potential_eval(t, out);
}

// harmonic_path_cpu.cpp:85-118 from four potentials (any failure -> invalid gradient).
__device__ __forceinline__ int gradient_from(int r, float v0, float v1, float v2, float v3, float cd, float *px, float *py)
{
    if (r != kPathOk) {
        return kPathInvalidGradient;
    }
    const float two_cd = __fmul_rn(2.0f, cd);
    float gx = __fdiv_rn(__fsub_rn(v1, v0), two_cd);
    float gy = __fdiv_rn(__fsub_rn(v3, v2), two_cd);
    const double dx = (double)gx, dy = (double)gy;
    const float denom = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    *px = __fdiv_rn(gx, denom);
    *py = __fdiv_rn(gy, denom);
    return kPathOk;
}

template <bool MULTI>
__device__ int gradient_2d(const FieldView2D &f, float x, float y, float cd, float *px, float *py)
{
    PotentialTaps t0, t1, t2, t3;
    potential_fetch<MULTI>(f, __fsub_rn(x, cd), y, t0);
    potential_fetch<MULTI>(f, __fadd_rn(x, cd), y, t1);
    potential_fetch<MULTI>(f, x, __fsub_rn(y, cd), t2);
    potential_fetch<MULTI>(f, x, __fadd_rn(y, cd), t3);
    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
    int r = potential_eval(t0, &v0);
    r += potential_eval(t1, &v1);
    r += potential_eval(t2, &v2);
    r += potential_eval(t3, &v3);
    return gradient_from(r, v0, v1, v2, v3, cd, px, py);
}

struct PathState {        // lets a long path continue across launches
    float x, y;
    float hx[5], hy[5];   // the previous points, most recent first
    uint32_t nhist;       // valid entries in hx/hy
    uint32_t points;      // points emitted so far (including the start)
    int status;           // -1 running, otherwise the final return code
};

// One WARP per path (kPathWarps paths per CTA).  A streamline is sequential -- each point needs the gradient at
// the previous one -- so the time per point is the length of the dependent chain: the warp shortens it by doing
// the independent pieces of a step side by side.  Lane p < 4 fetches and blends the p-th potential of the central
// difference (six independent loads, one memory round trip), lane 4 fetches the lock bit of the current cell,
// lanes 0..4 each measure the distance to one of the five previous points (the reference's stuck test,
// harmonic_path_cpu.cpp:121-151, one double sqrt each instead of five in a row); a vote and four shuffles bring
// the results together and every lane then advances the same state with the reference's float operations in
// the reference's order.  Emits up to `chunk` further points into out[path * chunk * 2 ...] and returns; the host
// relaunches while any path is still running.  `points` counts over all launches.
constexpr int kPathWarps = 4;

template <bool MULTI>
__global__ void __launch_bounds__(32 * kPathWarps)
path_2d_kernel(FieldView2D f, uint32_t count, const float *__restrict__ starts, float step, float cd,
               uint64_t max_floats, uint32_t chunk, PathState *states, float *__restrict__ out,
               uint32_t *__restrict__ emitted, uint32_t first_launch)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * kPathWarps + (threadIdx.x >> 5);
    if (i >= count) {
        return;     // warp-uniform
    }
    // The state every lane carries: the current point, counters, status.  The five previous points (the stuck
    // test's history, most recent first) are spread over the warp instead: lane h < 5 keeps entry h, so that
    // "shift the history" is two shuffles and "distance to entry h" needs no selection.
    float x, y, hx = 0.0f, hy = 0.0f;
    uint32_t nhist, points;
    int status;
    float *o = out + (uint64_t)i * chunk * 2;
    uint32_t n = 0;
    if (first_launch) {
        x = starts[2 * i];
        y = starts[2 * i + 1];
        nhist = 0;
        points = 0;
        status = -1;
        const uint32_t xc = f2u_x86(__fadd_rn(x, 0.5f)), yc = f2u_x86(__fadd_rn(y, 0.5f));
        bool blocked = xc >= f.m1 || yc >= f.m0;
        if (!blocked) {
            const RowRef rr = row_ref<MULTI>(f, yc);
            const bool locked = ((__ldg(rr.mask + (xc >> 5)) >> (xc & 31u)) & 1u) == 0u;
            blocked = locked && __ldg(rr.u + xc) < 0.0f;
        }
        if (blocked) {
            status = kPathInvalidLocation;
        } else {
            if (lane == 0) {
                o[0] = x;
                o[1] = y;
            }
            n = 1;
            points = 1;
        }
    } else {
        const PathState &s = states[i];
        x = s.x;
        y = s.y;
        nhist = s.nhist;
        points = s.points;
        status = s.status;
        if (lane < 5u) {
            hx = s.hx[lane];
            hy = s.hy[lane];
        }
    }
    const float half_step = __fdiv_rn(step, 2.0f);
    const uint32_t which = lane & 3u;      // this lane's potential of the central difference
    while (status == -1 && n < chunk) {
        // ---- this lane's share of the step -------------------------------------------------------------
        // EVERY lane runs the same instructions (no lane-dependent branch, which would serialise the pieces):
        // the potential at (x -/+ cd, y) or (x, y -/+ cd) chosen by lane & 3, the lock bit of the current cell,
        // and the distance to this lane's history entry.  The three chains are independent of each other, so
        // their latencies overlap inside the one instruction stream.
        const float qx = (which == 0u) ? __fsub_rn(x, cd) : ((which == 1u) ? __fadd_rn(x, cd) : x);
        const float qy = (which == 2u) ? __fsub_rn(y, cd) : ((which == 3u) ? __fadd_rn(y, cd) : y);
        PotentialTaps t;
        potential_fetch<MULTI>(f, qx, qy, t);
        // loop condition of harmonic_path_cpu.cpp:185-187, evaluated on the current point
        const uint32_t xc = f2u_x86(__fadd_rn(x, 0.5f)), yc = f2u_x86(__fadd_rn(y, 0.5f));
        const bool outside = xc >= f.m1 || yc >= f.m0;
        const uint32_t cword = outside ? 0u : __ldg(row_ref<MULTI>(f, yc).mask + (xc >> 5));
        const bool stop_here = outside || ((cword >> (xc & 31u)) & 1u) == 0u;
        // the stuck test (harmonic_path_cpu.cpp:121-151): lane h measures the distance to the h-th previous point
        const double ddx = (double)__fsub_rn(x, hx), ddy = (double)__fsub_rn(y, hy);
        const float dist = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy))));
        const bool near = lane < nhist && dist < half_step;
        float v = 0.0f;
        const int pr = potential_eval(t, &v);
        // ---- together ------------------------------------------------------------------------------------
        bool stop = __any_sync(0xffffffffu, stop_here || near);
        if (!stop && (uint64_t)points * 2ull >= max_floats) {
            stop = true;
        }
        if (stop) {
            status = (points <= 2u) ? kPathInvalidPath : kPathOk;
            break;
        }
        const int bad = __any_sync(0xffffffffu, pr != kPathOk) ? 1 : 0;
        const float v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 1);
        const float v2 = __shfl_sync(0xffffffffu, v, 2), v3 = __shfl_sync(0xffffffffu, v, 3);
        float gx, gy;
        if (gradient_from(bad, v0, v1, v2, v3, cd, &gx, &gy) != kPathOk) {
            status = kPathInvalidGradient;
            break;
        }
        // history: entry h <- entry h - 1, entry 0 <- the current point
        const float px = __shfl_up_sync(0xffffffffu, hx, 1), py = __shfl_up_sync(0xffffffffu, hy, 1);
        hx = (lane == 0u) ? x : px;
        hy = (lane == 0u) ? y : py;
        if (nhist < 5u) {
            nhist++;
        }
        x = __fadd_rn(x, __fmul_rn(gx, step));
        y = __fadd_rn(y, __fmul_rn(gy, step));
        if (lane == 0) {
            o[2 * n] = x;
            o[2 * n + 1] = y;
        }
        n++;
        points++;
    }
    if (lane < 5u) {
        states[i].hx[lane] = hx;
        states[i].hy[lane] = hy;
    }
    if (lane == 0) {
        states[i].x = x;
        states[i].y = y;
        states[i].nhist = nhist;
        states[i].points = points;
        states[i].status = status;
        emitted[i] = n;
    }
}

__global__ void potential_gradient_kernel(FieldView2D f, float x, float y, float cd, int want_gradient,
                                          float *out, int *ret)
{
    if (want_gradient) {
        *ret = gradient_2d<true>(f, x, y, cd, out, out + 1);
    } else {
        *ret = potential_2d<true>(f, x, y, out);
    }
}

}  // namespace epic_b200
