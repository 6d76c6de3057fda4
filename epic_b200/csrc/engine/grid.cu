// grid.cu -- see grid.h.
#include "grid.h"

#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <thread>

namespace epic_b200 {

namespace {

enum {
    kSuccess = 0,
    kInvalidData = 2,
    kInvalidCudaParam = 3,
    kDeviceMalloc = 4,
    kMemcpyToDevice = 5,
    kMemcpyToHost = 6,
    kKernelExecution = 8,
    kDeviceSynchronize = 9
};

struct OnDevice {
    int prev = -1;
    explicit OnDevice(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) {
            cudaSetDevice(dev);
        } else {
            prev = -1;
        }
    }
    ~OnDevice()
    {
        if (prev >= 0) {
            cudaSetDevice(prev);
        }
    }
};

struct Range {      // NVTX range (a no-op unless a profiler injected itself)
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

// Run fn(i) for every slab, on one host thread per slab when there are several (host <-> device copies of
// pageable memory block the calling thread, so N slabs need N threads to use N PCIe links at once).
template <class F>
int for_each_slab(size_t n, bool threads, F fn)
{
    std::vector<int> rc(n, kSuccess);
    if (n == 1 || !threads) {
        for (size_t i = 0; i < n; ++i) {
            rc[i] = fn(i);
        }
    } else {
        std::vector<std::thread> pool;
        for (size_t i = 0; i < n; ++i) {
            pool.emplace_back([&rc, &fn, i]() { rc[i] = fn(i); });
        }
        for (std::thread &t : pool) {
            t.join();
        }
    }
    for (int r : rc) {
        if (r != kSuccess) {
            return r;
        }
    }
    return kSuccess;
}

}  // namespace

int Grid::create(Grid **out, unsigned n, const uint64_t *gm, const FieldConfig &cfg_in)
{
    *out = nullptr;
    if ((n != 2 && n != 3) || gm == nullptr) {
        return kInvalidData;
    }
    Grid *g = new Grid();
    g->n_ = n;
    for (unsigned i = 0; i < n; ++i) {
        g->gm_[i] = gm[i];
    }
    if (const char *e = getenv("EPIC_VERBOSE")) {
        g->verbose_ = atoi(e) != 0;
    }
    FieldConfig cfg = cfg_in;
    // how many slabs: every slab needs room for its ghost exchange (4 passes deep) and a sensible tile
    const int T = Field::default_sweeps_per_pass(n, cfg);
    const uint64_t min_rows = (n == 2) ? (uint64_t)std::max(4 * T, 16) : 8;
    size_t want = cfg.ndevices > 1 ? (size_t)cfg.ndevices : 1;
    want = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(want, gm[0] / min_rows));
    {
        // A small map is bound by launch latency, not by arithmetic: sharding it only adds exchanges.  Use a device
        // per EPIC_MIN_SLAB_CELLS cells (default 2^20; the demo maps stay on one GPU, a 4096^2 grid uses all
        // eight).  0 takes the device list literally (tests).
        uint64_t min_cells = 1ull << 20;
        if (const char *e = getenv("EPIC_MIN_SLAB_CELLS")) {
            min_cells = strtoull(e, nullptr, 10);
        }
        uint64_t cells = 1;
        for (unsigned i = 0; i < n; ++i) {
            cells *= gm[i];
        }
        if (min_cells > 0) {
            want = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(want, cells / min_cells));
        }
    }
    if (want <= 1) {
        if (cfg.ndevices >= 1) {
            cfg.device = cfg.devices[0];
        }
        Field *f = nullptr;
        const int r = Field::create(&f, n, gm, 0, gm[0], 0, cfg);
        if (r != kSuccess) {
            delete g;
            return r;
        }
        g->slabs_.push_back(f);
        *out = g;
        return kSuccess;
    }

    // one stream per distinct device, shared by the slabs that live on it (slabs on one device then run in
    // launch order, which is the order their in-kernel waits need)
    int result = kSuccess;
    std::vector<int> slab_stream(want, 0);
    for (size_t i = 0; i < want && result == kSuccess; ++i) {
        const int dev = cfg.devices[i];
        size_t s = 0;
        while (s < g->stream_device_.size() && g->stream_device_[s] != dev) {
            ++s;
        }
        if (s == g->stream_device_.size()) {
            OnDevice on(dev);
            cudaStream_t st = nullptr;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError();
                result = kDeviceMalloc;
                break;
            }
            g->streams_.push_back(st);
            g->stream_device_.push_back(dev);
        }
        slab_stream[i] = (int)s;
    }
    // all pairs of devices see each other (halo stores go to the neighbours, the convergence all-reduce and the
    // streamline kernels reach every slab)
    for (size_t a = 0; a < g->stream_device_.size() && result == kSuccess; ++a) {
        for (size_t b = 0; b < g->stream_device_.size(); ++b) {
            if (a == b) {
                continue;
            }
            OnDevice on(g->stream_device_[a]);
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->stream_device_[a], g->stream_device_[b]);
            const cudaError_t e = can ? cudaDeviceEnablePeerAccess(g->stream_device_[b], 0) : cudaErrorInvalidDevice;
            cudaGetLastError();
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                fprintf(stderr, "Error[epic_b200]: devices %d and %d of EPIC_DEVICES cannot access each other's memory.\n",
                        g->stream_device_[a], g->stream_device_[b]);
                result = kInvalidCudaParam;
                break;
            }
        }
    }
    for (size_t i = 0; i < want && result == kSuccess; ++i) {
        const uint64_t lo = gm[0] * i / want, hi = gm[0] * (i + 1) / want;
        FieldConfig c = cfg;
        c.device = cfg.devices[i];
        c.stream = g->streams_[(size_t)slab_stream[i]];
        c.use_stream = true;
        Field *f = nullptr;
        result = Field::create(&f, n, gm, lo, hi - lo, (unsigned)T, c);
        if (result == kSuccess) {
            g->slabs_.push_back(f);
            OnDevice on(c.device);
            unsigned long long *box = nullptr;
            const size_t bytes = 2 * (size_t)kMaxSlabs * sizeof(unsigned long long);
            if (cudaMalloc(&box, bytes) != cudaSuccess || cudaMemset(box, 0, bytes) != cudaSuccess) {
                cudaGetLastError();
                result = kDeviceMalloc;
            }
            g->inbox_.push_back(box);
        }
    }
    for (size_t i = 0; i + 1 < g->slabs_.size() && result == kSuccess; ++i) {
        if (g->slabs_[i]->set_peer_local(1, g->slabs_[i + 1]) != kSuccess ||
            g->slabs_[i + 1]->set_peer_local(0, g->slabs_[i]) != kSuccess) {
            fprintf(stderr, "Error[epic_b200]: peer-to-peer halo setup between slabs %zu and %zu failed.\n", i, i + 1);
            result = kInvalidCudaParam;
        }
    }
    if (result == kSuccess && g->slabs_[0]->sweeps_per_pass() != T) {
        result = kInvalidCudaParam;
    }
    if (result != kSuccess) {
        delete g;
        return result;
    }
    g->sync();
    *out = g;
    return kSuccess;
}

Grid::~Grid()
{
    for (Field *f : slabs_) {
        if (f != nullptr) {
            f->sync();
        }
    }
    for (Field *f : slabs_) {
        delete f;
    }
    for (size_t i = 0; i < inbox_.size(); ++i) {
        if (inbox_[i] != nullptr) {
            OnDevice on(slabs_.size() > i && slabs_[i] ? slabs_[i]->device() : 0);
            cudaFree(inbox_[i]);
        }
    }
    for (size_t s = 0; s < streams_.size(); ++s) {
        OnDevice on(stream_device_[s]);
        cudaStreamDestroy(streams_[s]);
    }
}

void Grid::held(size_t i, uint64_t *first, uint64_t *layers) const
{
    const Field *f = slabs_[i];
    const uint64_t g = f->ghost();
    const uint64_t lo = f->row0() > g ? f->row0() - g : 0;
    const uint64_t hi = std::min<uint64_t>(gm_[0], f->row0() + f->rows() + g);
    *first = lo;
    *layers = hi - lo;
}

int Grid::sync()
{
    int r = kSuccess;
    for (Field *f : slabs_) {
        const int q = f->sync();
        r = (r == kSuccess) ? q : r;
    }
    return r;
}

int Grid::upload_u(const float *host)
{
    if (host == nullptr) {
        return kInvalidData;
    }
    Range range("epic_b200::upload_u");
    if (slabs_.size() > 1 && sync() != kSuccess) {    // neighbours may still be storing into ghost layers
        return kDeviceSynchronize;
    }
    const uint64_t layer_cells = (n_ == 2) ? gm_[1] : gm_[1] * gm_[2];
    return for_each_slab(slabs_.size(), true, [&](size_t i) {
        uint64_t first, layers;
        held(i, &first, &layers);
        return slabs_[i]->upload_u(host + first * layer_cells, first, layers);
    });
}

int Grid::upload_locked(const uint32_t *host)
{
    if (host == nullptr) {
        return kInvalidData;
    }
    Range range("epic_b200::upload_locked");
    if (slabs_.size() > 1 && sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    const uint64_t layer_cells = (n_ == 2) ? gm_[1] : gm_[1] * gm_[2];
    return for_each_slab(slabs_.size(), true, [&](size_t i) {
        uint64_t first, layers;
        held(i, &first, &layers);
        return slabs_[i]->upload_locked(host + first * layer_cells, first, layers);
    });
}

int Grid::download_u(float *host)
{
    if (host == nullptr) {
        return kInvalidData;
    }
    Range range("epic_b200::download_u");
    const uint64_t layer_cells = (n_ == 2) ? gm_[1] : gm_[1] * gm_[2];
    return for_each_slab(slabs_.size(), true, [&](size_t i) {
        Field *f = slabs_[i];
        return f->download_u(host + f->row0() * layer_cells, f->row0(), f->rows());
    });
}

int Grid::run(uint32_t it0, uint32_t count, bool check_last)
{
    if (slabs_.size() == 1) {
        return slabs_[0]->run(it0, count, check_last);
    }
    // pass by pass over the slabs: a slab's pass p + 1 waits (in the kernel) for its neighbours' pass p, so the
    // launch order that never makes a device wait for work the host has not issued yet is round-robin
    const uint32_t T = (uint32_t)sweeps_per_pass();
    for (uint32_t done = 0; done < count;) {
        const uint32_t c = std::min(T, count - done);
        for (Field *f : slabs_) {
            const int r = f->run(it0 + done, c, check_last && done + c == count);
            if (r != kSuccess) {
                return r;
            }
        }
        done += c;
    }
    return kSuccess;
}

int Grid::read_delta(float *delta)
{
    if (delta == nullptr) {
        return kInvalidData;
    }
    float best = 0.0f;
    for (Field *f : slabs_) {
        float d = 0.0f;
        const int r = f->read_delta(&d);
        if (r != kSuccess) {
            return r;
        }
        best = std::max(best, d);
    }
    *delta = best;
    return kSuccess;
}

int Grid::solve(float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iteration, float *delta)
{
    Range range("epic_b200::solve");
    const auto t0 = std::chrono::steady_clock::now();
    const int r = (slabs_.size() == 1) ? slabs_[0]->solve(epsilon, stagger, m_max, iteration, delta)
                                       : solve_sharded(epsilon, stagger, m_max, iteration, delta);
    if (r == kSuccess) {
        stats_.last_solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats_.last_solve_iterations = *iteration;
        stats_.last_solve_delta = *delta;
        if (verbose_) {
            const GridStats s = stats();
            double cells = 1.0;
            for (unsigned i = 0; i < n_; ++i) {
                cells *= (double)gm_[i];
            }
            fprintf(stderr, "{\"epic_b200\": \"solve\", \"slabs\": %u, \"iterations\": %u, \"delta\": %.9g, \"seconds\": %.6f, "
                            "\"gcups\": %.3f, \"launches\": %llu, \"skipped_tiles\": [",
                    s.slabs, *iteration, (double)*delta, s.last_solve_seconds,
                    cells / 2.0 * (double)*iteration / s.last_solve_seconds / 1e9, (unsigned long long)s.launches);
            for (uint32_t i = 0; i < s.slabs; ++i) {
                fprintf(stderr, "%s%llu", i ? ", " : "", (unsigned long long)s.skipped_by_slab[i]);
            }
            fprintf(stderr, "]}\n");
        }
    }
    return r;
}

// The reference's execute loop over several slabs.  Period 0 is the check sweep at iteration 0; period k >= 1
// covers iterations (k-1)*stagger+1 .. k*stagger and ends with the check sweep (reference harmonic_gpu.cu:266-282).
// After each period every slab publishes its delta to all slabs and decides on the maximum on the device
// (decide_all_kernel); once the rule is met the passes already queued retire as no-ops.  The host only reads the
// control block two periods behind, so the devices never wait for it.
//
// Slabs on distinct devices are each driven by their own host thread: every thread issues the same periods and
// stops at the same one, because the decision it reads back is the same on every slab.  Slabs that share a
// device share a stream; they are issued from one thread, pass by pass in slab order (see run()).
int Grid::solve_sharded(float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iteration, float *delta)
{
    if (!(epsilon > 0.0f) || stagger == 0 || iteration == nullptr || delta == nullptr) {
        return kInvalidData;
    }
    const size_t W = slabs_.size();
    DecideWiring wiring[kMaxSlabs];
    for (size_t i = 0; i < W; ++i) {
        wiring[i].nslabs = (uint32_t)W;
        wiring[i].me = (uint32_t)i;
        for (size_t j = 0; j < W; ++j) {
            wiring[i].inbox[j] = inbox_[j];
        }
    }
    if (sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    struct Tracking {
        std::vector<Field *> &s;
        std::vector<bool> before;
        explicit Tracking(std::vector<Field *> &slabs) : s(slabs)
        {
            for (Field *f : s) {
                before.push_back(f->tracking_);
                f->tracking_ = true;
            }
        }
        ~Tracking()
        {
            for (size_t i = 0; i < s.size(); ++i) {
                s[i]->tracking_ = before[i];
            }
        }
    } tracking(slabs_);
    for (Field *f : slabs_) {
        const int r = f->solve_begin(epsilon, m_max);
        if (r != kSuccess) {
            return r;
        }
    }
    if (sync() != kSuccess) {       // no slab may publish into an inbox... of a slab that still clears its control block
        return kDeviceSynchronize;
    }
    const bool distinct = streams_.size() == W;
    bool threaded = distinct;
    if (const char *e = getenv("EPIC_GRID_THREADS")) {
        threaded = threaded && atoi(e) != 0;
    }
    const uint32_t tag_base = tag_;
    const uint32_t T = (uint32_t)sweeps_per_pass();
    const int kSlots = Field::kSlots;
    SolveSnapshot fin;
    uint64_t periods_issued = 0;

    // one slab's share of period `period` (threaded mode), or every slab's (group = all slabs)
    auto issue_period = [&](const std::vector<size_t> &group, uint64_t period, uint64_t it) -> int {
        const uint32_t count = (period == 0) ? 1u : stagger;
        for (uint32_t done = 0; done < count;) {
            const uint32_t c = std::min(T, count - done);
            for (size_t i : group) {
                const int r = slabs_[i]->run((uint32_t)(it + done), c, done + c == count);
                if (r != kSuccess) {
                    return r;
                }
            }
            done += c;
        }
        const uint32_t tag = tag_base + (uint32_t)period + 1u;
        for (size_t i : group) {
            const int r = slabs_[i]->publish_delta(wiring[i], tag);
            if (r != kSuccess) {
                return r;
            }
        }
        for (size_t i : group) {
            int r = slabs_[i]->decide_all(wiring[i], tag, count, true);
            if (r == kSuccess) {
                r = slabs_[i]->snapshot((int)(period % kSlots));
            }
            if (r != kSuccess) {
                return r;
            }
        }
        return kSuccess;
    };
    auto drive = [&](const std::vector<size_t> &group, SolveSnapshot *out, uint64_t *issued) -> int {
        uint64_t it = 0;
        for (uint64_t period = 0;; ++period) {
            if (period >= 2) {
                SolveSnapshot s;
                const int r = slabs_[group[0]]->wait_snapshot((int)((period - 2) % kSlots), &s);
                if (r != kSuccess) {
                    return r;
                }
                if (s.done) {
                    *out = s;
                    *issued = period;
                    return kSuccess;
                }
            }
            const uint32_t count = (period == 0) ? 1u : stagger;
            if (it + count > 0xffffffffull) {
                return kInvalidData;   // the reference's 32-bit iteration counter would wrap
            }
            const int r = issue_period(group, period, it);
            if (r != kSuccess) {
                return r;
            }
            it += count;
        }
    };

    int result = kSuccess;
    if (threaded) {
        std::vector<int> rc(W, kSuccess);
        std::vector<SolveSnapshot> fins(W);
        std::vector<uint64_t> issued(W, 0);
        std::vector<std::thread> pool;
        for (size_t i = 0; i < W; ++i) {
            pool.emplace_back([&, i]() {
                const std::vector<size_t> group{i};
                rc[i] = drive(group, &fins[i], &issued[i]);
            });
        }
        for (std::thread &t : pool) {
            t.join();
        }
        for (size_t i = 0; i < W; ++i) {
            if (rc[i] != kSuccess) {
                result = rc[i];
            } else if (issued[i] != issued[0] || fins[i].final_iteration != fins[0].final_iteration ||
                       fins[i].final_buffer != fins[0].final_buffer) {
                result = kKernelExecution;     // the slabs disagree: cannot happen with a consistent all-reduce
            }
        }
        fin = fins[0];
        periods_issued = issued[0];
    } else {
        std::vector<size_t> group;
        for (size_t i = 0; i < W; ++i) {
            group.push_back(i);
        }
        result = drive(group, &fin, &periods_issued);
    }
    tag_ = tag_base + (uint32_t)periods_issued;
    if (sync() != kSuccess && result == kSuccess) {
        result = kDeviceSynchronize;
    }
    if (result == kSuccess && fin.failed) {
        fprintf(stderr, "Error[epic_b200]: a slab of the grid stopped answering during the convergence check.\n");
        result = kKernelExecution;
    }
    for (Field *f : slabs_) {
        const int r = f->solve_end(fin);
        if (result == kSuccess && r != kSuccess) {
            result = r;
        }
    }
    if (result == kSuccess) {
        *iteration = fin.final_iteration;
        *delta = fin.last_delta;
    }
    return result;
}

int Grid::set_cells_2d(uint32_t k, const uint32_t *v, const uint32_t *types)
{
    // every slab applies the edits that fall into the layers it holds (owned or ghost)
    for (Field *f : slabs_) {
        const int r = f->set_cells_2d(k, v, types);
        if (r != kSuccess) {
            return r;
        }
    }
    return kSuccess;
}

int Grid::ingest_occupancy_2d(const signed char *host, int threshold, int no_change)
{
    if (host == nullptr || n_ != 2) {
        return kInvalidData;
    }
    if (slabs_.size() > 1 && sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    return for_each_slab(slabs_.size(), true, [&](size_t i) {
        uint64_t first, layers;
        held(i, &first, &layers);
        return slabs_[i]->ingest_occupancy_2d(host + first * gm_[1], first, layers, threshold, no_change);
    });
}

int Grid::reset_free_cells_2d()
{
    if (slabs_.size() > 1 && sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    for (Field *f : slabs_) {
        const int r = f->reset_free_cells_2d();
        if (r != kSuccess) {
            return r;
        }
    }
    return kSuccess;
}

// Streamlines read every slab from the first slab's device (peer access over NVLink for the others).
int Grid::potential_2d(float x, float y, float *value)
{
    if (slabs_.size() == 1) {
        return slabs_[0]->potential_2d(x, y, value);
    }
    if (n_ != 2 || value == nullptr) {
        return kInvalidData;
    }
    if (sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    return slabs_[0]->potential_grid(slabs_, x, y, value);
}

int Grid::gradient_2d(float x, float y, float cd, float *px, float *py)
{
    if (slabs_.size() == 1) {
        return slabs_[0]->gradient_2d(x, y, cd, px, py);
    }
    if (n_ != 2 || px == nullptr || py == nullptr) {
        return kInvalidData;
    }
    if (sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    return slabs_[0]->gradient_grid(slabs_, x, y, cd, px, py);
}

int Grid::paths_2d(uint32_t count, const float *starts, float step, float cd, uint32_t max_length, int *ret,
                   uint32_t *k, float **paths)
{
    Range range("epic_b200::paths_2d");
    if (slabs_.size() == 1) {
        return slabs_[0]->paths_2d(count, starts, step, cd, max_length, ret, k, paths);
    }
    if (n_ != 2 || count == 0 || starts == nullptr || ret == nullptr || k == nullptr || paths == nullptr) {
        return kInvalidData;
    }
    if (sync() != kSuccess) {
        return kDeviceSynchronize;
    }
    return slabs_[0]->paths_grid(slabs_, count, starts, step, cd, max_length, ret, k, paths);
}

GridStats Grid::stats() const
{
    GridStats s = stats_;
    s.slabs = (uint32_t)slabs_.size();
    s.launches = 0;
    s.skipped_tiles = 0;
    for (size_t i = 0; i < slabs_.size(); ++i) {
        s.launches += slabs_[i]->launches();
        s.skipped_by_slab[i] = slabs_[i]->skipped_tiles();
        s.skipped_tiles += slabs_[i]->skipped_tiles();
    }
    return s;
}

}  // namespace epic_b200
