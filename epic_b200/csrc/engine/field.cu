// field.cu -- see field.h.  Compiled with -fmad=false (bit-exact float arithmetic; every fused
// multiply-add in the kernels is an explicit intrinsic).
#include "field.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "../kernels/aux_kernels.cuh"
#include "../kernels/sweep2d.cuh"
#include "../kernels/sweep3d.cuh"

namespace epic_b200 {

namespace {

// libepic error codes (reference libepic/include/epic/error_codes.h:31-46)
enum {
    kSuccess = 0,
    kConverged = 1,
    kInvalidData = 2,
    kInvalidCudaParam = 3,
    kDeviceMalloc = 4,
    kMemcpyToDevice = 5,
    kMemcpyToHost = 6,
    kDeviceFree = 7,
    kKernelExecution = 8,
    kDeviceSynchronize = 9
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The driver entry point is looked up at run time so that the library has no link-time dependency
// on libcuda.so (it must load, for symbol checks, on machines without a driver).
EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = (EncodeTiledFn)p;
        }
    }
    return fn;
}

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

StreamValue32Fn stream_memop(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) {
        return (StreamValue32Fn)p;
    }
    return nullptr;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) {
            cudaSetDevice(dev);
        } else {
            prev = -1;
        }
    }
    ~DeviceGuard()
    {
        if (prev >= 0) {
            cudaSetDevice(prev);
        }
    }
};

uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

FieldView2D empty_view()
{
    FieldView2D f;
    memset(&f, 0, sizeof(f));
    return f;
}

template <class K>
bool allow_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

template <class Math, int NT>
bool allow_smem_modes()
{
    return allow_smem(sweep2d_kernel<Math, NT, kPlain>, 227 * 1024) && allow_smem(sweep2d_kernel<Math, NT, kTrack>, 227 * 1024) &&
           allow_smem(sweep2d_kernel<Math, NT, kP2P>, 227 * 1024) &&
           allow_smem(sweep2d_kernel<Math, NT, kTrack | kP2P>, 227 * 1024);
}

bool allow_smem_2d()
{
    return allow_smem_modes<StrictMath, 256>() && allow_smem_modes<StrictMath, 512>() && allow_smem_modes<FastMath, 256>() &&
           allow_smem_modes<FastMath, 512>();
}

}  // namespace

FieldConfig config_from_env()
{
    FieldConfig cfg;
    if (const char *s = getenv("EPIC_MATH")) {
        if (strcmp(s, "fast") == 0) {
            cfg.math = MATH_FAST;
        } else if (strcmp(s, "strict") != 0) {
            fprintf(stderr, "Warning[epic_b200]: EPIC_MATH='%s' is neither 'strict' nor 'fast'; using strict.\n", s);
        }
    }
    if (const char *s = getenv("EPIC_SWEEPS_PER_PASS")) {
        cfg.sweeps_per_pass = atoi(s);
    }
    if (const char *s = getenv("EPIC_TILE_ROWS")) {
        cfg.tile_rows = atoi(s);
    }
    if (const char *s = getenv("EPIC_THREADS")) {
        cfg.threads = atoi(s);
    }
    if (const char *s = getenv("EPIC_DEVICE")) {
        cfg.device = atoi(s);
    }
    if (const char *s = getenv("EPIC_DEVICES")) {
        int have = 0;
        cudaGetDeviceCount(&have);
        cudaGetLastError();
        if (strcmp(s, "all") == 0) {
            for (int i = 0; i < have && i < kMaxSlabs; ++i) {
                cfg.devices[cfg.ndevices++] = i;
            }
        } else if (strchr(s, ',') == nullptr) {
            const int n = atoi(s);                 // a count: the first n devices (one specific device: EPIC_DEVICE)
            for (int i = 0; i < n && i < have && i < kMaxSlabs; ++i) {
                cfg.devices[cfg.ndevices++] = i;
            }
        } else {
            const char *q = s;
            while (*q != 0 && cfg.ndevices < kMaxSlabs) {
                char *end = nullptr;
                const long d = strtol(q, &end, 10);
                if (end == q) {
                    break;
                }
                if (d >= 0 && d < have) {
                    cfg.devices[cfg.ndevices++] = (int)d;
                } else {
                    fprintf(stderr, "Warning[epic_b200]: EPIC_DEVICES names device %ld, which does not exist; ignored.\n", d);
                }
                q = (*end == ',') ? end + 1 : end;
            }
        }
    }
    return cfg;
}

int Field::create(Field **out, unsigned n, const uint64_t *gm, uint64_t row0, uint64_t rows, unsigned ghost,
                  const FieldConfig &cfg)
{
    *out = nullptr;
    if ((n != 2 && n != 3) || gm == nullptr || rows == 0) {
        return kInvalidData;
    }
    for (unsigned i = 0; i < n; ++i) {
        if (gm[i] == 0 || gm[i] > 0x7fffffffull) {
            return kInvalidData;
        }
    }
    if (row0 + rows > gm[0]) {
        return kInvalidData;
    }
    Field *f = new Field();
    f->cfg_ = cfg;
    if (f->cfg_.device < 0) {
        if (cudaGetDevice(&f->cfg_.device) != cudaSuccess) {
            delete f;
            return kDeviceMalloc;
        }
    }
    DeviceGuard guard(f->cfg_.device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, f->cfg_.device) != cudaSuccess) {
        delete f;
        return kDeviceMalloc;
    }
    if (prop.major < 10) {
        fprintf(stderr, "Error[epic_b200]: this library contains sm_100a code only; device %d is sm_%d%d.\n",
                f->cfg_.device, prop.major, prop.minor);
        delete f;
        return kInvalidCudaParam;
    }
    f->sms_ = prop.multiProcessorCount;
    f->n_ = n;
    for (unsigned i = 0; i < n; ++i) {
        f->gm_[i] = gm[i];
    }
    f->row0_ = row0;
    f->rows_ = rows;
    f->ghost_ = ghost;
    f->grow0_ = (int64_t)row0 - (int64_t)ghost;
    f->buf_layers_ = rows + 2ull * ghost;
    f->own_lo_ = ghost;
    f->own_hi_ = ghost + rows;
    const uint64_t inner = gm[n - 1];
    f->pitch_ = std::max<uint64_t>(round_up(inner, 32), (n == 2) ? (uint64_t)kTileW : (uint64_t)k3W);
    f->mask_wpr_ = f->pitch_ / 32;
    f->layer_floats_ = (n == 2) ? f->pitch_ : gm[1] * f->pitch_;

    // Tile geometry (2-D): the largest tile that still gives every SM a CTA.
    f->T_ = (n == 2) ? 4 : k3HR;
    if (cfg.sweeps_per_pass > 0 && n == 2) {
        f->T_ = std::min(cfg.sweeps_per_pass, 8);
    }
    if (n == 3) {
        // 3-D: a column of BH x 128 tiles walked along x0 (sweep3d.cuh); BH rows include 2 halo rows per side
        f->TH_ = 34;
        if (cfg.tile_rows >= 8 && cfg.tile_rows <= 40 && cfg.tile_rows % 2 == 0) {
            f->TH_ = cfg.tile_rows;
        }
        f->NT_ = k3Threads;
        const size_t smem = sweep3d_smem_bytes((uint32_t)f->TH_);
        if (!allow_smem(sweep3d_kernel<StrictMath>, smem) || !allow_smem(sweep3d_kernel<FastMath>, smem) ||
            !allow_smem(sweep3d_kernel<StrictMath, true>, smem) || !allow_smem(sweep3d_kernel<FastMath, true>, smem)) {
            cudaGetLastError();
            delete f;
            return kInvalidCudaParam;
        }
        f->attr_done_ = true;
        int k = 0;
        const cudaError_t e = (cfg.math == MATH_STRICT)
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep3d_kernel<StrictMath>, k3Threads, smem)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep3d_kernel<FastMath>, k3Threads, smem);
        if (e != cudaSuccess || k < 1) {
            cudaGetLastError();
            k = 1;
        }
        f->ctas_per_sm_ = k;
        if (gm[1] * f->pitch_ >= 0x80000000ull) {
            delete f;
            return kInvalidData;   // the 3-D kernel addresses a layer with 32-bit offsets
        }
        // Layers per CTA: a CTA costs (its layers + the 2 * 2 halo layers); the grid runs in waves of
        // ctas_per_sm_ CTAs per SM.  Take the split along x0 with the cheapest estimated pass.
        {
            const uint64_t ntx3 = (gm[2] + k3OutW - 1) / k3OutW;
            const uint64_t nty3 = (gm[1] + (f->TH_ - 2 * k3HR) - 1) / (f->TH_ - 2 * k3HR);
            const uint64_t tiles_xy = ntx3 * nty3;
            const uint64_t slots = (uint64_t)f->ctas_per_sm_ * (uint64_t)f->sms_;
            uint64_t best_cost = 0;
            uint32_t best_chunk = (uint32_t)rows;
            const uint64_t max_ntz = std::max<uint64_t>(1, std::min<uint64_t>(256, rows / 4));
            for (uint64_t ntz = 1; ntz <= max_ntz; ++ntz) {
                const uint64_t chunk = (rows + ntz - 1) / ntz;
                const uint64_t ctas = tiles_xy * ((rows + chunk - 1) / chunk);
                const uint64_t cost = ((ctas + slots - 1) / slots) * (chunk + 2 * k3HR + 2);
                if (best_cost == 0 || cost < best_cost) {
                    best_cost = cost;
                    best_chunk = (uint32_t)chunk;
                }
            }
            f->zchunk_ = best_chunk;
            f->chg_bytes_ = (size_t)round_up(tiles_xy * ((rows + best_chunk - 1) / best_chunk), 256);
        }
    }
    if (n == 2) {
        // Tile geometry: estimate the time of one pass for every candidate (threads, tile rows) and keep
        // the cheapest.  A tile costs its output rows over `eff`; an SM works through its tiles k at a time (k =
        // resident CTAs); `eff` is the measured relative speed of the candidate on a grid large enough
        // to hide wave effects (tools/sweep_timing.py, profiles/r01i_tile_candidates.txt).  Small grids
        // end up with small tiles (every SM gets work, short passes), slabs of a sharded grid with a
        // tile height that fills the last wave.
        if (!allow_smem_2d()) {
            cudaGetLastError();
            delete f;
            return kInvalidCudaParam;
        }
        f->attr_done_ = true;
        struct Cand { int nt, th; float eff_strict, eff_fast; };
        // strict column: gpurun_out/r02b_tiles.log -> profiles/r02b_tile_candidates.txt (round 2: with the expf table in
        // registers the 256-thread kernel, 72 registers, overtook the 512-thread one at its 64-register cap);
        // entries that were not re-measured are the round-1 figures scaled by their measured neighbours
        static const Cand cands[] = {{512, 96, 0.954f, 0.952f}, {512, 88, 0.940f, 0.936f}, {512, 80, 0.925f, 0.919f},
                                     {512, 72, 0.915f, 0.902f}, {512, 64, 0.906f, 0.978f}, {256, 96, 0.940f, 0.930f},
                                     {256, 64, 1.000f, 0.971f}, {256, 56, 0.987f, 0.948f}, {256, 48, 0.960f, 1.000f},
                                     {256, 40, 0.921f, 0.992f}, {256, 32, 0.883f, 0.981f}, {256, 24, 0.778f, 0.865f},
                                     {256, 16, 0.54f, 0.60f}};
        const int HC = 4 * ((f->T_ + 3) / 4);
        const int out_w = kTileW - 2 * HC;
        const uint64_t ntx = (gm[1] + out_w - 1) / out_w;
        double best = 0.0;
        for (const Cand &c : cands) {
            if (c.th <= 2 * f->T_ + 4) {
                continue;
            }
            const size_t smem = sweep2d_smem_bytes((uint32_t)c.th, (uint32_t)c.nt);
            int k = 0;
            cudaError_t e;
            if (cfg.math == MATH_STRICT) {
                e = (c.nt == 512) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<StrictMath, 512>, 512, smem)
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<StrictMath, 256>, 256, smem);
            } else {
                e = (c.nt == 512) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<FastMath, 512>, 512, smem)
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<FastMath, 256>, 256, smem);
            }
            if (e != cudaSuccess || k < 1) {
                cudaGetLastError();
                k = 1;
            }
            const uint64_t nty = (rows + (c.th - 2 * f->T_) - 1) / (c.th - 2 * f->T_);
            const double tiles = (double)(ntx * nty);
            const double eff = (cfg.math == MATH_STRICT) ? c.eff_strict : c.eff_fast;
            const double per_tile = (double)(c.th - 2 * f->T_) / eff;   // output rows per tile over relative speed
            const double by_sm = ceil(tiles / f->sms_) * per_tile;                          // perfect packing
            const double by_wave = ceil(tiles / ((double)f->sms_ * k)) * k * per_tile;      // whole waves
            const double t = 0.5 * (by_sm + by_wave);
            if (best == 0.0 || t < best) {
                best = t;
                f->TH_ = c.th;
                f->NT_ = c.nt;
            }
        }
        // An override must leave an output region at least as tall as the halo (TH - 2T >= T): static-tile skipping
        // looks at the 3 x 3 neighbourhood of a tile and the peer-to-peer edge classification at the first / last
        // tile rows, both of which assume that a tile's input does not reach beyond its direct neighbours.
        if (cfg.tile_rows >= 3 * f->T_ && cfg.tile_rows > 2 * f->T_ + 1 && cfg.tile_rows <= 200) {
            f->TH_ = cfg.tile_rows;
        } else if (cfg.tile_rows != 0) {
            fprintf(stderr, "Warning[epic_b200]: EPIC_TILE_ROWS=%d does not fit %d sweeps per pass (need %d..200); ignored.\n",
                    cfg.tile_rows, f->T_, 3 * f->T_);
        }
        if (cfg.threads == 256 || cfg.threads == 512) {
            f->NT_ = cfg.threads;
        }
        {   // resident CTAs per SM of the final choice: the L2 prefetch of the sweep kernel looks that far ahead
            const size_t smem = sweep2d_smem_bytes((uint32_t)f->TH_, (uint32_t)f->NT_);
            int k = 0;
            cudaError_t e;
            if (cfg.math == MATH_STRICT) {
                e = (f->NT_ == 512) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<StrictMath, 512>, 512, smem)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<StrictMath, 256>, 256, smem);
            } else {
                e = (f->NT_ == 512) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<FastMath, 512>, 512, smem)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, sweep2d_kernel<FastMath, 256>, 256, smem);
            }
            if (e != cudaSuccess || k < 1) {
                cudaGetLastError();
                k = 1;
            }
            f->ctas_per_sm_ = k;
        }
    }

    // Device memory.  The 2-D buffers are padded to at least one tile so a TMA box never exceeds
    // the tensor it reads from.
    const uint64_t alloc_layers = f->alloc_layers();
    const size_t ubytes = (size_t)alloc_layers * f->layer_floats_ * sizeof(float);
    const uint64_t mask_rows = (n == 2) ? alloc_layers : alloc_layers * gm[1];
    const size_t mbytes = (size_t)mask_rows * f->mask_wpr_ * sizeof(uint32_t);
    bool ok = true;
    const int nbuf = 2;
    for (int i = 0; i < nbuf && ok; ++i) {
        ok = cudaMalloc(&f->u_[i], ubytes) == cudaSuccess && cudaMemset(f->u_[i], 0, ubytes) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&f->freemask_, mbytes) == cudaSuccess && cudaMemset(f->freemask_, 0, mbytes) == cudaSuccess;
    ok = ok && cudaMalloc(&f->ctrl_, sizeof(Ctrl)) == cudaSuccess && cudaMemset(f->ctrl_, 0, sizeof(Ctrl)) == cudaSuccess;
    ok = ok && cudaMalloc(&f->flags_, 256) == cudaSuccess && cudaMemset(f->flags_, 0, 256) == cudaSuccess;
    {
        if (n == 2) {
            const int HC = 4 * ((f->T_ + 3) / 4);
            const uint64_t ntx = (gm[1] + (kTileW - 2 * HC) - 1) / (kTileW - 2 * HC);
            const uint64_t nty = (rows + (f->TH_ - 2 * f->T_) - 1) / (f->TH_ - 2 * f->T_);
            f->chg_bytes_ = (size_t)round_up(ntx * nty, 256);
        }
        for (int i = 0; i < 2 && ok; ++i) {
            ok = cudaMalloc(&f->chg_[i], f->chg_bytes_) == cudaSuccess && cudaMemset(f->chg_[i], 1, f->chg_bytes_) == cudaSuccess;
        }
        if (const char *e = getenv("EPIC_SKIP_STATIC")) {
            // 0: never; 1 (default): inside solves to epsilon; all: also for the passes of run() / the libepic
            // update calls (an anytime caller that keeps relaxing a mostly converged field)
            f->skip_static_ = strcmp(e, "all") == 0 || atoi(e) != 0;
            f->track_runs_ = strcmp(e, "all") == 0;
            f->tracking_ = f->track_runs_;
        }
        if (const char *e = getenv("EPIC_P2P_SYNC")) {
            f->kernel_sync_ = strcmp(e, "stream") != 0;
        }
    }
    ok = ok && cudaMallocHost(&f->ctrl_host_, sizeof(Ctrl) * kSlots) == cudaSuccess;
    if (ok) {
        memset(f->ctrl_host_, 0, sizeof(Ctrl) * kSlots);
        f->device_bytes_ = ubytes * nbuf + mbytes + sizeof(Ctrl);
        if (cfg.use_stream) {
            f->stream_ = cfg.stream;
        } else {
            ok = cudaStreamCreateWithFlags(&f->stream_, cudaStreamNonBlocking) == cudaSuccess;
            f->own_stream_ = ok;
        }
    }
    for (int i = 0; i < kSlots && ok; ++i) {
        ok = cudaEventCreateWithFlags(&f->events_[i], cudaEventDisableTiming) == cudaSuccess;
        f->events_ok_ = ok;
    }
    if (!ok) {
        cudaGetLastError();
        delete f;
        return kDeviceMalloc;
    }
    {
        const int r = f->build_tensor_maps();
        if (r != kSuccess) {
            delete f;
            return r;
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        delete f;
        return kDeviceSynchronize;
    }
    *out = f;
    return kSuccess;
}

Field::~Field()
{
    if (cfg_.device >= 0) {
        DeviceGuard guard(cfg_.device);
        if (stream_ != nullptr || cfg_.use_stream) {
            cudaStreamSynchronize(stream_);
        }
        for (int i = 0; i < 2; ++i) {
            if (u_[i]) cudaFree(u_[i]);
        }
        close_peers();
        for (PeriodGraph &g : graphs_) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
        }
        if (freemask_) cudaFree(freemask_);
        for (int i = 0; i < 2; ++i) {
            if (chg_[i]) cudaFree(chg_[i]);
        }
        if (flags_) cudaFree(flags_);
        if (ctrl_) cudaFree(ctrl_);
        if (ctrl_host_) cudaFreeHost(ctrl_host_);
        if (staging_) cudaFree(staging_);
        if (events_ok_) {
            for (int i = 0; i < kSlots; ++i) cudaEventDestroy(events_[i]);
        }
        if (own_stream_) cudaStreamDestroy(stream_);
    }
}

int Field::build_tensor_maps()
{
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) {
        fprintf(stderr, "Error[epic_b200]: cuTensorMapEncodeTiled is not available from the driver.\n");
        return kInvalidCudaParam;
    }
    // 3-D fields are presented as a 2-D tensor of pitch x (layers * m1) rows: a box is one layer-tile
    const uint64_t tensor_rows = (n_ == 2) ? alloc_layers() : alloc_layers() * gm_[1];
    for (int i = 0; i < 2; ++i) {
        cuuint64_t gdim[2] = {(cuuint64_t)pitch_, (cuuint64_t)tensor_rows};
        cuuint64_t gstride[1] = {(cuuint64_t)pitch_ * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)((n_ == 2) ? kTileW : k3W), (cuuint32_t)TH_};
        cuuint32_t estride[2] = {1, 1};
        const CUresult r = enc(&tmap_[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, u_[i], gdim, gstride, box, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fprintf(stderr, "Error[epic_b200]: cuTensorMapEncodeTiled failed (%d).\n", (int)r);
            return kInvalidCudaParam;
        }
    }
    return kSuccess;
}

// ------------------------------------------------------------------------------------------------
// Host <-> device

static bool range_ok(int64_t grow0, uint64_t buf_layers, uint64_t first, uint64_t layers)
{
    const int64_t b = (int64_t)first - grow0;
    return layers > 0 && b >= 0 && (uint64_t)b + layers <= buf_layers;
}

int Field::upload_u(const float *host, uint64_t first, uint64_t layers)
{
    if (host == nullptr || !range_ok(grow0_, buf_layers_, first, layers)) {
        return kInvalidData;
    }
    chg_stale_ = true;
    DeviceGuard guard(cfg_.device);
    if (wait_peers() != kSuccess) {   // neighbours may still be storing into the ghost layers
        return kDeviceSynchronize;
    }
    const uint64_t inner = gm_[n_ - 1];
    const uint64_t rows_per_layer = (n_ == 2) ? 1 : gm_[1];
    float *dst = u_[cur_] + (uint64_t)((int64_t)first - grow0_) * layer_floats_;
    if (cudaMemcpy2DAsync(dst, pitch_ * sizeof(float), host, inner * sizeof(float), inner * sizeof(float),
                          layers * rows_per_layer, cudaMemcpyHostToDevice, stream_) != cudaSuccess ||
        cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToDevice;
    }
    return kSuccess;
}

int Field::download_u(float *host, uint64_t first, uint64_t layers)
{
    if (host == nullptr || !range_ok(grow0_, buf_layers_, first, layers)) {
        return kInvalidData;
    }
    DeviceGuard guard(cfg_.device);
    const uint64_t inner = gm_[n_ - 1];
    const uint64_t rows_per_layer = (n_ == 2) ? 1 : gm_[1];
    const float *src = u_[cur_] + (uint64_t)((int64_t)first - grow0_) * layer_floats_;
    if (cudaMemcpy2DAsync(host, inner * sizeof(float), src, pitch_ * sizeof(float), inner * sizeof(float),
                          layers * rows_per_layer, cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToHost;
    }
    return kSuccess;
}

static int ensure_staging(void **staging, size_t *have, size_t want)
{
    if (*have >= want) {
        return kSuccess;
    }
    if (*staging) {
        cudaFree(*staging);
        *staging = nullptr;
        *have = 0;
    }
    if (cudaMalloc(staging, want) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceMalloc;
    }
    *have = want;
    return kSuccess;
}

int Field::upload_locked(const uint32_t *host, uint64_t first, uint64_t layers)
{
    if (host == nullptr || !range_ok(grow0_, buf_layers_, first, layers)) {
        return kInvalidData;
    }
    chg_stale_ = true;
    DeviceGuard guard(cfg_.device);
    const uint64_t inner = gm_[n_ - 1];
    const uint64_t rows_per_layer = (n_ == 2) ? 1 : gm_[1];
    const uint64_t total_rows = layers * rows_per_layer;
    const uint64_t row_bytes = inner * sizeof(uint32_t);
    const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(total_rows, (64ull << 20) / row_bytes));
    int r = ensure_staging(&staging_, &staging_bytes_, chunk_rows * row_bytes);
    if (r != kSuccess) {
        return r;
    }
    const uint64_t mask_row0 = (uint64_t)((int64_t)first - grow0_) * rows_per_layer;
    for (uint64_t done = 0; done < total_rows; done += chunk_rows) {
        const uint64_t nrows = std::min(chunk_rows, total_rows - done);
        if (cudaMemcpyAsync(staging_, host + done * inner, nrows * row_bytes, cudaMemcpyHostToDevice, stream_) !=
            cudaSuccess) {
            cudaGetLastError();
            return kMemcpyToDevice;
        }
        const uint64_t warps = nrows * mask_wpr_;
        const uint64_t blocks = (warps * 32 + 255) / 256;
        pack_locked_kernel<<<(unsigned)blocks, 256, 0, stream_>>>((const uint32_t *)staging_,
                                                                   freemask_ + (mask_row0 + done) * mask_wpr_, nrows,
                                                                   (uint32_t)inner, (uint32_t)mask_wpr_);
        launches_++;
        if (cudaGetLastError() != cudaSuccess) {
            return kKernelExecution;
        }
        // the staging buffer is reused by the next chunk: pageable copies are synchronous w.r.t. the
        // host buffer only, so wait for the pack kernel
        if (cudaStreamSynchronize(stream_) != cudaSuccess) {
            return kDeviceSynchronize;
        }
    }
    return kSuccess;
}

int Field::download_locked(uint32_t *host, uint64_t first, uint64_t layers)
{
    if (host == nullptr || !range_ok(grow0_, buf_layers_, first, layers)) {
        return kInvalidData;
    }
    DeviceGuard guard(cfg_.device);
    const uint64_t inner = gm_[n_ - 1];
    const uint64_t rows_per_layer = (n_ == 2) ? 1 : gm_[1];
    const uint64_t total_rows = layers * rows_per_layer;
    const uint64_t row_bytes = inner * sizeof(uint32_t);
    const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(total_rows, (64ull << 20) / row_bytes));
    int r = ensure_staging(&staging_, &staging_bytes_, chunk_rows * row_bytes);
    if (r != kSuccess) {
        return r;
    }
    const uint64_t mask_row0 = (uint64_t)((int64_t)first - grow0_) * rows_per_layer;
    for (uint64_t done = 0; done < total_rows; done += chunk_rows) {
        const uint64_t nrows = std::min(chunk_rows, total_rows - done);
        const uint64_t cells = nrows * inner;
        unpack_locked_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream_>>>(
            freemask_ + (mask_row0 + done) * mask_wpr_, (uint32_t *)staging_, nrows, (uint32_t)inner,
            (uint32_t)mask_wpr_);
        launches_++;
        if (cudaGetLastError() != cudaSuccess) {
            return kKernelExecution;
        }
        if (cudaMemcpyAsync(host + done * inner, staging_, nrows * row_bytes, cudaMemcpyDeviceToHost, stream_) !=
                cudaSuccess ||
            cudaStreamSynchronize(stream_) != cudaSuccess) {
            cudaGetLastError();
            return kMemcpyToHost;
        }
    }
    return kSuccess;
}

float *Field::layer_ptr(int64_t layer)
{
    const int64_t b = layer - grow0_;
    if (b < 0 || (uint64_t)b >= buf_layers_) {
        return nullptr;
    }
    return u_[cur_] + (uint64_t)b * layer_floats_;
}

int Field::sync()
{
    DeviceGuard guard(cfg_.device);
    return cudaStreamSynchronize(stream_) == cudaSuccess ? kSuccess : kDeviceSynchronize;
}

// ------------------------------------------------------------------------------------------------
// Peer-to-peer halos

int Field::peer_export(PeerInfo *out)
{
    if (out == nullptr) {
        return kInvalidData;
    }
    DeviceGuard guard(cfg_.device);
    memset(out, 0, sizeof(*out));
    const int nbuf = 2;
    for (int i = 0; i < nbuf; ++i) {
        if (cudaIpcGetMemHandle(&out->u[i], u_[i]) != cudaSuccess) {
            cudaGetLastError();
            return kInvalidCudaParam;
        }
    }
    if (cudaIpcGetMemHandle(&out->flags, flags_) != cudaSuccess) {
        cudaGetLastError();
        return kInvalidCudaParam;
    }
    out->own_lo = own_lo_;
    out->own_hi = own_hi_;
    out->layer_floats = layer_floats_;
    out->buf_layers = buf_layers_;
    out->ghost = ghost_;
    out->device = cfg_.device;
    return kSuccess;
}

int Field::set_peer_ipc(int dir, const PeerInfo *info)
{
    if (dir < 0 || dir > 1 || info == nullptr || info->layer_floats != layer_floats_ || ghost_ == 0 ||
        info->ghost != ghost_) {
        return kInvalidData;
    }
    DeviceGuard guard(cfg_.device);
    if (stream_memop("cuStreamWaitValue32") == nullptr || stream_memop("cuStreamWriteValue32") == nullptr) {
        return kInvalidCudaParam;
    }
    Peer &p = peer_[dir];
    const int nbuf = 2;
    bool ok = true;
    for (int i = 0; i < nbuf && ok; ++i) {
        ok = cudaIpcOpenMemHandle((void **)&p.u[i], info->u[i], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    }
    ok = ok && cudaIpcOpenMemHandle((void **)&p.flags, info->flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        return kInvalidCudaParam;
    }
    p.own_lo = info->own_lo;
    p.own_hi = info->own_hi;
    p.ipc = true;
    p.on = true;
    return kSuccess;
}

int Field::set_peer_local(int dir, Field *other)
{
    if (dir < 0 || dir > 1 || other == nullptr || other->layer_floats_ != layer_floats_ || ghost_ == 0 ||
        other->ghost_ != ghost_) {
        return kInvalidData;
    }
    if (stream_memop("cuStreamWaitValue32") == nullptr || stream_memop("cuStreamWriteValue32") == nullptr) {
        return kInvalidCudaParam;
    }
    if (other->cfg_.device != cfg_.device) {
        DeviceGuard guard(cfg_.device);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, cfg_.device, other->cfg_.device);
        if (!can) {
            return kInvalidCudaParam;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(other->cfg_.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
            return kInvalidCudaParam;
        }
        cudaGetLastError();
    }
    Peer &p = peer_[dir];
    p.u[0] = other->u_[0];
    p.u[1] = other->u_[1];
    p.flags = other->flags_;
    p.own_lo = other->own_lo_;
    p.own_hi = other->own_hi_;
    p.ipc = false;
    p.on = true;
    return kSuccess;
}

void Field::close_peers()
{
    for (Peer &p : peer_) {
        if (p.on && p.ipc) {
            for (int i = 0; i < 2; ++i) {
                if (p.u[i]) cudaIpcCloseMemHandle(p.u[i]);
            }
            if (p.flags) cudaIpcCloseMemHandle(p.flags);
        }
        p = Peer();
    }
}

int Field::wait_peers()
{
    if (!has_peers()) {
        return kSuccess;
    }
    static StreamValue32Fn wait_fn = stream_memop("cuStreamWaitValue32");
    for (int d = 0; d < 2; ++d) {
        if (peer_[d].on &&
            wait_fn((CUstream)stream_, (CUdeviceptr)(flags_ + d), pass_count_, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
            return kDeviceSynchronize;
        }
    }
    return kSuccess;
}

int Field::signal_peers()
{
    pass_count_++;
    if (!has_peers()) {
        return kSuccess;
    }
    static StreamValue32Fn write_fn = stream_memop("cuStreamWriteValue32");
    for (int d = 0; d < 2; ++d) {
        // the upper neighbour (d = 0) sees me as its lower one: its from_down word is flags[1]
        if (peer_[d].on &&
            write_fn((CUstream)stream_, (CUdeviceptr)(peer_[d].flags + (1 - d)), pass_count_, CU_STREAM_WRITE_VALUE_DEFAULT) !=
                CUDA_SUCCESS) {
            return kKernelExecution;
        }
    }
    return kSuccess;
}

// ------------------------------------------------------------------------------------------------
// Sweeps

int Field::launch_pass_2d(uint32_t it0, uint32_t count, bool check_last)
{
    Sweep2DParams p;
    memset(&p, 0, sizeof(p));
    p.dst = u_[cur_ ^ 1];
    p.freemask = freemask_;
    p.ctrl_done = &ctrl_->done;
    p.delta_bits = &ctrl_->delta_bits;
    p.pitch = pitch_;
    p.mask_wpr = (uint32_t)mask_wpr_;
    p.m0 = (uint32_t)gm_[0];
    p.m1 = (uint32_t)gm_[1];
    p.grow0 = (int32_t)grow0_;
    p.buf_rows = (uint32_t)buf_layers_;
    p.own_lo = (uint32_t)own_lo_;
    p.own_hi = (uint32_t)own_hi_;
    p.TH = (uint32_t)TH_;
    p.T = (uint32_t)T_;
    p.HC = 4u * ((p.T + 3u) / 4u);
    p.out_h = p.TH - 2u * p.T;
    p.out_w = kTileW - 2u * p.HC;
    p.ntx = (uint32_t)((gm_[1] + p.out_w - 1) / p.out_w);
    const uint32_t nty = (uint32_t)((rows_ + p.out_h - 1) / p.out_h);
    p.nty = nty;
    p.count = count;
    p.parity0 = (uint32_t)(((int64_t)it0 + grow0_) & 1);
    p.check = check_last ? 1u : 0u;
    const size_t smem = sweep2d_smem_bytes(p.TH, (uint32_t)NT_);
    p.prefetch_stride = (uint32_t)std::max(1, ctas_per_sm_) * (uint32_t)sms_;
    const uint32_t grid = p.ntx * nty;
    p.halo_rows = (uint32_t)std::min<uint64_t>(ghost_, rows_);
    if (peer_[0].on) {
        p.peer_up = peer_[0].u[cur_ ^ 1] + peer_[0].own_hi * pitch_;
    }
    if (peer_[1].on) {
        p.peer_down = peer_[1].u[cur_ ^ 1] + (peer_[1].own_lo - p.halo_rows) * pitch_;
    }
    const bool kernel_sync = has_peers() && kernel_sync_;
    if (kernel_sync) {
        // passes are ordered between the GPUs inside the kernel (see Sweep2DParams::p2p_sync)
        p.p2p_sync = 1;
        p.pass_index = pass_count_ + 1;
        p.wait_up = flags_ + 0;
        p.wait_dn = flags_ + 1;
        p.signal_up = peer_[0].on ? peer_[0].flags + 1 : nullptr;   // the upper neighbour's "from below" word
        p.signal_dn = peer_[1].on ? peer_[1].flags + 0 : nullptr;
        p.edge_count = flags_ + 2;
        p.n_edge_up = p.ntx;
        uint32_t rows_dn = 0;
        for (uint32_t ty = 0; ty < nty; ++ty) {
            if ((int64_t)own_lo_ + (int64_t)ty * p.out_h - p.T + p.TH > (int64_t)own_hi_) {
                rows_dn++;
            }
        }
        p.n_edge_dn = rows_dn * p.ntx;
    } else if (wait_peers() != kSuccess) {
        return kKernelExecution;
    }

    if (!attr_done_) {  // per device, so per field
        if (!allow_smem_2d()) {
            cudaGetLastError();
            return kInvalidCudaParam;
        }
        attr_done_ = true;
    }
    // Static-tile skipping: inside solve(), or when the driver of a sharded solve has switched it on
    // (set_tracking); tiles that read ghost layers always run, see the kernel.
    const bool track = tracking_ && skip_static_ && chg_[0] != nullptr && (size_t)p.ntx * nty <= chg_bytes_;
    if (track) {
        if (chg_stale_) {
            if (cudaMemsetAsync(chg_[0], 1, chg_bytes_, stream_) != cudaSuccess ||
                cudaMemsetAsync(chg_[1], 1, chg_bytes_, stream_) != cudaSuccess) {
                cudaGetLastError();
                return kKernelExecution;
            }
            chg_stale_ = false;
        }
        p.chg_prev = chg_[cur_];
        p.chg_out = chg_[cur_ ^ 1];
        p.skipped = &ctrl_->skipped;
    } else {
        chg_stale_ = true;
    }
    const int mode = (track ? kTrack : kPlain) | (has_peers() ? kP2P : kPlain);
#define EPIC_LAUNCH_2D(MATH, NT, MODE) \
    sweep2d_kernel<MATH, NT, MODE><<<grid, NT, smem, stream_>>>(tmap_[cur_], p, m)
#define EPIC_LAUNCH_2D_MODES(MATH, NT)                                   \
    switch (mode) {                                                      \
    case kPlain: EPIC_LAUNCH_2D(MATH, NT, kPlain); break;                \
    case kTrack: EPIC_LAUNCH_2D(MATH, NT, kTrack); break;                \
    case kP2P: EPIC_LAUNCH_2D(MATH, NT, kP2P); break;                    \
    default: EPIC_LAUNCH_2D(MATH, NT, kTrack | kP2P); break;             \
    }
    if (cfg_.math == MATH_STRICT) {
        StrictMath m;
        m.init(kLog4);
        if (NT_ == 512) {
            EPIC_LAUNCH_2D_MODES(StrictMath, 512)
        } else {
            EPIC_LAUNCH_2D_MODES(StrictMath, 256)
        }
    } else {
        FastMath m;
        m.ln2n = 1.3862943611198906f;
        if (NT_ == 512) {
            EPIC_LAUNCH_2D_MODES(FastMath, 512)
        } else {
            EPIC_LAUNCH_2D_MODES(FastMath, 256)
        }
    }
#undef EPIC_LAUNCH_2D_MODES
#undef EPIC_LAUNCH_2D
    if (track && count < 2) {
        // a pass of one half-sweep has only exercised one colour: its "nothing changed" proves nothing
        // about the other one, so the next pass must not skip on these flags
        if (cudaMemsetAsync(chg_[cur_ ^ 1], 1, chg_bytes_, stream_) != cudaSuccess) {
            cudaGetLastError();
            return kKernelExecution;
        }
    }
    launches_++;
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    cur_ ^= 1;
    if (kernel_sync) {
        pass_count_++;      // the kernel's edge tiles publish it to the neighbours
        return kSuccess;
    }
    return signal_peers();
}

int Field::launch_pass_3d(uint32_t it0, uint32_t count, bool check_last)
{
    Sweep3DParams p;
    memset(&p, 0, sizeof(p));
    p.dst = u_[cur_ ^ 1];
    p.freemask = freemask_;
    p.ctrl_done = &ctrl_->done;
    p.delta_bits = &ctrl_->delta_bits;
    p.pitch = pitch_;
    p.layer_floats = layer_floats_;
    p.mask_wpr = (uint32_t)mask_wpr_;
    p.m0 = (uint32_t)gm_[0];
    p.m1 = (uint32_t)gm_[1];
    p.m2 = (uint32_t)gm_[2];
    p.grow0 = grow0_;
    p.buf_layers = (uint32_t)buf_layers_;
    p.own_lo = (uint32_t)own_lo_;
    p.own_hi = (uint32_t)own_hi_;
    p.BH = (uint32_t)TH_;
    p.ntx = (uint32_t)((gm_[2] + k3OutW - 1) / k3OutW);
    p.nty = (uint32_t)((gm_[1] + (p.BH - 2 * k3HR) - 1) / (p.BH - 2 * k3HR));
    p.zchunk = zchunk_;
    const uint32_t ntz = (uint32_t)((rows_ + p.zchunk - 1) / p.zchunk);
    p.count = count;
    p.it0 = it0;
    p.check = check_last ? 1u : 0u;
    p.halo_layers = (uint32_t)std::min<uint64_t>(ghost_, rows_);
    if (peer_[0].on) {
        p.peer_up = peer_[0].u[cur_ ^ 1] + peer_[0].own_hi * layer_floats_;
    }
    if (peer_[1].on) {
        p.peer_down = peer_[1].u[cur_ ^ 1] + (peer_[1].own_lo - p.halo_layers) * layer_floats_;
    }
    if (wait_peers() != kSuccess) {
        return kKernelExecution;
    }
    const size_t smem = sweep3d_smem_bytes(p.BH);
    const uint32_t grid = p.ntx * p.nty * ntz;
    p.ntz = ntz;
    const bool track = tracking_ && skip_static_ && chg_[0] != nullptr && (size_t)grid <= chg_bytes_ && p.zchunk >= 2u * k3HR;
    if (track) {
        if (chg_stale_) {
            if (cudaMemsetAsync(chg_[0], 1, chg_bytes_, stream_) != cudaSuccess ||
                cudaMemsetAsync(chg_[1], 1, chg_bytes_, stream_) != cudaSuccess) {
                cudaGetLastError();
                return kKernelExecution;
            }
            chg_stale_ = false;
        }
        p.chg_prev = chg_[cur_];
        p.chg_out = chg_[cur_ ^ 1];
        p.skipped = &ctrl_->skipped;
    } else {
        chg_stale_ = true;
    }
    if (cfg_.math == MATH_STRICT) {
        StrictMath m;
        m.init(kLog6);
        if (track) sweep3d_kernel<StrictMath, true><<<grid, k3Threads, smem, stream_>>>(tmap_[cur_], p, m);
        else sweep3d_kernel<StrictMath><<<grid, k3Threads, smem, stream_>>>(tmap_[cur_], p, m);
    } else {
        FastMath m;
        m.ln2n = 1.791759469228055f;
        if (track) sweep3d_kernel<FastMath, true><<<grid, k3Threads, smem, stream_>>>(tmap_[cur_], p, m);
        else sweep3d_kernel<FastMath><<<grid, k3Threads, smem, stream_>>>(tmap_[cur_], p, m);
    }
    if (track && count < 2) {   // one colour only: see launch_pass_2d
        if (cudaMemsetAsync(chg_[cur_ ^ 1], 1, chg_bytes_, stream_) != cudaSuccess) {
            cudaGetLastError();
            return kKernelExecution;
        }
    }
    launches_++;
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    cur_ ^= 1;
    return signal_peers();
}

int Field::launch_pass(uint32_t it0, uint32_t count, bool check_last)
{
    return (n_ == 2) ? launch_pass_2d(it0, count, check_last) : launch_pass_3d(it0, count, check_last);
}

int Field::run(uint32_t it0, uint32_t count, bool check_last)
{
    DeviceGuard guard(cfg_.device);
    uint32_t done = 0;
    while (done < count) {
        const uint32_t c = std::min<uint32_t>((uint32_t)T_, count - done);
        const int r = launch_pass(it0 + done, c, check_last && (done + c == count));
        if (r != kSuccess) {
            return r;
        }
        done += c;
    }
    return kSuccess;
}

int Field::run_period(uint32_t it0, uint32_t count)
{
    // Peer-to-peer halos order passes between GPUs with stream memory operations, and the legacy default
    // stream cannot be captured: those cases launch directly.
    const bool direct = graphs_off_ || has_peers() || stream_ == nullptr || stream_ == cudaStreamLegacy ||
                        stream_ == cudaStreamPerThread || count < 2u * (uint32_t)T_;
    if (!direct) {
        const uint32_t key = ((uint32_t)cur_ << 1) | (it0 & 1u);
        PeriodGraph &g = graphs_[key];
        if (g.exec != nullptr && g.count != count) {
            cudaGraphExecDestroy(g.exec);
            g.exec = nullptr;
        }
        if (g.exec == nullptr) {
            const int cur_before = cur_;
            const uint64_t launches_before = launches_;
            const uint32_t passes_before = pass_count_;
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                ok = run(it0, count, true) == kSuccess;
                decide_period_kernel<<<1, 1, 0, stream_>>>(ctrl_, count, (uint32_t)cur_);
                launches_++;
                ok = (cudaStreamEndCapture(stream_, &graph) == cudaSuccess) && ok && graph != nullptr;
            }
            if (ok) {
                ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
            }
            if (graph != nullptr) {
                cudaGraphDestroy(graph);
            }
            if (ok) {
                g.key = key;
                g.count = count;
                g.cur_after = cur_;
                g.launches = launches_ - launches_before;
                g.passes = pass_count_ - passes_before;
            } else {
                cudaGetLastError();
                g.exec = nullptr;
                graphs_off_ = true;   // this driver / stream cannot capture: launch directly from now on
            }
            // nothing has run yet: rewind the host-side bookkeeping the capture advanced
            cur_ = cur_before;
            launches_ = launches_before;
            pass_count_ = passes_before;
        }
        if (g.exec != nullptr) {
            if (cudaGraphLaunch(g.exec, stream_) != cudaSuccess) {
                cudaGetLastError();
                return kKernelExecution;
            }
            cur_ = g.cur_after;
            launches_ += g.launches;
            pass_count_ += g.passes;
            return kSuccess;
        }
    }
    const int r = run(it0, count, true);
    if (r != kSuccess) {
        return r;
    }
    decide_period_kernel<<<1, 1, 0, stream_>>>(ctrl_, count, (uint32_t)cur_);
    launches_++;
    return cudaGetLastError() == cudaSuccess ? kSuccess : kKernelExecution;
}

int Field::read_delta(float *delta)
{
    DeviceGuard guard(cfg_.device);
    take_delta_kernel<<<1, 1, 0, stream_>>>(ctrl_, 0);
    launches_++;
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    if (cudaMemcpyAsync(&ctrl_host_[0], ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToHost;
    }
    if (cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    *delta = ctrl_host_[0].last_delta;
    skipped_tiles_ += ctrl_host_[0].taken_skipped;
    return kSuccess;
}

int Field::solve(float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iteration, float *delta)
{
    if (!(epsilon > 0.0f) || stagger == 0 || iteration == nullptr || delta == nullptr) {
        return kInvalidData;
    }
    DeviceGuard guard(cfg_.device);
    struct Tracking {   // static-tile skipping is on for the passes this function issues
        bool &flag;
        bool before;
        explicit Tracking(bool &f) : flag(f), before(f) { flag = true; }
        ~Tracking() { flag = before; }
    } tracking(tracking_);
    struct ClearCtrl {  // an early error return must not leave `done` set: later run() calls would silently do nothing
        Field *f;
        bool armed = true;
        ~ClearCtrl()
        {
            if (armed) {
                cudaStreamSynchronize(f->stream_);
                cudaMemsetAsync(f->ctrl_, 0, sizeof(Ctrl), f->stream_);
                cudaStreamSynchronize(f->stream_);
                cudaGetLastError();
            }
        }
    } clear_ctrl{this};
    if (cudaMemsetAsync(ctrl_, 0, sizeof(Ctrl), stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToDevice;
    }
    set_rule_kernel<<<1, 1, 0, stream_>>>(ctrl_, epsilon, m_max);
    launches_++;
    // Period 0 is the check sweep at iteration 0; period k >= 1 covers iterations
    // (k-1)*stagger+1 .. k*stagger and ends with the check sweep at k*stagger
    // (reference harmonic_gpu.cu:266-282).  Two periods are kept in flight.
    uint64_t it = 0;
    const Ctrl *fin = nullptr;
    for (uint64_t period = 0; fin == nullptr; ++period) {
        if (period >= 2) {
            const int slot = (int)((period - 2) % kSlots);
            if (cudaEventSynchronize(events_[slot]) != cudaSuccess) {
                cudaGetLastError();
                return kDeviceSynchronize;
            }
            if (ctrl_host_[slot].done) {
                fin = &ctrl_host_[slot];
                break;
            }
        }
        const uint32_t count = (period == 0) ? 1u : stagger;
        if (it + count > 0xffffffffull) {
            return kInvalidData;  // the reference's 32-bit iteration counter would wrap
        }
        int r;
        if (period == 0) {
            r = run((uint32_t)it, count, true);
            if (r == kSuccess) {
                decide_kernel<<<1, 1, 0, stream_>>>(ctrl_, epsilon, (uint32_t)(it + count), m_max, (uint32_t)cur_);
                launches_++;
                r = cudaGetLastError() == cudaSuccess ? kSuccess : kKernelExecution;
            }
        } else {
            r = run_period((uint32_t)it, count);
        }
        if (r != kSuccess) {
            return r;
        }
        it += count;
        const int slot = (int)(period % kSlots);
        if (cudaMemcpyAsync(&ctrl_host_[slot], ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
            cudaEventRecord(events_[slot], stream_) != cudaSuccess) {
            cudaGetLastError();
            return kMemcpyToHost;
        }
    }
    if (cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    cur_ = (int)fin->final_buffer;
    skipped_tiles_ += fin->skipped;
    *iteration = fin->final_iteration;
    *delta = fin->last_delta;
    clear_ctrl.armed = false;
    // leave the flag clear so that later update / update_and_check calls sweep again
    if (cudaMemsetAsync(ctrl_, 0, sizeof(Ctrl), stream_) != cudaSuccess || cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    return kSuccess;
}

static_assert(kViewSlabs >= kMaxSlabs && kDecideSlots >= kMaxSlabs, "a Grid's slabs must fit the device-side tables");

static FieldView2D grid_view(const std::vector<Field *> &slabs, void (Field::*add)(FieldView2D *) const)
{
    FieldView2D view = empty_view();
    for (const Field *f : slabs) {
        (f->*add)(&view);
    }
    return view;
}

int Field::potential_grid(const std::vector<Field *> &slabs, float x, float y, float *value)
{
    return potential_view(grid_view(slabs, &Field::add_to_view), x, y, value);
}

int Field::gradient_grid(const std::vector<Field *> &slabs, float x, float y, float cd, float *px, float *py)
{
    return gradient_view(grid_view(slabs, &Field::add_to_view), x, y, cd, px, py);
}

int Field::paths_grid(const std::vector<Field *> &slabs, uint32_t count, const float *starts, float step, float cd,
                      uint32_t max_length, int *ret, uint32_t *k, float **paths)
{
    return paths_view(grid_view(slabs, &Field::add_to_view), count, starts, step, cd, max_length, ret, k, paths);
}

// ------------------------------------------------------------------------------------------------
// Hooks for Grid::solve (a grid sharded over several slabs of this process)

int Field::default_sweeps_per_pass(unsigned n, const FieldConfig &cfg)
{
    if (n != 2) {
        return k3HR;
    }
    return cfg.sweeps_per_pass > 0 ? std::min(cfg.sweeps_per_pass, 8) : 4;
}

int Field::solve_begin(float epsilon, uint32_t m_max)
{
    DeviceGuard guard(cfg_.device);
    if (cudaMemsetAsync(ctrl_, 0, sizeof(Ctrl), stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToDevice;
    }
    set_rule_kernel<<<1, 1, 0, stream_>>>(ctrl_, epsilon, m_max);
    launches_++;
    return cudaGetLastError() == cudaSuccess ? kSuccess : kKernelExecution;
}

static DecideAllParams decide_params(Ctrl *ctrl, const DecideWiring &w, uint32_t tag, uint32_t count, uint32_t buffer, bool rule)
{
    DecideAllParams p;
    memset(&p, 0, sizeof(p));
    p.ctrl = ctrl;
    p.inbox = w.inbox[w.me];
    for (uint32_t i = 0; i < w.nslabs; ++i) {
        p.peer_inbox[i] = w.inbox[i];
    }
    p.nslabs = w.nslabs;
    p.me = w.me;
    p.tag = tag;
    p.count = count;
    p.buffer = buffer;
    p.use_ctrl_rule = rule ? 1u : 0u;
    return p;
}

int Field::publish_delta(const DecideWiring &w, uint32_t tag)
{
    DeviceGuard guard(cfg_.device);
    publish_delta_kernel<<<1, 32, 0, stream_>>>(decide_params(ctrl_, w, tag, 0, (uint32_t)cur_, true));
    launches_++;
    return cudaGetLastError() == cudaSuccess ? kSuccess : kKernelExecution;
}

int Field::decide_all(const DecideWiring &w, uint32_t tag, uint32_t count, bool rule)
{
    DeviceGuard guard(cfg_.device);
    decide_all_kernel<<<1, 32, 0, stream_>>>(decide_params(ctrl_, w, tag, count, (uint32_t)cur_, rule));
    launches_++;
    return cudaGetLastError() == cudaSuccess ? kSuccess : kKernelExecution;
}

int Field::snapshot(int slot)
{
    DeviceGuard guard(cfg_.device);
    if (cudaMemcpyAsync(&ctrl_host_[slot], ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaEventRecord(events_[slot], stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToHost;
    }
    return kSuccess;
}

int Field::wait_snapshot(int slot, SolveSnapshot *out)
{
    DeviceGuard guard(cfg_.device);
    if (cudaEventSynchronize(events_[slot]) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    const Ctrl &c = ctrl_host_[slot];
    out->done = c.done;
    out->final_iteration = c.final_iteration;
    out->final_buffer = c.final_buffer;
    out->skipped = c.skipped;
    out->failed = c.failed;
    out->last_delta = c.last_delta;
    return kSuccess;
}

int Field::solve_end(const SolveSnapshot &fin)
{
    DeviceGuard guard(cfg_.device);
    // this slab's own count of skipped tiles (the decision fields are the same on every slab)
    if (cudaMemcpyAsync(&ctrl_host_[0], ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    skipped_tiles_ += ctrl_host_[0].skipped;
    cur_ = (int)fin.final_buffer;
    chg_stale_ = true;
    if (cudaMemsetAsync(ctrl_, 0, sizeof(Ctrl), stream_) != cudaSuccess || cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    return kSuccess;
}

// ------------------------------------------------------------------------------------------------
// Sparse edits

int Field::set_cells_2d(uint32_t k, const uint32_t *v, const uint32_t *types)
{
    if (n_ != 2 || k == 0 || v == nullptr || types == nullptr) {
        return kInvalidData;
    }
    chg_stale_ = true;
    DeviceGuard guard(cfg_.device);
    // The reference's CPU twin applies the edits in order; when a cell appears more than once the
    // last edit wins.  last[i] = index of the last *valid* edit of the cell edit i targets.
    std::vector<uint32_t> last(k);
    {
        std::unordered_map<uint64_t, uint32_t> winner;
        winner.reserve((size_t)k * 2);
        for (uint32_t i = 0; i < k; ++i) {
            const uint32_t x = v[2 * i], y = v[2 * i + 1];
            if (x < gm_[1] && y < gm_[0] && types[i] <= 2u) {
                winner[((uint64_t)y << 32) | x] = i;
            }
        }
        for (uint32_t i = 0; i < k; ++i) {
            auto itw = winner.find(((uint64_t)v[2 * i + 1] << 32) | v[2 * i]);
            last[i] = (itw == winner.end()) ? i : itw->second;
        }
    }
    const size_t bytes = (size_t)k * 4 * sizeof(uint32_t);
    int r = ensure_staging(&staging_, &staging_bytes_, bytes);
    if (r != kSuccess) {
        return r;
    }
    uint32_t *d_v = (uint32_t *)staging_;
    uint32_t *d_types = d_v + 2 * (size_t)k;
    uint32_t *d_last = d_types + k;
    if (cudaMemcpyAsync(d_v, v, (size_t)k * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_) != cudaSuccess ||
        cudaMemcpyAsync(d_types, types, (size_t)k * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_) != cudaSuccess ||
        cudaMemcpyAsync(d_last, last.data(), (size_t)k * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_) !=
            cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToDevice;
    }
    set_cells_2d_kernel<<<(k + 255) / 256, 256, 0, stream_>>>(u_[cur_], freemask_, pitch_, (uint32_t)mask_wpr_,
                                                             (uint32_t)gm_[0], (uint32_t)gm_[1], grow0_,
                                                             (uint32_t)buf_layers_, k, d_v, d_types, d_last);
    launches_++;
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    if (cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    return kSuccess;
}

int Field::ingest_occupancy_2d(const signed char *host, uint64_t first, uint64_t layers, int threshold, int no_change)
{
    if (n_ != 2 || host == nullptr || !range_ok(grow0_, buf_layers_, first, layers)) {
        return kInvalidData;
    }
    chg_stale_ = true;
    DeviceGuard guard(cfg_.device);
    if (wait_peers() != kSuccess) {
        return kDeviceSynchronize;
    }
    const uint64_t m1 = gm_[1];
    const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(layers, (64ull << 20) / m1));
    int r = ensure_staging(&staging_, &staging_bytes_, chunk_rows * m1);
    if (r != kSuccess) {
        return r;
    }
    const uint32_t words = (uint32_t)((m1 + 31) / 32);
    for (uint64_t done = 0; done < layers; done += chunk_rows) {
        const uint64_t nrows = std::min(chunk_rows, layers - done);
        if (cudaMemcpyAsync(staging_, host + done * m1, nrows * m1, cudaMemcpyHostToDevice, stream_) != cudaSuccess) {
            cudaGetLastError();
            return kMemcpyToDevice;
        }
        const uint64_t b0 = (uint64_t)((int64_t)(first + done) - grow0_);
        const uint64_t threads = nrows * words * 32;
        reclassify_2d_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream_>>>(
            u_[cur_] + b0 * pitch_, freemask_ + b0 * mask_wpr_, pitch_, (uint32_t)mask_wpr_, (uint32_t)gm_[0], (uint32_t)m1,
            (int64_t)(first + done), (uint32_t)nrows, (const signed char *)staging_, (int64_t)(first + done), threshold,
            no_change, 0);
        launches_++;
        if (cudaGetLastError() != cudaSuccess) {
            return kKernelExecution;
        }
        if (cudaStreamSynchronize(stream_) != cudaSuccess) {   // the staging buffer is reused
            cudaGetLastError();
            return kDeviceSynchronize;
        }
    }
    return kSuccess;
}

int Field::reset_free_cells_2d()
{
    if (n_ != 2) {
        return kInvalidData;
    }
    chg_stale_ = true;
    DeviceGuard guard(cfg_.device);
    if (wait_peers() != kSuccess) {
        return kDeviceSynchronize;
    }
    const uint32_t words = (uint32_t)((gm_[1] + 31) / 32);
    const uint64_t threads = buf_layers_ * words * 32;
    reclassify_2d_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream_>>>(
        u_[cur_], freemask_, pitch_, (uint32_t)mask_wpr_, (uint32_t)gm_[0], (uint32_t)gm_[1], grow0_, (uint32_t)buf_layers_,
        nullptr, 0, 0, 0, 1);
    launches_++;
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    if (cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceSynchronize;
    }
    return kSuccess;
}

// ------------------------------------------------------------------------------------------------
// Streamlines

void Field::add_to_view(FieldView2D *view) const
{
    FieldView2D &f = *view;
    if (f.nslabs == 0) {
        f.pitch = pitch_;
        f.mask_wpr = (uint32_t)mask_wpr_;
        f.m0 = (uint32_t)gm_[0];
        f.m1 = (uint32_t)gm_[1];
    }
    const uint32_t i = f.nslabs++;
    f.u[i] = u_[cur_];
    f.mask[i] = freemask_;
    f.grow0[i] = grow0_;
    f.row_end[i] = (uint32_t)(row0_ + rows_);
}


int Field::potential_2d(float x, float y, float *value)
{
    if (n_ != 2 || value == nullptr || rows_ != gm_[0]) {
        return kInvalidData;
    }
    FieldView2D view = empty_view();
    add_to_view(&view);
    return potential_view(view, x, y, value);
}

int Field::potential_view(const FieldView2D &view, float x, float y, float *value)
{
    DeviceGuard guard(cfg_.device);
    int r = ensure_staging(&staging_, &staging_bytes_, 64);
    if (r != kSuccess) {
        return r;
    }
    float *d_out = (float *)staging_;
    int *d_ret = (int *)(d_out + 2);
    potential_gradient_kernel<<<1, 1, 0, stream_>>>(view, x, y, 0.0f, 0, d_out, d_ret);
    launches_++;
    float h[3];
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    if (cudaMemcpyAsync(h, staging_, 12, cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToHost;
    }
    int ret;
    memcpy(&ret, &h[2], 4);
    if (ret == kSuccess) {
        *value = h[0];
    }
    return ret;
}

int Field::gradient_2d(float x, float y, float cd, float *px, float *py)
{
    if (n_ != 2 || px == nullptr || py == nullptr || rows_ != gm_[0]) {
        return kInvalidData;
    }
    FieldView2D view = empty_view();
    add_to_view(&view);
    return gradient_view(view, x, y, cd, px, py);
}

int Field::gradient_view(const FieldView2D &view, float x, float y, float cd, float *px, float *py)
{
    DeviceGuard guard(cfg_.device);
    int r = ensure_staging(&staging_, &staging_bytes_, 64);
    if (r != kSuccess) {
        return r;
    }
    float *d_out = (float *)staging_;
    int *d_ret = (int *)(d_out + 2);
    potential_gradient_kernel<<<1, 1, 0, stream_>>>(view, x, y, cd, 1, d_out, d_ret);
    launches_++;
    float h[3];
    if (cudaGetLastError() != cudaSuccess) {
        return kKernelExecution;
    }
    if (cudaMemcpyAsync(h, staging_, 12, cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        return kMemcpyToHost;
    }
    int ret;
    memcpy(&ret, &h[2], 4);
    if (ret == kSuccess) {
        *px = h[0];
        *py = h[1];
    }
    return ret;
}

int Field::paths_2d(uint32_t count, const float *starts, float step, float cd, uint32_t max_length, int *ret,
                    uint32_t *k, float **paths)
{
    if (n_ != 2 || count == 0 || starts == nullptr || ret == nullptr || k == nullptr || paths == nullptr ||
        rows_ != gm_[0]) {
        return kInvalidData;
    }
    FieldView2D view = empty_view();
    add_to_view(&view);
    return paths_view(view, count, starts, step, cd, max_length, ret, k, paths);
}

int Field::paths_view(const FieldView2D &view, uint32_t count, const float *starts, float step, float cd,
                      uint32_t max_length, int *ret, uint32_t *k, float **paths)
{
    DeviceGuard guard(cfg_.device);
    // chunk: points emitted per path and launch
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(256, std::min<uint64_t>(65536, (32ull << 20) / ((uint64_t)count * 8)));
    const size_t out_bytes = (size_t)count * chunk * 2 * sizeof(float);
    const size_t st_bytes = round_up((size_t)count * sizeof(PathState), 256);
    const size_t em_bytes = round_up((size_t)count * sizeof(uint32_t), 256);
    const size_t in_bytes = round_up((size_t)count * 2 * sizeof(float), 256);
    unsigned char *base = nullptr;
    if (cudaMalloc(&base, out_bytes + st_bytes + em_bytes + in_bytes) != cudaSuccess) {
        cudaGetLastError();
        return kDeviceMalloc;
    }
    float *d_out = (float *)base;
    PathState *d_states = (PathState *)(base + out_bytes);
    uint32_t *d_emitted = (uint32_t *)(base + out_bytes + st_bytes);
    float *d_starts = (float *)(base + out_bytes + st_bytes + em_bytes);
    int result = kSuccess;
    std::vector<std::vector<float>> acc(count);
    std::vector<uint32_t> emitted(count);
    std::vector<PathState> states(count);
    std::vector<float> chunk_host;
    // `pathVector.size() < 2 * maxLength` is evaluated in 32-bit unsigned arithmetic by the reference
    // (harmonic_path_cpu.cpp:187), so the product wraps.
    const uint64_t max_floats = (uint64_t)(uint32_t)(2u * max_length);
    if (cudaMemcpyAsync(d_starts, starts, (size_t)count * 2 * sizeof(float), cudaMemcpyHostToDevice, stream_) !=
        cudaSuccess) {
        cudaGetLastError();
        result = kMemcpyToDevice;
    }
    bool running = true;
    for (uint32_t launch = 0; running && result == kSuccess; ++launch) {
        // a whole-grid field (one slab that starts at row 0) gets the kernel without the slab lookup
        if (view.nslabs == 1 && view.grow0[0] == 0) {
            path_2d_kernel<false><<<(count + kPathWarps - 1) / kPathWarps, 32 * kPathWarps, 0, stream_>>>(
                view, count, d_starts, step, cd, max_floats, chunk, d_states, d_out, d_emitted, launch == 0 ? 1u : 0u);
        } else {
            path_2d_kernel<true><<<(count + kPathWarps - 1) / kPathWarps, 32 * kPathWarps, 0, stream_>>>(
                view, count, d_starts, step, cd, max_floats, chunk, d_states, d_out, d_emitted, launch == 0 ? 1u : 0u);
        }
        launches_++;
        if (cudaGetLastError() != cudaSuccess) {
            result = kKernelExecution;
            break;
        }
        if (cudaMemcpyAsync(emitted.data(), d_emitted, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            stream_) != cudaSuccess ||
            cudaMemcpyAsync(states.data(), d_states, (size_t)count * sizeof(PathState), cudaMemcpyDeviceToHost,
                            stream_) != cudaSuccess ||
            cudaStreamSynchronize(stream_) != cudaSuccess) {
            cudaGetLastError();
            result = kMemcpyToHost;
            break;
        }
        running = false;
        // One copy per launch when many streamlines are traced (their segments are scattered on the host);
        // a single streamline copies just what it emitted.
        const bool bulk = count > 1;
        uint32_t most = 0;
        if (bulk) {
            for (uint32_t i = 0; i < count; ++i) {
                most = std::max(most, emitted[i]);
            }
            if (most > 0) {
                // the first `most` points of every streamline's segment, packed: one strided copy
                chunk_host.resize((size_t)count * most * 2);
                if (cudaMemcpy2D(chunk_host.data(), (size_t)most * 2 * sizeof(float), d_out, (size_t)chunk * 2 * sizeof(float),
                                 (size_t)most * 2 * sizeof(float), count, cudaMemcpyDeviceToHost) != cudaSuccess) {
                    cudaGetLastError();
                    result = kMemcpyToHost;
                    break;
                }
            }
        }
        for (uint32_t i = 0; i < count; ++i) {
            if (emitted[i] > 0) {
                const size_t old = acc[i].size();
                acc[i].resize(old + (size_t)emitted[i] * 2);
                if (bulk) {
                    memcpy(acc[i].data() + old, chunk_host.data() + (size_t)i * most * 2, (size_t)emitted[i] * 2 * sizeof(float));
                } else if (cudaMemcpy(acc[i].data() + old, d_out + (size_t)i * chunk * 2,
                                      (size_t)emitted[i] * 2 * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
                    cudaGetLastError();
                    result = kMemcpyToHost;
                }
            }
            if (states[i].status == -1) {
                running = true;
            }
        }
    }
    cudaFree(base);
    if (result != kSuccess) {
        return result;
    }
    for (uint32_t i = 0; i < count; ++i) {
        ret[i] = states[i].status;
        k[i] = 0;
        paths[i] = nullptr;
        if (states[i].status == kSuccess) {
            k[i] = (uint32_t)(acc[i].size() / 2);
            paths[i] = new float[acc[i].size()];   // released by the caller with delete[]
            memcpy(paths[i], acc[i].data(), acc[i].size() * sizeof(float));
        }
    }
    return kSuccess;
}

}  // namespace epic_b200
