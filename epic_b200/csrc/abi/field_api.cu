// field_api.cu -- the slab-level C ABI declared in include/epic_b200.h, a thin shell over Field.
#include "../../../include/epic_b200.h"

#include <math.h>
#include <string.h>

#include <vector>

#include "../engine/field.h"
#include "../kernels/math_policies.cuh"

using epic_b200::Field;
using epic_b200::FieldConfig;

struct epic_b200_field {
    Field *impl;
};

namespace {

// out[i] = strict expf / logf of the float whose bit pattern is first + i*stride
__global__ void selftest_math_kernel(uint32_t first, uint32_t count, uint32_t stride, int which, float *out)
{
    __shared__ epic_b200::MathTables tables;
    epic_b200::load_math_tables(&tables, threadIdx.x, blockDim.x);
    __syncthreads();
    epic_b200::StrictMath math;
    math.init(epic_b200::kLog4);
    math.bind(&tables);
    // exp_nonpos is warp-synchronous (its table lookup is a shuffle): every lane evaluates, only the store is guarded
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const float x = __uint_as_float(first + (i < count ? i : 0u) * stride);
    const float e = math.exp_nonpos(which == 0 ? x : 0.0f);
    const float l = math.log_sum(which == 0 ? 1.0f : x);
    if (i < count) {
        out[i] = which == 0 ? e : l;
    }
}

}  // namespace

extern "C" {

int epic_b200_field_create(epic_b200_field **out, unsigned int n, const uint64_t *m, uint64_t row0, uint64_t rows,
                           unsigned int ghost, int math, int device, void *stream, int use_stream)
{
    if (out == nullptr) {
        return 2;
    }
    *out = nullptr;
    FieldConfig cfg = epic_b200::config_from_env();
    if (math == EPIC_B200_MATH_STRICT) {
        cfg.math = epic_b200::MATH_STRICT;
    } else if (math == EPIC_B200_MATH_FAST) {
        cfg.math = epic_b200::MATH_FAST;
    }
    if (device >= 0) {
        cfg.device = device;
    }
    cfg.stream = (cudaStream_t)stream;
    cfg.use_stream = use_stream != 0;
    Field *impl = nullptr;
    const int r = Field::create(&impl, n, m, row0, rows, ghost, cfg);
    if (r != 0) {
        return r;
    }
    *out = new epic_b200_field{impl};
    return 0;
}

void epic_b200_field_destroy(epic_b200_field *f)
{
    if (f != nullptr) {
        delete f->impl;
        delete f;
    }
}

int epic_b200_field_info(epic_b200_field *f, epic_b200_info *info)
{
    if (f == nullptr || info == nullptr) {
        return 2;
    }
    info->pitch = f->impl->pitch();
    info->layer_floats = f->impl->layer_floats();
    info->launches = f->impl->launches();
    info->device_bytes = f->impl->device_bytes();
    info->sweeps_per_pass = (uint32_t)f->impl->sweeps_per_pass();
    info->tile_rows = (uint32_t)f->impl->tile_rows();
    info->math = (uint32_t)f->impl->math();
    info->device = f->impl->device();
    info->skipped_tiles = f->impl->skipped_tiles();
    return 0;
}

int epic_b200_field_set_tracking(epic_b200_field *f, int on)
{
    if (f == nullptr) {
        return 2;
    }
    f->impl->set_tracking(on != 0);
    return 0;
}

int epic_b200_field_upload_u(epic_b200_field *f, const float *host, uint64_t first, uint64_t layers)
{
    return f ? f->impl->upload_u(host, first, layers) : 2;
}
int epic_b200_field_upload_locked(epic_b200_field *f, const uint32_t *host, uint64_t first, uint64_t layers)
{
    return f ? f->impl->upload_locked(host, first, layers) : 2;
}
int epic_b200_field_download_u(epic_b200_field *f, float *host, uint64_t first, uint64_t layers)
{
    return f ? f->impl->download_u(host, first, layers) : 2;
}
int epic_b200_field_download_locked(epic_b200_field *f, uint32_t *host, uint64_t first, uint64_t layers)
{
    return f ? f->impl->download_locked(host, first, layers) : 2;
}
int epic_b200_field_run(epic_b200_field *f, uint32_t it0, uint32_t count, int check_last)
{
    return f ? f->impl->run(it0, count, check_last != 0) : 2;
}
int epic_b200_field_read_delta(epic_b200_field *f, float *delta)
{
    return (f && delta) ? f->impl->read_delta(delta) : 2;
}
int epic_b200_field_solve(epic_b200_field *f, float epsilon, uint32_t stagger, uint32_t m_max, uint32_t *iterations,
                          float *delta)
{
    return f ? f->impl->solve(epsilon, stagger, m_max, iterations, delta) : 2;
}
int epic_b200_field_sync(epic_b200_field *f)
{
    return f ? f->impl->sync() : 2;
}
void *epic_b200_field_layer_ptr(epic_b200_field *f, int64_t layer)
{
    return f ? (void *)f->impl->layer_ptr(layer) : nullptr;
}
int epic_b200_field_peer_export(epic_b200_field *f, void *blob, uint64_t blob_bytes)
{
    if (f == nullptr || blob == nullptr || blob_bytes < sizeof(epic_b200::PeerInfo)) {
        return 2;
    }
    return f->impl->peer_export((epic_b200::PeerInfo *)blob);
}
int epic_b200_field_set_peer_ipc(epic_b200_field *f, int dir, const void *blob, uint64_t blob_bytes)
{
    if (f == nullptr || blob == nullptr || blob_bytes < sizeof(epic_b200::PeerInfo)) {
        return 2;
    }
    return f->impl->set_peer_ipc(dir, (const epic_b200::PeerInfo *)blob);
}
int epic_b200_field_set_peer_local(epic_b200_field *f, int dir, epic_b200_field *other)
{
    if (f == nullptr || other == nullptr) {
        return 2;
    }
    return f->impl->set_peer_local(dir, other->impl);
}
int epic_b200_field_set_cells_2d(epic_b200_field *f, uint32_t k, const uint32_t *v, const uint32_t *types)
{
    return f ? f->impl->set_cells_2d(k, v, types) : 2;
}
int epic_b200_field_potential_2d(epic_b200_field *f, float x, float y, float *value)
{
    return f ? f->impl->potential_2d(x, y, value) : 2;
}
int epic_b200_field_gradient_2d(epic_b200_field *f, float x, float y, float cd, float *px, float *py)
{
    return f ? f->impl->gradient_2d(x, y, cd, px, py) : 2;
}
int epic_b200_field_paths_2d(epic_b200_field *f, uint32_t count, const float *starts, float step, float cd,
                             uint32_t max_length, int *results, uint32_t *k, float **paths)
{
    return f ? f->impl->paths_2d(count, starts, step, cd, max_length, results, k, paths) : 2;
}
void epic_b200_free_path(float *path)
{
    delete[] path;
}
int epic_b200_selftest_math(uint32_t stride, uint64_t *exp_checked, uint64_t *exp_mismatches, uint64_t *log_checked,
                            uint64_t *log_mismatches)
{
    if (stride == 0 || !exp_checked || !exp_mismatches || !log_checked || !log_mismatches) {
        return 2;
    }
    const uint32_t chunk = 1u << 24;
    float *d_out = nullptr;
    if (cudaMalloc(&d_out, chunk * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return 4;
    }
    std::vector<float> host(chunk);
    // which 0: every x <= 0 (bit patterns 0x80000000 .. 0xff800000); which 1: every x in [1, 8] (the log
    // argument is a sum of at most 6 terms, each <= 1 and one of them == 1)
    const uint64_t lo[2] = {0x80000000ull, 0x3f800000ull}, hi[2] = {0xff800000ull + 1, 0x41000000ull + 1};
    uint64_t checked[2] = {0, 0}, bad[2] = {0, 0};
    int result = 0;
    for (int which = 0; which < 2 && result == 0; ++which) {
        for (uint64_t b = lo[which]; b < hi[which] && result == 0; b += (uint64_t)chunk * stride) {
            const uint64_t n = std::min<uint64_t>(chunk, (hi[which] - b + stride - 1) / stride);
            selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256>>>((uint32_t)b, (uint32_t)n, stride, which, d_out);
            if (cudaMemcpy(host.data(), d_out, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
                cudaGetLastError();
                result = 6;
                break;
            }
            for (uint64_t i = 0; i < n; ++i) {
                const uint32_t bits = (uint32_t)(b + i * stride);
                float x, want;
                memcpy(&x, &bits, 4);
                want = which == 0 ? expf(x) : logf(x);
                checked[which]++;
                if (memcmp(&want, &host[i], 4) != 0 && !(want != want && host[i] != host[i])) {
                    bad[which]++;
                }
            }
        }
    }
    cudaFree(d_out);
    *exp_checked = checked[0];
    *exp_mismatches = bad[0];
    *log_checked = checked[1];
    *log_mismatches = bad[1];
    return result;
}

const char *epic_b200_version(void)
{
    return "epic_b200 0.1 sm_100a";
}

}  // extern "C"
