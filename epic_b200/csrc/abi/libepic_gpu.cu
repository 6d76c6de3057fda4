// libepic_gpu.cu -- the GPU half of the libepic C ABI (include/epic/libepic.h) on top of
// engine/field.{h,cu}.
//
// Replaces, entry point for entry point, the host functions of the reference's
// libepic/src/harmonic/harmonic_gpu.cu:168-434, harmonic_model_gpu.cu:34-204 and
// harmonic_utilities_gpu.cu:66-138.  Same argument checks, same return codes, same
// "Error[function]: text" lines on stderr.
//
// Device residency.  The reference keeps three raw cudaMalloc pointers (d_m, d_u, d_locked) plus
// d_delta in the caller's struct; callers only zero-initialise them and hand the struct back
// (src/epic_nav_core_plugin.cpp:67-70, src/epic_navigation_node_harmonic.cpp:70-73), so here all
// four are handles to ONE `Context` that owns the Grid (engine/grid.h: one Field per device named by
// EPIC_DEVICES -- padded ping-pong buffers, 1-bit free mask, stream, control block -- a single one by default).  A handle is non-null exactly when the reference's pointer would be.
// Contexts are kept in a registry, so a stale or foreign pointer is recognised and rejected with
// EPIC_ERROR_INVALID_DATA instead of being dereferenced.
//
// Deferred sweeps.  harmonic_update_gpu does not need to return with the sweep finished (nothing on
// the host can observe the device field); it queues the sweep, and sweeps are issued to the GPU in
// passes of T (temporal blocking) as soon as T of them are queued.  Every call that observes or edits
// the field (update_and_check, get_potential_values, set_cells, path queries, uninitialize) first
// issues what is queued.  The observable results are those of the reference's call-by-call execution.
//
// There is no CPU fallback in this file: if CUDA is unavailable the calls fail with the reference's
// device error codes, and the reference's callers then choose harmonic_complete_cpu themselves
// (src/epic_nav_core_plugin.cpp:258-263).
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <unordered_set>

#include "../../../include/epic/libepic.h"
#include "../../../include/epic_b200.h"
#include "../engine/grid.h"

using epic_b200::FieldConfig;
using epic_b200::Grid;

namespace {

enum { BIT_DIM = 1, BIT_U = 2, BIT_LOCKED = 4, BIT_DELTA = 8 };

struct Context {
    Grid *field = nullptr;    // the grid on its device(s): one slab, or one per entry of EPIC_DEVICES
    unsigned n = 0;
    uint64_t m[3] = {0, 0, 0};
    unsigned live = 0;        // which of the four handles point here
    uint32_t queued = 0;      // update_gpu sweeps not yet issued
    uint32_t queued_from = 0; // iteration of the first queued sweep
};

std::mutex g_mutex;
std::unordered_set<Context *> g_contexts;

void complain(const char *fn, const char *text)
{
    fprintf(stderr, "Error[%s]: %s\n", fn, text);
}

bool registered(const void *p)
{
    return p != nullptr && g_contexts.count((Context *)p) != 0;
}

Context *find_context(const epic::Harmonic *h)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    const void *cands[4] = {h->d_m, h->d_u, h->d_locked, h->d_delta};
    for (const void *c : cands) {
        if (registered(c)) {
            return (Context *)c;
        }
    }
    return nullptr;
}

bool same_dims(const Context *c, const epic::Harmonic *h)
{
    if (c->n != h->n) {
        return false;
    }
    for (unsigned i = 0; i < h->n; ++i) {
        if (c->m[i] != h->m[i]) {
            return false;
        }
    }
    return true;
}

void drop_field(Context *c)
{
    delete c->field;
    c->field = nullptr;
    c->queued = 0;
}

// The context the struct's handles refer to, created (or re-dimensioned) from h->n / h->m as needed.
Context *obtain_context(epic::Harmonic *h)
{
    Context *c = find_context(h);
    if (c == nullptr) {
        c = new Context();
        std::lock_guard<std::mutex> lock(g_mutex);
        g_contexts.insert(c);
    }
    if (!same_dims(c, h)) {
        drop_field(c);
        c->n = h->n;
        for (unsigned i = 0; i < 3; ++i) {
            c->m[i] = (i < h->n) ? h->m[i] : 0;
        }
    }
    return c;
}

int ensure_field(Context *c)
{
    if (c->field != nullptr) {
        return EPIC_SUCCESS;
    }
    const FieldConfig cfg = epic_b200::config_from_env();
    return Grid::create(&c->field, c->n, c->m, cfg);
}

void release(Context *c, unsigned bit)
{
    c->live &= ~bit;
    if (c->live == 0) {
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            g_contexts.erase(c);
        }
        drop_field(c);
        delete c;
    }
}

// Issue queued update sweeps.  `all` = false keeps a remainder smaller than one pass queued.
int issue_queued(Context *c, bool all)
{
    if (c->field == nullptr || c->queued == 0) {
        return EPIC_SUCCESS;
    }
    const uint32_t T = (uint32_t)c->field->sweeps_per_pass();
    const uint32_t n = all ? c->queued : (c->queued / T) * T;
    if (n == 0) {
        return EPIC_SUCCESS;
    }
    const int r = c->field->run(c->queued_from, n, false);
    c->queued_from += n;
    c->queued -= n;
    return r;
}

bool dims_supported(const epic::Harmonic *h)
{
    return h->n == 2 || h->n == 3;
}

}  // namespace

namespace epic {

// ---- harmonic_model_gpu.cu:34-204 ----------------------------------------------------------------

int harmonic_initialize_dimension_size_gpu(Harmonic *harmonic)
{
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr) {
        complain("harmonic_initialize_dimension_size_gpu", "Invalid input.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (!dims_supported(harmonic)) {
        complain("harmonic_initialize_dimension_size_gpu", "Only n = 2 and n = 3 are implemented.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = obtain_context(harmonic);
    c->live |= BIT_DIM;
    harmonic->d_m = (unsigned int *)c;
    return EPIC_SUCCESS;
}

int harmonic_uninitialize_dimension_size_gpu(Harmonic *harmonic)
{
    if (harmonic == nullptr) {
        return EPIC_ERROR_INVALID_DATA;
    }
    if (Context *c = find_context(harmonic)) {
        if (harmonic->d_m == (unsigned int *)c) {
            release(c, BIT_DIM);
        }
    }
    harmonic->d_m = nullptr;
    return EPIC_SUCCESS;
}

int harmonic_initialize_potential_values_gpu(Harmonic *harmonic)
{
    const char *fn = "harmonic_initialize_potential_values_gpu";
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr || harmonic->u == nullptr) {
        complain(fn, "Invalid input.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (!dims_supported(harmonic)) {
        complain(fn, "Only n = 2 and n = 3 are implemented.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = obtain_context(harmonic);
    int r = ensure_field(c);
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to allocate device-side memory for the potential values.");
        if (c->live == 0) {
            release(c, 0);
        }
        return (r == EPIC_ERROR_INVALID_DATA) ? r : EPIC_ERROR_DEVICE_MALLOC;
    }
    c->queued = 0;
    r = c->field->upload_u(harmonic->u);
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to copy memory from host to device for the potential values.");
        if (c->live == 0) {
            release(c, 0);
        }
        return EPIC_ERROR_MEMCPY_TO_DEVICE;
    }
    c->live |= BIT_U;
    harmonic->d_u = (float *)c;
    return EPIC_SUCCESS;
}

int harmonic_uninitialize_potential_values_gpu(Harmonic *harmonic)
{
    if (harmonic == nullptr) {
        return EPIC_ERROR_INVALID_DATA;
    }
    if (Context *c = find_context(harmonic)) {
        if (harmonic->d_u == (float *)c) {
            release(c, BIT_U);
        }
    }
    harmonic->d_u = nullptr;
    return EPIC_SUCCESS;
}

int harmonic_initialize_locked_gpu(Harmonic *harmonic)
{
    const char *fn = "harmonic_initialize_locked_gpu";
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr || harmonic->locked == nullptr) {
        complain(fn, "Invalid input.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (!dims_supported(harmonic)) {
        complain(fn, "Only n = 2 and n = 3 are implemented.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = obtain_context(harmonic);
    int r = ensure_field(c);
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to allocate device-side memory for the locked cells.");
        if (c->live == 0) {
            release(c, 0);
        }
        return (r == EPIC_ERROR_INVALID_DATA) ? r : EPIC_ERROR_DEVICE_MALLOC;
    }
    r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->upload_locked(harmonic->locked);
    }
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to copy memory from host to device for the locked cells.");
        if (c->live == 0) {
            release(c, 0);
        }
        return EPIC_ERROR_MEMCPY_TO_DEVICE;
    }
    c->live |= BIT_LOCKED;
    harmonic->d_locked = (unsigned int *)c;
    return EPIC_SUCCESS;
}

int harmonic_uninitialize_locked_gpu(Harmonic *harmonic)
{
    if (harmonic == nullptr) {
        return EPIC_ERROR_INVALID_DATA;
    }
    if (Context *c = find_context(harmonic)) {
        if (harmonic->d_locked == (unsigned int *)c) {
            release(c, BIT_LOCKED);
        }
    }
    harmonic->d_locked = nullptr;
    return EPIC_SUCCESS;
}

// The context behind a struct whose u and locked are resident, or null.
static Context *resident(Harmonic *harmonic)
{
    if (harmonic == nullptr || harmonic->d_u == nullptr || harmonic->d_locked == nullptr) {
        return nullptr;
    }
    Context *c = find_context(harmonic);
    if (c == nullptr || c->field == nullptr || (c->live & (BIT_U | BIT_LOCKED)) != (BIT_U | BIT_LOCKED) ||
        harmonic->d_u != (float *)c || harmonic->d_locked != (unsigned int *)c) {
        return nullptr;
    }
    return c;
}

int harmonic_update_model_gpu(Harmonic *harmonic)
{
    const char *fn = "harmonic_update_model_gpu";
    Context *c = resident(harmonic);
    if (c == nullptr || harmonic->n == 0 || harmonic->m == nullptr || harmonic->u == nullptr ||
        harmonic->locked == nullptr || !same_dims(c, harmonic)) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    c->queued = 0;  // the field those sweeps would have produced is overwritten below
    if (c->field->upload_u(harmonic->u) != EPIC_SUCCESS) {
        complain(fn, "Failed to copy memory from host to device for the potential values.");
        return EPIC_ERROR_MEMCPY_TO_DEVICE;
    }
    if (c->field->upload_locked(harmonic->locked) != EPIC_SUCCESS) {
        complain(fn, "Failed to copy memory from host to device for the locked cells.");
        return EPIC_ERROR_MEMCPY_TO_DEVICE;
    }
    return EPIC_SUCCESS;
}

// ---- harmonic_gpu.cu:168-434 ---------------------------------------------------------------------

int harmonic_initialize_gpu(Harmonic *harmonic, unsigned int numThreads)
{
    (void)numThreads;
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr || harmonic->d_delta != nullptr) {
        complain("harmonic_initialize_gpu", "Invalid input.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (!dims_supported(harmonic)) {
        complain("harmonic_initialize_gpu", "Only n = 2 and n = 3 are implemented.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = obtain_context(harmonic);
    c->live |= BIT_DELTA;
    harmonic->d_delta = (float *)c;
    return EPIC_SUCCESS;
}

int harmonic_uninitialize_gpu(Harmonic *harmonic)
{
    if (harmonic == nullptr) {
        return EPIC_ERROR_INVALID_DATA;
    }
    int result = EPIC_SUCCESS;
    if (harmonic->d_delta != nullptr) {
        Context *c = find_context(harmonic);
        if (c != nullptr && harmonic->d_delta == (float *)c) {
            if (issue_queued(c, true) != EPIC_SUCCESS) {
                result = EPIC_ERROR_DEVICE_FREE;
            }
            release(c, BIT_DELTA);
        }
    }
    harmonic->d_delta = nullptr;
    return result;
}

int harmonic_update_gpu(Harmonic *harmonic, unsigned int numThreads)
{
    (void)numThreads;
    Context *c = resident(harmonic);
    if (c == nullptr) {
        complain("harmonic_update_gpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (c->queued == 0) {
        c->queued_from = harmonic->currentIteration;
    } else if (c->queued_from + c->queued != harmonic->currentIteration) {
        // the caller moved currentIteration: what is queued belongs to the old numbering
        const int r = issue_queued(c, true);
        if (r != EPIC_SUCCESS) {
            complain("harmonic_update_gpu", "Failed to execute the 'Gauss-Seidel update' kernel.");
            return EPIC_ERROR_KERNEL_EXECUTION;
        }
        c->queued_from = harmonic->currentIteration;
    }
    c->queued++;
    if (issue_queued(c, false) != EPIC_SUCCESS) {
        complain("harmonic_update_gpu", "Failed to execute the 'Gauss-Seidel update' kernel.");
        return EPIC_ERROR_KERNEL_EXECUTION;
    }
    harmonic->currentIteration++;
    return EPIC_SUCCESS;
}

int harmonic_update_and_check_gpu(Harmonic *harmonic, unsigned int numThreads)
{
    (void)numThreads;
    const char *fn = "harmonic_update_and_check_gpu";
    Context *c = resident(harmonic);
    if (c == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r;
    if (c->queued != 0 && c->queued_from + c->queued == harmonic->currentIteration) {
        // the check sweep closes the pass that holds the queued sweeps
        const uint32_t from = c->queued_from, n = c->queued;
        c->queued = 0;
        r = c->field->run(from, n + 1, true);
    } else {
        r = issue_queued(c, true);
        if (r == EPIC_SUCCESS) {
            r = c->field->run(harmonic->currentIteration, 1, true);
        }
    }
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to execute the 'Gauss-Seidel update' kernel.");
        return EPIC_ERROR_KERNEL_EXECUTION;
    }
    r = c->field->read_delta(&harmonic->delta);
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to copy memory from device to host for the max delta.");
        return (r == EPIC_ERROR_DEVICE_SYNCHRONIZE) ? r : EPIC_ERROR_MEMCPY_TO_HOST;
    }
    harmonic->currentIteration++;
    return (harmonic->delta < harmonic->epsilon) ? EPIC_SUCCESS_AND_CONVERGED : EPIC_SUCCESS;
}

int harmonic_get_potential_values_gpu(Harmonic *harmonic)
{
    const char *fn = "harmonic_get_potential_values_gpu";
    if (harmonic == nullptr || harmonic->u == nullptr || harmonic->d_u == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = find_context(harmonic);
    if (c == nullptr || c->field == nullptr || harmonic->d_u != (float *)c) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (issue_queued(c, true) != EPIC_SUCCESS || c->field->download_u(harmonic->u) != EPIC_SUCCESS) {
        complain(fn, "Failed to copy memory from device to host for the potential values.");
        return EPIC_ERROR_MEMCPY_TO_HOST;
    }
    return EPIC_SUCCESS;
}

int harmonic_execute_gpu(Harmonic *harmonic, unsigned int numThreads)
{
    const char *fn = "harmonic_execute_gpu";
    if (harmonic == nullptr || harmonic->m == nullptr || harmonic->u == nullptr || harmonic->locked == nullptr ||
        harmonic->epsilon <= 0.0 || harmonic->d_m == nullptr || harmonic->d_u == nullptr ||
        harmonic->d_locked == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    if (numThreads % 32 != 0) {
        complain(fn, "Must specficy a number of threads divisible by 32 (the number of threads in a warp).");
        return EPIC_ERROR_INVALID_CUDA_PARAM;
    }
    Context *c = resident(harmonic);
    if (c == nullptr || !same_dims(c, harmonic) || harmonic->numIterationsToStaggerCheck == 0) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    // Sweeps queued by earlier harmonic_update_gpu calls have already been applied to d_u in the reference
    // (execute restarts the iteration count, not the field): issue them before the solve starts from them.
    if (issue_queued(c, true) != EPIC_SUCCESS) {
        complain(fn, "Failed to execute the 'Gauss-Seidel update' kernel.");
        return EPIC_ERROR_KERNEL_EXECUTION;
    }
    harmonic->currentIteration = 0;

    int result = harmonic_initialize_gpu(harmonic, numThreads);
    if (result != EPIC_SUCCESS) {
        complain(fn, "Failed to initialize GPU variables.");
        return result;
    }
    // information must be able to cross the whole grid before convergence is accepted
    uint32_t mMax = 0;
    for (unsigned i = 0; i < harmonic->n; ++i) {
        mMax = harmonic->m[i] > mMax ? harmonic->m[i] : mMax;
    }
    harmonic->delta = harmonic->epsilon + 1.0f;

    uint32_t iterations = 0;
    float delta = 0.0f;
    result = c->field->solve(harmonic->epsilon, harmonic->numIterationsToStaggerCheck, mMax, &iterations, &delta);
    if (result != EPIC_SUCCESS) {
        complain(fn, "Failed to perform the Gauss-Seidel update and check step.");
        return result;
    }
    harmonic->currentIteration = iterations;
    harmonic->delta = delta;

    result = harmonic_get_potential_values_gpu(harmonic);
    if (result != EPIC_SUCCESS) {
        complain(fn, "Failed to get all the potential values.");
        return result;
    }
    result = harmonic_uninitialize_gpu(harmonic);
    if (result != EPIC_SUCCESS) {
        complain(fn, "Failed to uninitialize GPU variables.");
        return result;
    }
    return EPIC_SUCCESS;
}

int harmonic_complete_gpu(Harmonic *harmonic, unsigned int numThreads)
{
    if (harmonic == nullptr) {
        complain("harmonic_complete_gpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int result = harmonic_initialize_dimension_size_gpu(harmonic);
    if (result != EPIC_SUCCESS) {
        return result;
    }
    result = harmonic_initialize_potential_values_gpu(harmonic);
    if (result != EPIC_SUCCESS) {
        return result;
    }
    result = harmonic_initialize_locked_gpu(harmonic);
    if (result != EPIC_SUCCESS) {
        return result;
    }
    result = harmonic_execute_gpu(harmonic, numThreads);
    if (result != EPIC_SUCCESS) {
        return result;
    }
    result = EPIC_SUCCESS;
    if (harmonic_uninitialize_dimension_size_gpu(harmonic) != EPIC_SUCCESS) {
        result = EPIC_ERROR_DEVICE_FREE;
    }
    if (harmonic_uninitialize_potential_values_gpu(harmonic) != EPIC_SUCCESS) {
        result = EPIC_ERROR_DEVICE_FREE;
    }
    if (harmonic_uninitialize_locked_gpu(harmonic) != EPIC_SUCCESS) {
        result = EPIC_ERROR_DEVICE_FREE;
    }
    return result;
}

// ---- harmonic_utilities_gpu.cu:66-138 --------------------------------------------------------------

int harmonic_utilities_set_cells_2d_gpu(Harmonic *harmonic, unsigned int numThreads, unsigned int k, unsigned int *v,
                                        unsigned int *types)
{
    (void)numThreads;
    const char *fn = "harmonic_utilities_set_cells_2d_gpu";
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr || harmonic->u == nullptr ||
        harmonic->locked == nullptr || k == 0 || v == nullptr || types == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = resident(harmonic);
    if (c == nullptr || c->n != 2) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->set_cells_2d(k, v, types);
    }
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to execute the 'set cells' kernel.");
    }
    return r;
}

// ---- extensions: dense map ingest on the device-resident field ----------------------------------------

int harmonic_utilities_set_occupancy_grid_2d_gpu(Harmonic *harmonic, const signed char *data, int obstacleThreshold,
                                                 int noChangeValue)
{
    const char *fn = "harmonic_utilities_set_occupancy_grid_2d_gpu";
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr || data == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = resident(harmonic);
    if (c == nullptr || c->n != 2) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->ingest_occupancy_2d(data, obstacleThreshold, noChangeValue);
    }
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to classify the occupancy grid on the device.");
    }
    return r;
}

int harmonic_utilities_reset_free_cells_2d_gpu(Harmonic *harmonic)
{
    const char *fn = "harmonic_utilities_reset_free_cells_2d_gpu";
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = resident(harmonic);
    if (c == nullptr || c->n != 2) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->reset_free_cells_2d();
    }
    if (r != EPIC_SUCCESS) {
        complain(fn, "Failed to reset the free cells on the device.");
    }
    return r;
}

// ---- extensions: streamlines on the device-resident field -------------------------------------------

int harmonic_compute_potential_2d_gpu(Harmonic *harmonic, float x, float y, float &potential)
{
    Context *c = resident(harmonic);
    if (c == nullptr || c->n != 2) {
        complain("harmonic_compute_potential_2d_gpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->potential_2d(x, y, &potential);
    }
    if (r == EPIC_ERROR_INVALID_LOCATION) {
        complain("harmonic_compute_potential_2d_gpu", "Invalid location.");
    }
    return r;
}

int harmonic_compute_gradient_2d_gpu(Harmonic *harmonic, float x, float y, float cdPrecision, float &partialX,
                                     float &partialY)
{
    Context *c = resident(harmonic);
    if (c == nullptr || c->n != 2) {
        complain("harmonic_compute_gradient_2d_gpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->gradient_2d(x, y, cdPrecision, &partialX, &partialY);
    }
    if (r == EPIC_ERROR_INVALID_GRADIENT) {
        complain("harmonic_compute_gradient_2d_gpu", "Failed to compute potential values.");
    }
    return r;
}

int harmonic_compute_paths_2d_gpu(Harmonic *harmonic, unsigned int numPaths, const float *starts, float stepSize,
                                  float cdPrecision, unsigned int maxLength, int *results, unsigned int *k,
                                  float **paths)
{
    Context *c = resident(harmonic);
    if (c == nullptr || c->n != 2 || numPaths == 0 || starts == nullptr || results == nullptr || k == nullptr ||
        paths == nullptr) {
        complain("harmonic_compute_paths_2d_gpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    int r = issue_queued(c, true);
    if (r == EPIC_SUCCESS) {
        r = c->field->paths_2d(numPaths, starts, stepSize, cdPrecision, maxLength, results, k, paths);
    }
    return r;
}

int harmonic_compute_path_2d_gpu(Harmonic *harmonic, float x, float y, float stepSize, float cdPrecision,
                                 unsigned int maxLength, unsigned int &k, float *&path)
{
    const char *fn = "harmonic_compute_path_2d_gpu";
    if (path != nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const float start[2] = {x, y};
    int ret = EPIC_SUCCESS;
    unsigned int kk = 0;
    float *p = nullptr;
    const int r = harmonic_compute_paths_2d_gpu(harmonic, 1, start, stepSize, cdPrecision, maxLength, &ret, &kk, &p);
    if (r != EPIC_SUCCESS) {
        return r;
    }
    if (ret == EPIC_ERROR_INVALID_LOCATION) {
        complain(fn, "Invalid location.");
    } else if (ret == EPIC_ERROR_INVALID_GRADIENT) {
        complain(fn, "Could not compute gradient.");
    } else if (ret == EPIC_ERROR_INVALID_PATH) {
        complain(fn, "Could not compute a valid path.");
    }
    if (ret == EPIC_SUCCESS) {
        k = kk;
        path = p;
    }
    return ret;
}

int harmonic_compute_path_poses_2d_gpu(Harmonic *harmonic, float x, float y, float stepSize, float cdPrecision,
                                       unsigned int maxLength, float originX, float originY, float resolution,
                                       unsigned int &k, float *&poses)
{
    if (poses != nullptr) {
        complain("harmonic_compute_path_poses_2d_gpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    float *raw = nullptr;
    unsigned int kk = 0;
    const int r = harmonic_compute_path_2d_gpu(harmonic, x, y, stepSize, cdPrecision, maxLength, kk, raw);
    if (r != EPIC_SUCCESS) {
        return r;
    }
    // the streamline is traced on the device-resident field; the pose arithmetic stays on the host so that
    // the yaw is the host libm's atan2, as in the callers
    poses = new float[3 * (size_t)kk];
    harmonic_path_to_poses_2d(raw, kk, originX, originY, resolution, poses);
    delete[] raw;
    k = kk;
    return EPIC_SUCCESS;
}

}  // namespace epic

extern "C" int epic_b200_harmonic_stats(const void *harmonic, epic_b200_stats *out)
{
    if (harmonic == nullptr || out == nullptr) {
        return EPIC_ERROR_INVALID_DATA;
    }
    Context *c = find_context((const epic::Harmonic *)harmonic);
    if (c == nullptr || c->field == nullptr) {
        return EPIC_ERROR_INVALID_DATA;
    }
    const epic_b200::GridStats s = c->field->stats();
    memset(out, 0, sizeof(*out));
    out->slabs = s.slabs;
    out->last_solve_iterations = s.last_solve_iterations;
    out->last_solve_delta = s.last_solve_delta;
    out->last_solve_seconds = s.last_solve_seconds;
    out->launches = s.launches;
    for (uint32_t i = 0; i < s.slabs && i < 16; ++i) {
        out->skipped_tiles[i] = s.skipped_by_slab[i];
    }
    return EPIC_SUCCESS;
}
