// libepic_cpu.cpp -- the host-side entry points of the libepic C ABI (include/epic/libepic.h).
//
// libepic's ABI contains CPU functions that callers select explicitly: the log-space CPU solver
// (reference libepic/src/harmonic/harmonic_cpu.cpp:38-220), the host twin of set_cells
// (harmonic_utilities_cpu.cpp:38-76), streamline extraction on the host copy of u
// (harmonic_path_cpu.cpp:41-232) and the legacy linear-space SOR code
// (harmonic_legacy_cpu.cpp:36-141, harmonic_legacy_path_cpu.cpp:41-235).  The ROS plugin calls
// harmonic_compute_path_2d_cpu on every plan (src/epic_nav_core_plugin.cpp:298) and the Python
// wrapper binds all of them at import (python/epic/epic_harmonic.py:61-124), so a drop-in library
// must export them with identical results.  None of the *_gpu entry points calls into this file.
//
// Built without -march / -ffast-math and with -ffp-contract=off: every float multiply and add is a
// separate IEEE operation, as in the reference's own build (libepic/Makefile:2).
#include <math.h>
#include <stdint.h>
#include <cmath>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "../../../include/epic/libepic.h"

namespace {

void complain(const char *fn, const char *text)
{
    fprintf(stderr, "Error[%s]: %s\n", fn, text);
}

// One red-black half-sweep over an n-dimensional grid (n = 2 or 3) with NB = 2n neighbours.
// Stride table: st[d] = distance in cells between neighbours along dimension d.
template <int N>
void half_sweep(epic::Harmonic *h, bool check)
{
    const unsigned int *m = h->m;
    float *u = h->u;
    const unsigned int *locked = h->locked;
    if (check) {
        h->delta = 0.0f;
    }
    const double log2n = log(2.0 * N);
    const unsigned int inner = m[N - 1];
    const unsigned int mid = (N == 3) ? m[1] : 1;
    const unsigned int plane = mid * inner;  // cells per x0 layer
    for (unsigned int x0 = 1; x0 + 1 < m[0]; x0++) {
        for (unsigned int xm = (N == 3) ? 1 : 0; xm + ((N == 3) ? 1 : 0) < mid; xm++) {
            // first active cell of this pencil (harmonic_cpu.cpp:49-51, :94-101)
            unsigned int offset = (h->currentIteration % 2) != (x0 % 2);
            if (N == 3 && xm % 2 == 0) {
                offset = !offset;
            }
            const unsigned int base = x0 * plane + xm * inner;
            for (unsigned int xi = 1 + offset; xi + 1 < inner; xi += 2) {
                const unsigned int c = base + xi;
                if (locked[c]) {
                    continue;
                }
                const float before = u[c];
                float nb[2 * N];
                nb[0] = u[c - plane];
                nb[1] = u[c + plane];
                if (N == 3) {
                    nb[2] = u[c - inner];
                    nb[3] = u[c + inner];
                }
                nb[2 * N - 2] = u[c - 1];
                nb[2 * N - 1] = u[c + 1];
                float top = std::max(nb[0], nb[1]);
                for (int j = 2; j < 2 * N; j++) {
                    top = std::max(top, nb[j]);
                }
                float acc = expf(nb[0] - top) + expf(nb[1] - top);
                for (int j = 2; j < 2 * N; j++) {
                    acc = acc + expf(nb[j] - top);
                }
                const float shifted = top + logf(acc);
                u[c] = (float)((double)shifted - log2n);
                if (check) {
                    h->delta = std::max(h->delta, (float)fabs(before - u[c]));
                }
            }
        }
    }
}

void sweep(epic::Harmonic *h, bool check)
{
    if (h->n == 2) {
        half_sweep<2>(h, check);
    } else if (h->n == 3) {
        half_sweep<3>(h, check);
    }
}

// (unsigned int)value exactly as the reference's x86-64 build converts it
template <typename F>
inline unsigned int trunc_u(F value)
{
    return (unsigned int)(int64_t)value;
}

// The streamline code exists in a float flavour (on Harmonic) and a double flavour (legacy); the
// arithmetic is the same, only the scalar type and the validity rules differ.
template <typename F>
struct GridView {
    unsigned int w, h;
    const unsigned int *locked;
    const F *u;
};

template <typename F>
int bilinear(const GridView<F> &g, F x, F y, F &out)
{
    const F half = (F)0.5;
    const unsigned int cx = trunc_u(x + half), cy = trunc_u(y + half);
    if (cx >= g.w || cy >= g.h || (g.locked[cy * g.w + cx] == 1 && g.u[cy * g.w + cx] < (F)0.0)) {
        return EPIC_ERROR_INVALID_LOCATION;
    }
    const unsigned int xl = trunc_u(x - half), xr = trunc_u(x + half);
    const unsigned int yt = trunc_u(y - half), yb = trunc_u(y + half);
    if (xl >= g.w || xr >= g.w || yt >= g.h || yb >= g.h) {
        return EPIC_ERROR_INVALID_LOCATION;  // the reference indexes outside the arrays here
    }
    const F alpha = x - xl;
    const F beta = y - yt;
    const F one = ((F)1.0 - alpha) * g.u[yt * g.w + xl] + alpha * g.u[yt * g.w + xr];
    const F two = ((F)1.0 - alpha) * g.u[yb * g.w + xl] + alpha * g.u[yb * g.w + xr];
    out = ((F)1.0 - beta) * one + beta * two;
    return EPIC_SUCCESS;
}

template <typename F>
int unit_gradient(const GridView<F> &g, F x, F y, F cd, F &px, F &py)
{
    F v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    int r = bilinear(g, x - cd, y, v0);
    r += bilinear(g, x + cd, y, v1);
    r += bilinear(g, x, y - cd, v2);
    r += bilinear(g, x, y + cd, v3);
    if (r != EPIC_SUCCESS) {
        return EPIC_ERROR_INVALID_GRADIENT;
    }
    px = (v1 - v0) / ((F)2.0 * cd);
    py = (v3 - v2) / ((F)2.0 * cd);
    const F denom = (F)sqrt((double)px * (double)px + (double)py * (double)py);
    px /= denom;
    py /= denom;
    return EPIC_SUCCESS;
}

template <typename F>
bool revisits_recent_point(const std::vector<F> &p, F step)
{
    const size_t n = p.size();
    if (n % 2 == 1) {
        return true;
    }
    if (n == 0) {
        return false;
    }
    const F x = p[n - 2], y = p[n - 1];
    const size_t floor_i = (n > 12) ? n - 12 : 0;
    for (size_t i = n - 2; i > floor_i; i -= 2) {
        const F dx = x - p[i - 2], dy = y - p[i - 1];
        const F dist = (F)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
        if (dist < step / (F)2.0) {
            return true;
        }
    }
    return false;
}

// direction: +1 climbs the potential (log-space and legacy "flipped"), -1 descends it.
template <typename F>
int trace(const GridView<F> &g, F x, F y, F step, F cd, size_t max_values, int direction, const char *fn,
          unsigned int &k, F *&path)
{
    unsigned int cx = trunc_u(x + (F)0.5), cy = trunc_u(y + (F)0.5);
    std::vector<F> pts;
    pts.push_back(x);
    pts.push_back(y);
    while (g.locked[cy * g.w + cx] != 1 && !revisits_recent_point(pts, step) && pts.size() < max_values) {
        F px = 0, py = 0;
        if (unit_gradient(g, x, y, cd, px, py) != EPIC_SUCCESS) {
            complain(fn, "Could not compute gradient.");
            return EPIC_ERROR_INVALID_GRADIENT;
        }
        if (direction > 0) {
            x += px * step;
            y += py * step;
        } else {
            x -= px * step;
            y -= py * step;
        }
        pts.push_back(x);
        pts.push_back(y);
        cx = trunc_u(x + (F)0.5);
        cy = trunc_u(y + (F)0.5);
        if (cx >= g.w || cy >= g.h) {
            break;  // left the grid: the reference would index outside `locked`
        }
    }
    if (pts.size() / 2 <= 2) {
        complain(fn, "Could not compute a valid path.");
        return EPIC_ERROR_INVALID_PATH;
    }
    k = (unsigned int)(pts.size() / 2);
    path = new F[2 * (size_t)k];
    std::copy(pts.begin(), pts.begin() + 2 * (size_t)k, path);
    return EPIC_SUCCESS;
}

template <typename F>
int legacy_sor(unsigned int w, unsigned int h, F epsilon, F omega, const unsigned int *locked, F *u,
               unsigned int &iter)
{
    F delta = epsilon + (F)1.0;
    iter = 0;
    while (delta >= epsilon || iter < 10000u) {  // MIN_ITERATIONS, harmonic_legacy_cpu.cpp:34
        delta = (F)0.0;
        for (unsigned int y = 1; y + 1 < h; y++) {
            for (unsigned int x = 1; x + 1 < w; x++) {
                const unsigned int c = y * w + x;
                if (locked[c] == 1) {
                    continue;
                }
                const F before = u[c];
                u[c] = ((F)1.0 - omega) * u[c] + omega / (F)4.0 * (u[c - w] + u[c + w] + u[c - 1] + u[c + 1]);
                const F change = u[c] - before;
                const F mag = change < 0 ? -change : change;
                delta = (mag > delta || delta != delta) ? mag : delta;  // fmax
            }
        }
        iter++;
    }
    return EPIC_SUCCESS;
}

}  // namespace

namespace epic {

// ---- harmonic_cpu.cpp:136-220 --------------------------------------------------------------------

int harmonic_update_cpu(Harmonic *harmonic)
{
    sweep(harmonic, false);
    harmonic->currentIteration++;
    return EPIC_SUCCESS;
}

int harmonic_update_and_check_cpu(Harmonic *harmonic)
{
    sweep(harmonic, true);
    harmonic->currentIteration++;
    return (harmonic->delta < harmonic->epsilon) ? EPIC_SUCCESS_AND_CONVERGED : EPIC_SUCCESS;
}

int harmonic_complete_cpu(Harmonic *harmonic)
{
    if (harmonic == nullptr || harmonic->m == nullptr || harmonic->u == nullptr || harmonic->locked == nullptr ||
        harmonic->epsilon <= 0.0 || harmonic->numIterationsToStaggerCheck == 0) {
        complain("harmonic_complete_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    unsigned int longest = 0;
    for (unsigned int i = 0; i < harmonic->n; i++) {
        longest = std::max(longest, harmonic->m[i]);
    }
    harmonic->currentIteration = 0;
    harmonic->delta = harmonic->epsilon + 1.0;
    int result = EPIC_SUCCESS;
    while (result != EPIC_SUCCESS_AND_CONVERGED || harmonic->currentIteration < longest) {
        const bool check = harmonic->currentIteration % harmonic->numIterationsToStaggerCheck == 0;
        result = check ? harmonic_update_and_check_cpu(harmonic) : harmonic_update_cpu(harmonic);
    }
    return EPIC_SUCCESS;
}

// ---- harmonic_utilities_cpu.cpp:38-76 ---------------------------------------------------------------

int harmonic_utilities_set_cells_2d_cpu(Harmonic *harmonic, unsigned int k, unsigned int *v, unsigned int *types)
{
    const char *fn = "harmonic_utilities_set_cells_2d_cpu";
    if (harmonic == nullptr || harmonic->n == 0 || harmonic->m == nullptr || harmonic->u == nullptr ||
        harmonic->locked == nullptr || k == 0 || v == nullptr || types == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    static const float kValue[3] = {(float)EPIC_LOG_SPACE_GOAL, (float)EPIC_LOG_SPACE_OBSTACLE,
                                    (float)EPIC_LOG_SPACE_FREE};
    static const unsigned int kLocked[3] = {1, 1, 0};
    for (unsigned int i = 0; i < k; i++) {
        const unsigned int x = v[2 * i], y = v[2 * i + 1];
        if (y >= harmonic->m[0] || x >= harmonic->m[1]) {
            fprintf(stderr, "Warning[%s]: %s\n", fn, "Provided vector has invalid values outside area.");
            continue;
        }
        if (types[i] > EPIC_CELL_TYPE_FREE) {
            fprintf(stderr, "Warning[%s]: %s\n", fn, "Type is invalid. No change made.");
            continue;
        }
        const size_t c = (size_t)y * harmonic->m[1] + x;
        harmonic->u[c] = kValue[types[i]];
        harmonic->locked[c] = kLocked[types[i]];
    }
    return EPIC_SUCCESS;
}

// ---- extensions: the pose list the callers build from a raw path (src/epic_nav_core_plugin.cpp:310-328,
// src/epic_navigation_node_harmonic.cpp:655-668): world x, world y and yaw of the incoming segment per point.
// Same float operations as the callers' loops (separate multiply and add, atan2 of the float differences
// narrowed to float), so a client receives identical numbers.

int harmonic_path_to_poses_2d(const float *path, unsigned int k, float originX, float originY, float resolution,
                              float *poses)
{
    if (path == nullptr || poses == nullptr || k == 0) {
        return EPIC_ERROR_INVALID_DATA;
    }
    for (unsigned int i = 0; i < k; i++) {
        const float x = path[2 * i + 0], y = path[2 * i + 1];
        // point 0 has no incoming segment: it takes the heading of the first one (the callers put the
        // request's own start pose there)
        const unsigned int a = (i == 0) ? 0 : i - 1, b = (i == 0) ? ((k > 1) ? 1 : 0) : i;
        const float theta = std::atan2(path[2 * b + 1] - path[2 * a + 1], path[2 * b + 0] - path[2 * a + 0]);
        poses[3 * i + 0] = originX + x * resolution;
        poses[3 * i + 1] = originY + y * resolution;
        poses[3 * i + 2] = theta;
    }
    return EPIC_SUCCESS;
}

int harmonic_compute_path_poses_2d_cpu(Harmonic *harmonic, float x, float y, float stepSize, float cdPrecision,
                                       unsigned int maxLength, float originX, float originY, float resolution,
                                       unsigned int &k, float *&poses)
{
    if (poses != nullptr) {
        complain("harmonic_compute_path_poses_2d_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    float *raw = nullptr;
    unsigned int kk = 0;
    const int r = harmonic_compute_path_2d_cpu(harmonic, x, y, stepSize, cdPrecision, maxLength, kk, raw);
    if (r != EPIC_SUCCESS) {
        return r;
    }
    poses = new float[3 * (size_t)kk];
    harmonic_path_to_poses_2d(raw, kk, originX, originY, resolution, poses);
    delete[] raw;
    k = kk;
    return EPIC_SUCCESS;
}

// ---- extensions: dense map ingest (what the node's /map handler and reset-free-cells service compute
// before they call set_cells: src/epic_navigation_node_harmonic.cpp:383-426, :582-611) ---------------------

int harmonic_utilities_set_occupancy_grid_2d_cpu(Harmonic *harmonic, const signed char *data, int obstacleThreshold,
                                                 int noChangeValue)
{
    const char *fn = "harmonic_utilities_set_occupancy_grid_2d_cpu";
    if (harmonic == nullptr || harmonic->n != 2 || harmonic->m == nullptr || harmonic->u == nullptr ||
        harmonic->locked == nullptr || data == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const unsigned int h = harmonic->m[0], w = harmonic->m[1];
    for (unsigned int y = 1; y + 1 < h; y++) {
        for (unsigned int x = 1; x + 1 < w; x++) {
            const size_t c = (size_t)y * w + x;
            const bool goal = harmonic->u[c] == (float)EPIC_LOG_SPACE_GOAL && harmonic->locked[c] == 1;
            if (data[c] == noChangeValue || goal) {
                continue;
            }
            const bool obstacle = data[c] >= obstacleThreshold;
            harmonic->u[c] = obstacle ? (float)EPIC_LOG_SPACE_OBSTACLE : (float)EPIC_LOG_SPACE_FREE;
            harmonic->locked[c] = obstacle ? 1 : 0;
        }
    }
    return EPIC_SUCCESS;
}

int harmonic_utilities_reset_free_cells_2d_cpu(Harmonic *harmonic)
{
    if (harmonic == nullptr || harmonic->n != 2 || harmonic->m == nullptr || harmonic->u == nullptr ||
        harmonic->locked == nullptr) {
        complain("harmonic_utilities_reset_free_cells_2d_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const unsigned int h = harmonic->m[0], w = harmonic->m[1];
    for (unsigned int y = 1; y + 1 < h; y++) {
        for (unsigned int x = 1; x + 1 < w; x++) {
            if (harmonic->locked[(size_t)y * w + x] == 0) {
                harmonic->u[(size_t)y * w + x] = (float)EPIC_LOG_SPACE_FREE;
            }
        }
    }
    return EPIC_SUCCESS;
}

// ---- harmonic_path_cpu.cpp:41-232 -----------------------------------------------------------------------

static bool host_grid(const Harmonic *h, GridView<float> &g)
{
    if (h == nullptr || h->m == nullptr || h->u == nullptr || h->locked == nullptr) {
        return false;
    }
    g.w = h->m[1];
    g.h = h->m[0];
    g.locked = h->locked;
    g.u = h->u;
    return true;
}

int harmonic_compute_potential_2d_cpu(Harmonic *harmonic, float x, float y, float &potential)
{
    GridView<float> g;
    if (!host_grid(harmonic, g)) {
        complain("harmonic_compute_potential_2d_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const int r = bilinear(g, x, y, potential);
    if (r != EPIC_SUCCESS) {
        complain("harmonic_compute_potential_2d_cpu", "Invalid location.");
    }
    return r;
}

int harmonic_compute_gradient_2d_cpu(Harmonic *harmonic, float x, float y, float cdPrecision, float &partialX,
                                     float &partialY)
{
    GridView<float> g;
    if (!host_grid(harmonic, g)) {
        complain("harmonic_compute_gradient_2d_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const int r = unit_gradient(g, x, y, cdPrecision, partialX, partialY);
    if (r != EPIC_SUCCESS) {
        complain("harmonic_compute_gradient_2d_cpu", "Failed to compute potential values.");
    }
    return r;
}

int harmonic_compute_path_2d_cpu(Harmonic *harmonic, float x, float y, float stepSize, float cdPrecision,
                                 unsigned int maxLength, unsigned int &k, float *&path)
{
    const char *fn = "harmonic_compute_path_2d_cpu";
    GridView<float> g;
    if (!host_grid(harmonic, g) || path != nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const unsigned int cx = trunc_u(x + 0.5f), cy = trunc_u(y + 0.5f);
    if (cx >= g.w || cy >= g.h || (g.locked[cy * g.w + cx] == 1 && g.u[cy * g.w + cx] < 0.0f)) {
        complain(fn, "Invalid location.");
        return EPIC_ERROR_INVALID_LOCATION;
    }
    // `2 * maxLength` is a 32-bit unsigned product in the reference (:187)
    return trace(g, x, y, stepSize, cdPrecision, (size_t)(2u * maxLength), +1, fn, k, path);
}

int harmonic_free_path_cpu(float *&path)
{
    delete[] path;
    path = nullptr;
    return EPIC_SUCCESS;
}

// ---- legacy linear-space code: harmonic_legacy_cpu.cpp, harmonic_legacy_path_cpu.cpp ------------------------

int harmonic_legacy_sor_2d_float_cpu(unsigned int w, unsigned int h, float epsilon, float omega,
                                     unsigned int *locked, float *u, unsigned int &iter)
{
    return legacy_sor<float>(w, h, epsilon, omega, locked, u, iter);
}

int harmonic_legacy_sor_2d_double_cpu(unsigned int w, unsigned int h, double epsilon, double omega,
                                      unsigned int *locked, double *u, unsigned int &iter)
{
    return legacy_sor<double>(w, h, epsilon, omega, locked, u, iter);
}

int harmonic_legacy_sor_2d_long_double_cpu(unsigned int w, unsigned int h, long double epsilon, long double omega,
                                           unsigned int *locked, long double *u, unsigned int &iter)
{
    return legacy_sor<long double>(w, h, epsilon, omega, locked, u, iter);
}

static bool legacy_grid(unsigned int w, unsigned int h, unsigned int *locked, double *u, GridView<double> &g)
{
    if (w == 0 || h == 0 || locked == nullptr || u == nullptr) {
        return false;
    }
    g.w = w;
    g.h = h;
    g.locked = locked;
    g.u = u;
    return true;
}

int harmonic_legacy_compute_potential_2d_cpu(unsigned int w, unsigned int h, unsigned int *locked, double *u,
                                             double x, double y, double &potential)
{
    GridView<double> g;
    if (!legacy_grid(w, h, locked, u, g)) {
        complain("harmonic_legacy_compute_potential_2d_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    return bilinear(g, x, y, potential);
}

int harmonic_legacy_compute_gradient_2d_cpu(unsigned int w, unsigned int h, unsigned int *locked, double *u,
                                            double x, double y, double cdPrecision, double &partialX,
                                            double &partialY)
{
    GridView<double> g;
    if (!legacy_grid(w, h, locked, u, g)) {
        complain("harmonic_legacy_compute_gradient_2d_cpu", "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    return unit_gradient(g, x, y, cdPrecision, partialX, partialY);
}

int harmonic_legacy_compute_path_2d_cpu(unsigned int w, unsigned int h, unsigned int *locked, double *u, double x,
                                        double y, double stepSize, double cdPrecision, unsigned int maxLength,
                                        int flipped, unsigned int &k, double *&path)
{
    const char *fn = "harmonic_legacy_compute_path_2d_cpu";
    GridView<double> g;
    if (!legacy_grid(w, h, locked, u, g) || path != nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    const unsigned int cx = trunc_u(x + 0.5), cy = trunc_u(y + 0.5);
    // an obstacle is a locked cell holding 1 (or 0 when the potential is flipped), :162-166
    if (cx >= w || cy >= h ||
        (locked[cy * w + cx] == 1 &&
         ((flipped == 0 && u[cy * w + cx] == 1.0) || (flipped == 1 && u[cy * w + cx] == 0.0)))) {
        complain(fn, "Invalid location.");
        return EPIC_ERROR_INVALID_LOCATION;
    }
    return trace(g, x, y, stepSize, cdPrecision, (size_t)maxLength, flipped == 1 ? +1 : -1, fn, k, path);
}

int harmonic_legacy_free_path_cpu(double *&path)
{
    delete[] path;
    path = nullptr;
    return EPIC_SUCCESS;
}

}  // namespace epic
