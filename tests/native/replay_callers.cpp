// ROS-free replay of the two callers of libepic's harmonic path, written against the library's public
// headers only, the ones the reference's callers include (`<epic/harmonic/*.h>` resolves to include/epic/
// here and, in the build container, to the reference's own libepic/include/epic/ -- the same source compiles
// against both).
//
//   plan  the nav_core plugin: EpicNavCorePlugin::initialize + makePlan
//         (reference src/epic_nav_core_plugin.cpp:108-136, :234-338): cells from a costmap, border locked, one
//         goal cell, harmonic_complete_gpu (CPU twin on failure, as the plugin does), then
//         harmonic_compute_path_2d_cpu on the host copy and the pose / yaw post-processing.
//   node  the anytime node: initAlg, an OccupancyGrid message turned into one set-cells call over every
//         interior cell, add-goals, ticks of update_and_check + (steps-1) updates, get-cell, compute-path,
//         set-cells (new obstacles), reset-free-cells, more ticks, uninitAlg
//         (reference src/epic_navigation_node_harmonic.cpp:165-204, :206-282, :357-426, :438-674).
//
// Every stage prints `name value` lines (hashes of the arrays a ROS client would receive), so that the
// output of a GPU run can be compared with a CPU run of the same binary and with the same binary linked
// against the reference's CPU sources (-DREPLAY_CPU_ONLY, oracle/_ref).
//
// -DREPLAY_DENSE_INGEST (this library's headers only) replaces the node's two k = N set-cells lists by the dense
// map-ingest extensions; the output must not change.
//
// usage: replay_callers plan map.bin cpu|gpu goal_x goal_y start_wx start_wy
//        replay_callers node map.bin cpu|gpu goal_wx goal_wy start_wx start_wy ticks steps
//   map.bin = uint32 height, uint32 width, height*width bytes (plan: costmap costs 0..255;
//             node: occupancy values as int8: -2 no change, -1 unknown, 0..100)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <epic/constants.h>
#include <epic/error_codes.h>
#include <epic/harmonic/harmonic.h>
#include <epic/harmonic/harmonic_cpu.h>
#include <epic/harmonic/harmonic_gpu.h>
#include <epic/harmonic/harmonic_model_gpu.h>
#include <epic/harmonic/harmonic_path_cpu.h>
#include <epic/harmonic/harmonic_utilities_cpu.h>
#include <epic/harmonic/harmonic_utilities_gpu.h>

using namespace epic;

namespace {

const float kOriginX = -12.5f, kOriginY = 3.25f, kResolution = 0.05f;   // a map_server style origin / resolution

uint64_t fnv(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull)
{
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < bytes; ++i) {
        h = (h ^ p[i]) * 1099511628211ull;
    }
    return h;
}

void say(const char *name, uint64_t v) { printf("%s %016llx\n", name, (unsigned long long)v); }
void sayu(const char *name, unsigned v) { printf("%s %u\n", name, v); }
void sayf(const char *name, float v)
{
    uint32_t b;
    memcpy(&b, &v, 4);
    printf("%s %08x\n", name, b);
}

bool load_map(const char *path, unsigned &h, unsigned &w, std::vector<unsigned char> &cells)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    uint32_t hw[2];
    bool ok = fread(hw, 4, 2, f) == 2;
    if (ok) {
        h = hw[0];
        w = hw[1];
        cells.resize((size_t)h * w);
        ok = fread(cells.data(), 1, cells.size(), f) == cells.size();
    }
    fclose(f);
    return ok;
}

void lock_border(Harmonic &hm)
{
    const unsigned h = hm.m[0], w = hm.m[1];
    for (unsigned y = 0; y < h; ++y) {
        for (unsigned x = 0; x < w; x += (y == 0 || y == h - 1) ? 1 : w - 1) {
            hm.u[y * w + x] = EPIC_LOG_SPACE_OBSTACLE;
            hm.locked[y * w + x] = 1;
        }
    }
}

// the pose list a client receives: world x, world y, yaw per point after the first
uint64_t poses_hash(const float *raw, unsigned k)
{
    uint64_t hsh = 1469598103934665603ull;
    for (unsigned i = 1; i < k; ++i) {
        const float x = raw[2 * i], y = raw[2 * i + 1];
        const float yaw = atan2f(y - raw[2 * (i - 1) + 1], x - raw[2 * (i - 1)]);
        const float wx = kOriginX + x * kResolution, wy = kOriginY + y * kResolution;
        const float pose[3] = {wx, wy, yaw};
        hsh = fnv(pose, sizeof(pose), hsh);
    }
    return hsh;
}

bool world_to_map(float wx, float wy, unsigned w, unsigned h, float &mx, float &my)
{
    if (wx < kOriginX || wy < kOriginY || wx >= kOriginX + w * kResolution || wy >= kOriginY + h * kResolution) {
        return false;
    }
    mx = (wx - kOriginX) / kResolution;
    my = (wy - kOriginY) / kResolution;
    return true;
}

int report_path(Harmonic &hm, float x, float y, float step, float cd, unsigned max_length, const char *tag)
{
    unsigned k = 0;
    float *raw = nullptr;
    const int r = harmonic_compute_path_2d_cpu(&hm, x, y, step, cd, max_length, k, raw);
    printf("%s_result %d\n", tag, r);
    if (r == EPIC_SUCCESS) {
        printf("%s_points %u\n", tag, k);
        printf("%s_raw %016llx\n", tag, (unsigned long long)fnv(raw, (size_t)k * 2 * sizeof(float)));
        printf("%s_poses %016llx\n", tag, (unsigned long long)poses_hash(raw, k));
    }
    if (raw != nullptr) {
        delete[] raw;   // the callers release paths with delete[], not harmonic_free_path_cpu
    }
    return r;
}

int plan(const char *map, bool gpu, unsigned gx, unsigned gy, float swx, float swy)
{
    unsigned h = 0, w = 0;
    std::vector<unsigned char> cost;
    if (!load_map(map, h, w, cost)) return 2;
    Harmonic hm;
    memset(&hm, 0, sizeof(hm));
    hm.n = 2;
    hm.m = new unsigned int[2];
    hm.m[0] = h;
    hm.m[1] = w;
    hm.u = new float[(size_t)w * h];
    hm.locked = new unsigned int[(size_t)w * h];
    hm.epsilon = 1e-3f;
    hm.numIterationsToStaggerCheck = 100;
    for (unsigned y = 1; y + 1 < h; ++y) {
        for (unsigned x = 1; x + 1 < w; ++x) {
            const bool obstacle = cost[(size_t)y * w + x] >= 250;
            hm.u[y * w + x] = obstacle ? EPIC_LOG_SPACE_OBSTACLE : EPIC_LOG_SPACE_FREE;
            hm.locked[y * w + x] = obstacle ? 1 : 0;
        }
    }
    lock_border(hm);
    for (int rep = 0; rep < 2; ++rep) {   // two plans on the same object, as move_base would ask for
        // setGoal: earlier goals become free cells again
        for (unsigned y = 1; y + 1 < h; ++y) {
            for (unsigned x = 1; x + 1 < w; ++x) {
                if (hm.u[y * w + x] == EPIC_LOG_SPACE_GOAL) {
                    hm.u[y * w + x] = EPIC_LOG_SPACE_FREE;
                    hm.locked[y * w + x] = 0;
                }
            }
        }
        const unsigned tx = rep == 0 ? gx : gx + 3, ty = rep == 0 ? gy : gy + 2;
        hm.u[ty * w + tx] = EPIC_LOG_SPACE_GOAL;
        hm.locked[ty * w + tx] = 1;
        // a re-plan starts from the previous solution in the free cells, exactly like the plugin (it never resets u)
        int result = EPIC_ERROR_INVALID_DATA;
#ifndef REPLAY_CPU_ONLY
        if (gpu) {
            result = harmonic_complete_gpu(&hm, 1024);
            printf("complete_gpu_result %d\n", result);
        }
#endif
        if (result != EPIC_SUCCESS) {
            result = harmonic_complete_cpu(&hm);
        }
        printf("plan%d_result %d\n", rep, result);
        if (result != EPIC_SUCCESS) return 3;
        sayu(rep == 0 ? "plan0_iterations" : "plan1_iterations", hm.currentIteration);
        sayf(rep == 0 ? "plan0_delta" : "plan1_delta", hm.delta);
        say(rep == 0 ? "plan0_u" : "plan1_u", fnv(hm.u, (size_t)w * h * sizeof(float)));
        float x = 0.0f, y = 0.0f;
        world_to_map(swx, swy, w, h, x, y);
        const unsigned max_length = (unsigned)(h * w / 0.05f);
        report_path(hm, x, y, 0.05f, 0.5f, max_length, rep == 0 ? "plan0_path" : "plan1_path");
    }
    delete[] hm.u;
    delete[] hm.locked;
    delete[] hm.m;
    return 0;
}

struct Node {
    Harmonic hm;
    bool gpu = false;
    unsigned w = 0, h = 0;

    bool set_cells(std::vector<unsigned> &v, std::vector<unsigned> &types)
    {
        if (types.empty()) return false;
        if (harmonic_utilities_set_cells_2d_cpu(&hm, (unsigned)types.size(), &v[0], &types[0]) != EPIC_SUCCESS) return false;
#ifndef REPLAY_CPU_ONLY
        if (gpu && harmonic_utilities_set_cells_2d_gpu(&hm, 1024, (unsigned)types.size(), &v[0], &types[0]) != EPIC_SUCCESS) {
            return false;
        }
#endif
        return true;
    }

    void tick(unsigned steps)
    {
        int r;
#ifndef REPLAY_CPU_ONLY
        if (gpu) {
            r = harmonic_update_and_check_gpu(&hm, 1024);
            if (r == EPIC_SUCCESS) {
                for (unsigned i = 0; i + 1 < steps; ++i) {
                    if (harmonic_update_gpu(&hm, 1024) != EPIC_SUCCESS) return;
                }
            }
            return;
        }
#endif
        r = harmonic_update_and_check_cpu(&hm);
        if (r == EPIC_SUCCESS) {
            for (unsigned i = 0; i + 1 < steps; ++i) {
                if (harmonic_update_cpu(&hm) != EPIC_SUCCESS) return;
            }
        }
    }

    bool fetch()
    {
#ifndef REPLAY_CPU_ONLY
        if (gpu) return harmonic_get_potential_values_gpu(&hm) == EPIC_SUCCESS;
#endif
        return true;
    }
};

int node(const char *map, bool want_gpu, float gwx, float gwy, float swx, float swy, unsigned ticks, unsigned steps)
{
    unsigned h = 0, w = 0;
    std::vector<unsigned char> grid;
    if (!load_map(map, h, w, grid)) return 2;
    Node nd;
    Harmonic &hm = nd.hm;
    memset(&hm, 0, sizeof(hm));
    nd.w = w;
    nd.h = h;
    // initAlg
    hm.n = 2;
    hm.m = new unsigned int[2];
    hm.m[0] = h;
    hm.m[1] = w;
    hm.u = new float[(size_t)w * h];
    hm.locked = new unsigned int[(size_t)w * h];
    hm.epsilon = 1e-3f;
    hm.numIterationsToStaggerCheck = 100;
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        hm.u[i] = 0.0f;
        hm.locked[i] = 0;
    }
    lock_border(hm);
#ifndef REPLAY_CPU_ONLY
    if (want_gpu) {
        int r = harmonic_initialize_dimension_size_gpu(&hm);
        r += harmonic_initialize_potential_values_gpu(&hm);
        r += harmonic_initialize_locked_gpu(&hm);
        r += harmonic_initialize_gpu(&hm, 1024);
        nd.gpu = (r == EPIC_SUCCESS);
        printf("gpu_initialised %d\n", nd.gpu ? 1 : 0);
    }
#endif
    // the /map message: every interior cell that is not "no change" and not already a goal
    std::vector<unsigned> v, types;
    for (unsigned y = 1; y + 1 < h; ++y) {
        for (unsigned x = 1; x + 1 < w; ++x) {
            const signed char occ = (signed char)grid[(size_t)y * w + x];
            const bool is_goal = hm.u[y * w + x] == EPIC_LOG_SPACE_GOAL && hm.locked[y * w + x] == 1;
            if (occ == -2 || is_goal) continue;
            v.push_back(x);
            v.push_back(y);
            types.push_back(occ >= 50 ? EPIC_CELL_TYPE_OBSTACLE : EPIC_CELL_TYPE_FREE);
        }
    }
#ifdef REPLAY_DENSE_INGEST
    // the optional upgrade of INTEGRATION.md section 3: hand the message itself to the library
    {
        bool ok = harmonic_utilities_set_occupancy_grid_2d_cpu(&hm, (const signed char *)grid.data(), 50, -2) == EPIC_SUCCESS;
        if (ok && nd.gpu) {
            ok = harmonic_utilities_set_occupancy_grid_2d_gpu(&hm, (const signed char *)grid.data(), 50, -2) == EPIC_SUCCESS;
        }
        printf("map_cells %zu\nmap_set %d\n", types.size(), ok ? 1 : 0);
    }
#else
    printf("map_cells %zu\nmap_set %d\n", types.size(), nd.set_cells(v, types) ? 1 : 0);
#endif
    // add one goal (world coordinates), unless it falls into an obstacle
    float gx = 0.0f, gy = 0.0f;
    world_to_map(gwx, gwy, w, h, gx, gy);
    {
        const unsigned cx = (unsigned)(gx + 0.5f), cy = (unsigned)(gy + 0.5f);
        const bool obstacle = cx >= w || cy >= h || (hm.u[cy * w + cx] == EPIC_LOG_SPACE_OBSTACLE && hm.locked[cy * w + cx] == 1);
        v.clear();
        types.clear();
        if (!obstacle) {
            v.push_back((unsigned)gx);
            v.push_back((unsigned)gy);
            types.push_back(EPIC_CELL_TYPE_GOAL);
        }
        printf("goal_added %d\n", (!obstacle && nd.set_cells(v, types)) ? 1 : 0);
    }
    for (unsigned t = 0; t < ticks; ++t) nd.tick(steps);
    // get-cell + compute-path services
    float sx = 0.0f, sy = 0.0f;
    world_to_map(swx, swy, w, h, sx, sy);
    printf("fetch1 %d\n", nd.fetch() ? 1 : 0);
    sayu("iterations1", hm.currentIteration);
    sayf("cell1", hm.u[(unsigned)sy * w + (unsigned)sx]);
    say("u1", fnv(hm.u, (size_t)w * h * sizeof(float)));
    report_path(hm, sx, sy, 0.05f, 0.5f, 2000000u, "path1");
    // set-cells service: a wall of new obstacles across the middle rows, then keep relaxing
    v.clear();
    types.clear();
    for (unsigned x = w / 4; x < w / 2; ++x) {
        v.push_back(x);
        v.push_back(h / 2);
        types.push_back(EPIC_CELL_TYPE_OBSTACLE);
    }
    v.push_back(w + 7);   // out of range: skipped by the library
    v.push_back(1);
    types.push_back(EPIC_CELL_TYPE_OBSTACLE);
    printf("wall_set %d\n", nd.set_cells(v, types) ? 1 : 0);
    for (unsigned t = 0; t < ticks / 2; ++t) nd.tick(steps);
    printf("fetch2 %d\n", nd.fetch() ? 1 : 0);
    say("u2", fnv(hm.u, (size_t)w * h * sizeof(float)));
    say("locked2", fnv(hm.locked, (size_t)w * h * sizeof(unsigned)));
    // reset-free-cells service: every unlocked interior cell back to the free value
    v.clear();
    types.clear();
    for (unsigned y = 1; y + 1 < h; ++y) {
        for (unsigned x = 1; x + 1 < w; ++x) {
            if (hm.locked[y * w + x] == 0) {
                v.push_back(x);
                v.push_back(y);
                types.push_back(EPIC_CELL_TYPE_FREE);
            }
        }
    }
#ifdef REPLAY_DENSE_INGEST
    {
        bool ok = harmonic_utilities_reset_free_cells_2d_cpu(&hm) == EPIC_SUCCESS;
        if (ok && nd.gpu) {
            ok = harmonic_utilities_reset_free_cells_2d_gpu(&hm) == EPIC_SUCCESS;
        }
        printf("reset_cells %zu\nreset_set %d\n", types.size(), ok ? 1 : 0);
    }
#else
    printf("reset_cells %zu\nreset_set %d\n", types.size(), nd.set_cells(v, types) ? 1 : 0);
#endif
    for (unsigned t = 0; t < ticks; ++t) nd.tick(steps);
    printf("fetch3 %d\n", nd.fetch() ? 1 : 0);
    sayu("iterations3", hm.currentIteration);
    say("u3", fnv(hm.u, (size_t)w * h * sizeof(float)));
    report_path(hm, sx, sy, 0.2f, 0.4f, 1000000u, "path3");
    // uninitAlg: host arrays first, device handles after, as the node does
    delete[] hm.u;
    hm.u = nullptr;
    delete[] hm.locked;
    hm.locked = nullptr;
    delete[] hm.m;
    hm.m = nullptr;
    hm.n = 0;
#ifndef REPLAY_CPU_ONLY
    if (nd.gpu) {
        int r = harmonic_uninitialize_dimension_size_gpu(&hm);
        r += harmonic_uninitialize_potential_values_gpu(&hm);
        r += harmonic_uninitialize_locked_gpu(&hm);
        r += harmonic_uninitialize_gpu(&hm);
        printf("gpu_uninitialised %d\n", r == EPIC_SUCCESS ? 1 : 0);
    }
#endif
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s plan|node map.bin cpu|gpu a b c d [ticks steps]\n", argv[0]);
        return 1;
    }
    const bool gpu = strcmp(argv[3], "gpu") == 0;
    if (strcmp(argv[1], "plan") == 0) {
        return plan(argv[2], gpu, (unsigned)atoi(argv[4]), (unsigned)atoi(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]));
    }
    if (argc < 10) return 1;
    return node(argv[2], gpu, (float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]),
                (unsigned)atoi(argv[8]), (unsigned)atoi(argv[9]));
}
