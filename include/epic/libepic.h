/*
 * libepic.h -- the C ABI of the B200-native libepic.so, in one header.
 *
 * This is the drop-in boundary: the same `Harmonic` struct and the same 30 unmangled
 * entry points that the reference's ROS nodes and Python ctypes wrapper bind.  Each
 * declaration cites the reference header (relative to /root/reference/libepic/include/epic)
 * it replaces.  The per-file headers the reference's callers include
 * (epic/harmonic/harmonic_gpu.h, ...) are kept as one-line forwards to this file, so
 * src/epic_nav_core_plugin.cpp and src/epic_navigation_node_harmonic.cpp compile unchanged.
 *
 * References (`float &`, `unsigned int &`, `float *&`) are pointers at the ABI level; the
 * Python wrapper passes them with ctypes.byref (python/epic/epic_harmonic.py:61-124).
 */
#ifndef EPIC_B200_LIBEPIC_H
#define EPIC_B200_LIBEPIC_H

/* ---- error_codes.h:31-46 ---- */
#define EPIC_SUCCESS                        0
#define EPIC_SUCCESS_AND_CONVERGED          1
#define EPIC_ERROR_INVALID_DATA             2
#define EPIC_ERROR_INVALID_CUDA_PARAM       3
#define EPIC_ERROR_DEVICE_MALLOC            4
#define EPIC_ERROR_MEMCPY_TO_DEVICE         5
#define EPIC_ERROR_MEMCPY_TO_HOST           6
#define EPIC_ERROR_DEVICE_FREE              7
#define EPIC_ERROR_KERNEL_EXECUTION         8
#define EPIC_ERROR_DEVICE_SYNCHRONIZE       9
#define EPIC_ERROR_INVALID_LOCATION         10
#define EPIC_ERROR_INVALID_CELL_TYPE        11
#define EPIC_ERROR_INVALID_GRADIENT         12
#define EPIC_ERROR_INVALID_PATH             13

/* ---- constants.h:32-43 ---- */
#define EPIC_FLT_MAX                1e+300
#define EPIC_FLT_MIN                (-EPIC_FLT_MAX)
#define EPIC_CELL_TYPE_GOAL         0
#define EPIC_CELL_TYPE_OBSTACLE     1
#define EPIC_CELL_TYPE_FREE         2
#define EPIC_LOG_SPACE_GOAL         0.0
#define EPIC_LOG_SPACE_OBSTACLE     -1e6
#define EPIC_LOG_SPACE_FREE         -1e6

#ifdef __cplusplus
namespace epic {
#define EPIC_REF(T) T &
#define EPIC_API extern "C"
#else
#define EPIC_REF(T) T *
#define EPIC_API
#endif

/*
 * harmonic/harmonic.h:44-64.  80 bytes on x86-64; field order and types are frozen.
 * n: number of dimensions (2 or 3); m[n]: size of each dimension, last one fastest;
 * u: log-potentials; locked: 0 = free, 1 = locked (goal when u == 0, obstacle when u < 0);
 * the border cells are expected to be locked.
 * d_m, d_u, d_locked, d_delta: opaque to callers.  In this library they are handles to the
 * device-resident field (all of them refer to one object holding the padded ping-pong buffers,
 * the 1-bit free mask, the stream and the control block); they are null exactly when the
 * reference's pointers would be null, and the initialize / uninitialize calls set and clear them
 * with the reference's rules.
 */
typedef struct Harmonic {
    unsigned int n;
    unsigned int *m;
    float *u;
    unsigned int *locked;
    float epsilon;
    float delta;
    unsigned int numIterationsToStaggerCheck;
    unsigned int currentIteration;
    unsigned int *d_m;
    float *d_u;
    unsigned int *d_locked;
    float *d_delta;
} Harmonic;

/* ---- harmonic/harmonic_cpu.h:40-56 : the reference CPU solver, kept for callers that ask for it ---- */
EPIC_API int harmonic_complete_cpu(Harmonic *harmonic);
EPIC_API int harmonic_update_cpu(Harmonic *harmonic);
EPIC_API int harmonic_update_and_check_cpu(Harmonic *harmonic);

/* ---- harmonic/harmonic_gpu.h:39-86 : numThreads must be a multiple of 32, otherwise it is a hint ---- */
EPIC_API int harmonic_complete_gpu(Harmonic *harmonic, unsigned int numThreads);
EPIC_API int harmonic_initialize_gpu(Harmonic *harmonic, unsigned int numThreads);
EPIC_API int harmonic_execute_gpu(Harmonic *harmonic, unsigned int numThreads);
EPIC_API int harmonic_uninitialize_gpu(Harmonic *harmonic);
EPIC_API int harmonic_update_gpu(Harmonic *harmonic, unsigned int numThreads);
EPIC_API int harmonic_update_and_check_gpu(Harmonic *harmonic, unsigned int numThreads);
EPIC_API int harmonic_get_potential_values_gpu(Harmonic *harmonic);

/* ---- harmonic/harmonic_model_gpu.h:38-80 ---- */
EPIC_API int harmonic_initialize_dimension_size_gpu(Harmonic *harmonic);
EPIC_API int harmonic_uninitialize_dimension_size_gpu(Harmonic *harmonic);
EPIC_API int harmonic_initialize_potential_values_gpu(Harmonic *harmonic);
EPIC_API int harmonic_uninitialize_potential_values_gpu(Harmonic *harmonic);
EPIC_API int harmonic_initialize_locked_gpu(Harmonic *harmonic);
EPIC_API int harmonic_uninitialize_locked_gpu(Harmonic *harmonic);
EPIC_API int harmonic_update_model_gpu(Harmonic *harmonic);

/* ---- harmonic/harmonic_utilities_cpu.h:41, harmonic_utilities_gpu.h:42 ----
 * v = [x0, y0, x1, y1, ...] (x = column, y = row), types[i] in EPIC_CELL_TYPE_*. */
EPIC_API int harmonic_utilities_set_cells_2d_cpu(Harmonic *harmonic, unsigned int k, unsigned int *v,
                                                 unsigned int *types);
EPIC_API int harmonic_utilities_set_cells_2d_gpu(Harmonic *harmonic, unsigned int numThreads, unsigned int k,
                                                 unsigned int *v, unsigned int *types);

/* ---- harmonic/harmonic_path_cpu.h:42-82 : streamlines on the HOST copy of u ---- */
EPIC_API int harmonic_compute_potential_2d_cpu(Harmonic *harmonic, float x, float y, EPIC_REF(float) potential);
EPIC_API int harmonic_compute_gradient_2d_cpu(Harmonic *harmonic, float x, float y, float cdPrecision,
                                              EPIC_REF(float) partialX, EPIC_REF(float) partialY);
EPIC_API int harmonic_compute_path_2d_cpu(Harmonic *harmonic, float x, float y, float stepSize,
                                          float cdPrecision, unsigned int maxLength, EPIC_REF(unsigned int) k,
                                          EPIC_REF(float *) path);
EPIC_API int harmonic_free_path_cpu(EPIC_REF(float *) path);

/* ---- harmonic/harmonic_legacy_cpu.h:44-76, harmonic_legacy_path_cpu.h:43-90 : linear-space SOR (host only) ---- */
EPIC_API int harmonic_legacy_sor_2d_float_cpu(unsigned int w, unsigned int h, float epsilon, float omega,
                                              unsigned int *locked, float *u, EPIC_REF(unsigned int) iter);
EPIC_API int harmonic_legacy_sor_2d_double_cpu(unsigned int w, unsigned int h, double epsilon, double omega,
                                               unsigned int *locked, double *u, EPIC_REF(unsigned int) iter);
EPIC_API int harmonic_legacy_sor_2d_long_double_cpu(unsigned int w, unsigned int h, long double epsilon,
                                                    long double omega, unsigned int *locked, long double *u,
                                                    EPIC_REF(unsigned int) iter);
EPIC_API int harmonic_legacy_compute_potential_2d_cpu(unsigned int w, unsigned int h, unsigned int *locked,
                                                      double *u, double x, double y, EPIC_REF(double) potential);
EPIC_API int harmonic_legacy_compute_gradient_2d_cpu(unsigned int w, unsigned int h, unsigned int *locked,
                                                     double *u, double x, double y, double cdPrecision,
                                                     EPIC_REF(double) partialX, EPIC_REF(double) partialY);
EPIC_API int harmonic_legacy_compute_path_2d_cpu(unsigned int w, unsigned int h, unsigned int *locked, double *u,
                                                 double x, double y, double stepSize, double cdPrecision,
                                                 unsigned int maxLength, int flipped, EPIC_REF(unsigned int) k,
                                                 EPIC_REF(double *) path);
EPIC_API int harmonic_legacy_free_path_cpu(EPIC_REF(double *) path);

/* ---- Extension: the legacy linear-space SOR on the GPU (the reference has it on the host only,
 * harmonic_legacy_cpu.cpp:36-141).  Same arguments and results as the *_cpu twins above, bit for bit: the lexicographic
 * in-place sweep is executed as waves of independent cells (epic_b200/csrc/abi/legacy_gpu.cu).  long double is x87
 * arithmetic and stays host-only. ---- */
EPIC_API int harmonic_legacy_sor_2d_float_gpu(unsigned int w, unsigned int h, float epsilon, float omega,
                                              unsigned int *locked, float *u, EPIC_REF(unsigned int) iter);
EPIC_API int harmonic_legacy_sor_2d_double_gpu(unsigned int w, unsigned int h, double epsilon, double omega,
                                               unsigned int *locked, double *u, EPIC_REF(unsigned int) iter);

/* ---- Extensions of this library (not in the reference): streamlines on the DEVICE-resident field,
 * so that a path or a cell query does not need harmonic_get_potential_values_gpu's full-field copy
 * (the reference's anytime node does one per service call, src/epic_navigation_node_harmonic.cpp:531,622).
 * Same arithmetic as the *_cpu functions above, bit for bit.  They need the d_* handles to be live. ---- */
EPIC_API int harmonic_compute_potential_2d_gpu(Harmonic *harmonic, float x, float y, EPIC_REF(float) potential);
EPIC_API int harmonic_compute_gradient_2d_gpu(Harmonic *harmonic, float x, float y, float cdPrecision,
                                              EPIC_REF(float) partialX, EPIC_REF(float) partialY);
EPIC_API int harmonic_compute_path_2d_gpu(Harmonic *harmonic, float x, float y, float stepSize,
                                          float cdPrecision, unsigned int maxLength, EPIC_REF(unsigned int) k,
                                          EPIC_REF(float *) path);
/* numPaths starts [x0, y0, x1, y1, ...]; results[i], k[i], paths[i] (each new float[2*k[i]] or null). */
EPIC_API int harmonic_compute_paths_2d_gpu(Harmonic *harmonic, unsigned int numPaths, const float *starts,
                                           float stepSize, float cdPrecision, unsigned int maxLength, int *results,
                                           unsigned int *k, float **paths);

/* The pose list the callers build from a raw path (src/epic_nav_core_plugin.cpp:310-328,
 * src/epic_navigation_node_harmonic.cpp:655-668): poses = [world x, world y, yaw] * k with world = origin + cell *
 * resolution and yaw = atan2 of the incoming segment (point 0: of the first segment), in the callers' float
 * arithmetic.  harmonic_compute_path_poses_2d_* trace the path (host field / device-resident field) and return
 * new float[3*k], to be released with delete[] or harmonic_free_path_cpu. */
EPIC_API int harmonic_path_to_poses_2d(const float *path, unsigned int k, float originX, float originY,
                                       float resolution, float *poses);
EPIC_API int harmonic_compute_path_poses_2d_cpu(Harmonic *harmonic, float x, float y, float stepSize,
                                                float cdPrecision, unsigned int maxLength, float originX,
                                                float originY, float resolution, EPIC_REF(unsigned int) k,
                                                EPIC_REF(float *) poses);
EPIC_API int harmonic_compute_path_poses_2d_gpu(Harmonic *harmonic, float x, float y, float stepSize,
                                                float cdPrecision, unsigned int maxLength, float originX,
                                                float originY, float resolution, EPIC_REF(unsigned int) k,
                                                EPIC_REF(float *) poses);

/* ---- Extensions: dense map ingest.  The reference's node turns every /map message into a set_cells call
 * over EVERY interior cell and implements "reset free cells" the same way (k = N scatter lists, 12 bytes per
 * cell; src/epic_navigation_node_harmonic.cpp:383-426, :582-611).  These take the occupancy grid itself (one
 * signed byte per cell, row-major like nav_msgs/OccupancyGrid.data) and classify it in place: value ==
 * noChangeValue or a goal cell -> untouched; value >= obstacleThreshold -> obstacle; else free.  As with
 * set_cells, a caller that mirrors the field on the host calls the _cpu twin too. ---- */
EPIC_API int harmonic_utilities_set_occupancy_grid_2d_cpu(Harmonic *harmonic, const signed char *data,
                                                          int obstacleThreshold, int noChangeValue);
EPIC_API int harmonic_utilities_set_occupancy_grid_2d_gpu(Harmonic *harmonic, const signed char *data,
                                                          int obstacleThreshold, int noChangeValue);
EPIC_API int harmonic_utilities_reset_free_cells_2d_cpu(Harmonic *harmonic);
EPIC_API int harmonic_utilities_reset_free_cells_2d_gpu(Harmonic *harmonic);

#ifdef __cplusplus
}  /* namespace epic */
#endif

#endif /* EPIC_B200_LIBEPIC_H */
