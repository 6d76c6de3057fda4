// legacy_gpu.cu -- the reference's legacy linear-space SOR (libepic/src/harmonic/harmonic_legacy_cpu.cpp:36-141) on
// the GPU, bit for bit.  Extension entry points harmonic_legacy_sor_2d_{float,double}_gpu
// with the signatures of their *_cpu twins; the reference has no GPU version of this solver.
//
// The reference sweeps the grid lexicographically and IN PLACE (Gauss-Seidel with over-relaxation): cell (y, x) of
// iteration k reads (y-1, x) and (y, x-1) of iteration k and (y+1, x), (y, x+1) of iteration k-1.  That order looks
// sequential, but every dependency of a cell points to the wave before its own if waves are numbered
//
//     wave(y, x, k) = (x - 1) + (y - 1) + 2 k:
//
// (y-1, x, k) and (y, x-1, k) sit on wave - 1, and so do (y+1, x, k-1) and (y, x+1, k-1).  All cells of a wave are
// therefore independent, a wave holds every second anti-diagonal -- half the grid, exactly a red-black half-sweep
// whose cells carry different iteration numbers -- and updating the array in place wave after wave reproduces the
// lexicographic result exactly: a value written on wave w is read by its four consumers on wave w + 1 and
// overwritten on wave w + 2.  The arithmetic is the reference's expression with separate multiplies and adds (this
// file is compiled with -fmad=false), so float and double results are bit-identical to the CPU code.  (long double
// is x87 80-bit arithmetic, which a GPU does not have: that variant stays host-only.)
//
// Termination.  The reference stops after the first iteration k >= 9999 whose max |change| is below epsilon.  With
// iterations pipelined through the waves, later iterations have already begun when iteration k completes, so the
// solve runs twice: a discovery run on a scratch copy that records max |change| per iteration (atomicMax on the
// non-negative float bits) until the stopping iteration K-1 is known, then the real run from the caller's field
// with the iteration window clamped to [0, K).  Both runs are fully parallel; together they cost two solves.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../../include/epic/libepic.h"

namespace {

template <typename T> struct Bits;
template <> struct Bits<float> {
    typedef unsigned int type;
    static __device__ __forceinline__ type of(float v) { return __float_as_uint(v); }
    static float from(type b)
    {
        float v;
        memcpy(&v, &b, 4);
        return v;
    }
};
template <> struct Bits<double> {
    typedef unsigned long long type;
    static __device__ __forceinline__ type of(double v) { return (unsigned long long)__double_as_longlong(v); }
    static double from(type b)
    {
        double v;
        memcpy(&v, &b, 8);
        return v;
    }
};

// One wave: the interior cells with (x - 1) + (y - 1) = wave - 2k for some iteration k in [k_lo, k_hi).
// One thread per cell of the active parity.  delta[k - delta_base] collects max |change| of iteration k.
template <typename T>
__global__ void legacy_sor_wave_kernel(T *__restrict__ u, const unsigned int *__restrict__ locked, unsigned int w,
                                       unsigned int h, T one_minus_omega, T omega_quarter, uint64_t wave, uint64_t k_lo,
                                       uint64_t k_hi, typename Bits<T>::type *__restrict__ delta)
{
    const unsigned int half = (w - 2u + 1u) / 2u;     // cells of one parity per interior row (rounded up)
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)half * (h - 2u)) {
        return;
    }
    const unsigned int yi = (unsigned int)(i / half);            // y - 1
    const unsigned int j = (unsigned int)(i % half);
    const unsigned int xi = 2u * j + (unsigned int)((wave + yi) & 1u);   // x - 1, with (xi + yi) = wave (mod 2)
    if (xi >= w - 2u) {
        return;
    }
    const uint64_t d = (uint64_t)xi + yi;
    if (d > wave) {
        return;
    }
    const uint64_t k = (wave - d) >> 1;
    if (k < k_lo || k >= k_hi) {
        return;
    }
    const size_t c = (size_t)(yi + 1u) * w + (xi + 1u);
    if (locked[c] == 1u) {
        return;
    }
    const T before = u[c];
    // (1 - omega) * u + omega / 4 * (((up + down) + left) + right): harmonic_legacy_cpu.cpp:52-56
    T sum = u[c - w] + u[c + w];
    sum = sum + u[c - 1];
    sum = sum + u[c + 1];
    const T after = one_minus_omega * before + omega_quarter * sum;
    u[c] = after;
    T change = after - before;
    change = change < (T)0 ? -change : change;
    if (change > (T)0 && delta != nullptr) {
        atomicMax(delta + (k - k_lo), Bits<T>::of(change));
    }
}

void complain(const char *fn, const char *text)
{
    fprintf(stderr, "Error[%s]: %s\n", fn, text);
}

template <typename T>
int legacy_sor_gpu(const char *fn, unsigned int w, unsigned int h, T epsilon, T omega, const unsigned int *locked, T *u,
                   unsigned int &iter)
{
    typedef typename Bits<T>::type B;
    const unsigned int kMinIterations = 10000u;     // MIN_ITERATIONS, harmonic_legacy_cpu.cpp:34
    if (w == 0 || h == 0 || locked == nullptr || u == nullptr) {
        complain(fn, "Invalid data.");
        return EPIC_ERROR_INVALID_DATA;
    }
    iter = 0;
    if (w < 3 || h < 3) {
        iter = kMinIterations;      // no interior: every iteration sees delta = 0; the loop runs its minimum
        return EPIC_SUCCESS;
    }
    const size_t cells = (size_t)w * h;
    const uint64_t D = (uint64_t)(w - 3u) + (h - 3u);            // largest (x - 1) + (y - 1)
    const unsigned int half = (w - 2u + 1u) / 2u;
    const uint64_t threads = (uint64_t)half * (h - 2u);
    const unsigned int blocks = (unsigned int)((threads + 255) / 256);
    const T one_minus_omega = (T)1.0 - omega, omega_quarter = omega / (T)4.0;

    T *d_u = nullptr;
    unsigned int *d_locked = nullptr;
    B *d_delta = nullptr;
    // Iterations per window of the discovery run.  A window is a full barrier between two iterations (legal: the
    // lexicographic order finishes iteration k before it starts k + 1), so each window pays the pipeline's fill and
    // drain (D waves) once; its deltas are checked when it is complete, i.e. the discovery run overshoots the
    // stopping iteration by less than one window.
    const uint64_t kWindow = 4096;
    cudaStream_t stream = nullptr;
    int result = EPIC_SUCCESS;
    if (cudaMalloc(&d_u, cells * sizeof(T)) != cudaSuccess || cudaMalloc(&d_locked, cells * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&d_delta, kWindow * sizeof(B)) != cudaSuccess || cudaStreamCreate(&stream) != cudaSuccess) {
        cudaGetLastError();
        complain(fn, "Failed to allocate device-side memory.");
        result = EPIC_ERROR_DEVICE_MALLOC;
    }
    if (result == EPIC_SUCCESS &&
        (cudaMemcpyAsync(d_locked, locked, cells * sizeof(unsigned int), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
         cudaMemcpyAsync(d_u, u, cells * sizeof(T), cudaMemcpyHostToDevice, stream) != cudaSuccess)) {
        cudaGetLastError();
        complain(fn, "Failed to copy memory from host to device.");
        result = EPIC_ERROR_MEMCPY_TO_DEVICE;
    }

    // ---- discovery run: which iteration stops the loop? ----
    // Iterations complete in order; iteration k is complete after wave D + 2k.  Deltas live in a window of kWindow
    // iterations starting at `base`; the window is read back when its last iteration is complete.
    uint64_t K = 0;     // the loop's final iteration count
    std::vector<B> host(kWindow);
    for (uint64_t base = 0; result == EPIC_SUCCESS && K == 0; base += kWindow) {
        if (base + kWindow > 0xffffffffull) {
            complain(fn, "No convergence within the 32-bit iteration counter.");
            result = EPIC_ERROR_INVALID_DATA;
            break;
        }
        if (cudaMemsetAsync(d_delta, 0, kWindow * sizeof(B), stream) != cudaSuccess) {
            result = EPIC_ERROR_KERNEL_EXECUTION;
            break;
        }
        // iteration `base` starts on wave 2 * base, the window's last iteration ends on wave D + 2 * (base + kWindow - 1);
        // the kernel keeps to the window's own iterations (the first D of these waves also carried the previous
        // window's last iterations, which are complete)
        const uint64_t first = 2 * base, last = D + 2 * (base + kWindow - 1);
        for (uint64_t wave = first; wave <= last; ++wave) {
            legacy_sor_wave_kernel<T><<<blocks, 256, 0, stream>>>(d_u, d_locked, w, h, one_minus_omega, omega_quarter, wave,
                                                               base, base + kWindow, d_delta);
        }
        if (cudaGetLastError() != cudaSuccess ||
            cudaMemcpyAsync(host.data(), d_delta, kWindow * sizeof(B), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess) {
            cudaGetLastError();
            complain(fn, "Failed to execute the SOR wave kernels.");
            result = EPIC_ERROR_KERNEL_EXECUTION;
            break;
        }
        for (uint64_t i = 0; i < kWindow; ++i) {
            const uint64_t k = base + i;
            const T delta = Bits<T>::from(host[i]);
            // while (delta >= epsilon || iter < MIN): after iteration k, iter = k + 1
            if (!(delta >= epsilon) && k + 1 >= kMinIterations) {
                K = k + 1;
                break;
            }
        }
    }

    // ---- the real run: exactly K iterations from the caller's field ----
    if (result == EPIC_SUCCESS) {
        if (cudaMemcpyAsync(d_u, u, cells * sizeof(T), cudaMemcpyHostToDevice, stream) != cudaSuccess) {
            cudaGetLastError();
            result = EPIC_ERROR_MEMCPY_TO_DEVICE;
        }
        const uint64_t last = D + 2 * (K - 1);
        for (uint64_t wave = 0; result == EPIC_SUCCESS && wave <= last; ++wave) {
            legacy_sor_wave_kernel<T><<<blocks, 256, 0, stream>>>(d_u, d_locked, w, h, one_minus_omega, omega_quarter, wave, 0, K,
                                                               nullptr);
        }
        if (result == EPIC_SUCCESS &&
            (cudaGetLastError() != cudaSuccess ||
             cudaMemcpyAsync(u, d_u, cells * sizeof(T), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
             cudaStreamSynchronize(stream) != cudaSuccess)) {
            cudaGetLastError();
            complain(fn, "Failed to execute the SOR wave kernels.");
            result = EPIC_ERROR_KERNEL_EXECUTION;
        }
        if (result == EPIC_SUCCESS) {
            iter = (unsigned int)K;
        }
    }
    if (stream) cudaStreamDestroy(stream);
    if (d_u) cudaFree(d_u);
    if (d_locked) cudaFree(d_locked);
    if (d_delta) cudaFree(d_delta);
    return result;
}

}  // namespace

namespace epic {

int harmonic_legacy_sor_2d_float_gpu(unsigned int w, unsigned int h, float epsilon, float omega, unsigned int *locked,
                                     float *u, unsigned int &iter)
{
    return legacy_sor_gpu<float>("harmonic_legacy_sor_2d_float_gpu", w, h, epsilon, omega, locked, u, iter);
}

int harmonic_legacy_sor_2d_double_gpu(unsigned int w, unsigned int h, double epsilon, double omega, unsigned int *locked,
                                      double *u, unsigned int &iter)
{
    return legacy_sor_gpu<double>("harmonic_legacy_sor_2d_double_gpu", w, h, epsilon, omega, locked, u, iter);
}

}  // namespace epic
