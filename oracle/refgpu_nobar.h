/* TEST / BENCH INFRASTRUCTURE ONLY -- force-included (nvcc -include) when oracle/Makefile builds the
 * `refgpu_nobar` variant of the UNTOUCHED reference sources: compiles every __syncthreads() out.  See the Makefile
 * for why (the reference's divergent barriers never return on Volta and later). */
#include <cuda_runtime.h>
#define __syncthreads() ((void)0)
