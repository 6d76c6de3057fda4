// sweep3d.cuh -- red-black log-sum-exp half-sweep, 3-D (6 neighbours).
//
// The reference has no GPU kernel for n = 3 (harmonic_gpu.cu:334-336, :367-369 are empty branches);
// the semantics are those of the CPU path, harmonic_update_3d_cpu (harmonic_cpu.cpp:81-133):
// iteration `it` updates the interior cells with (it + x0 + x1 + x2) even, locked cells are
// skipped, the six neighbours enter the max and the sum in the order x0-1, x0+1, x1-1, x1+1,
// x2-1, x2+1, and delta is the max |u_prev - u_new| over the cells of the check sweep's colour.
//
// One half-sweep per launch, in place (a red-black half-sweep only reads the other colour, so
// there is no hazard).  A warp owns 128 consecutive x2 cells of one (x0, x1) pencil, one float4 per
// lane; the x2 neighbours that fall into the adjacent lane come by shuffle.  A CTA covers 8
// adjacent x1 rows so that the x1-1 / x1+1 rows are L1 hits, and consecutive CTAs walk x1 then x0
// so that the x0-1 / x0+1 planes are L2 hits: DRAM sees each cell once per sweep for reading and
// once for writing.
#pragma once

#include <stdint.h>

#include "math_policies.cuh"

namespace epic_b200 {

struct Sweep3DParams {
    float *u;                  // current buffer, buffer layer 0
    const uint32_t *freemask;  // 1 bit per cell, buffer layout
    const uint32_t *ctrl_done;
    uint32_t *delta_bits;
    uint64_t pitch;            // floats per x2 row
    uint64_t layer_floats;     // floats per x0 layer = m1 * pitch
    uint32_t mask_wpr;         // mask words per x2 row
    uint32_t m0, m1, m2;       // global dimensions
    int64_t grow0;             // global x0 of buffer layer 0
    uint32_t own_lo, own_hi;   // buffer layers updated by this slab
    uint32_t segs;             // 128-cell segments per x2 row
    uint32_t row_blocks;       // ceil(m1 / 8)
    uint32_t it;               // iteration (colour)
    uint32_t check;
    // peer-to-peer halos (sharded runs): new values of the first / last owned layer are also stored
    // into the neighbouring GPU's ghost layer
    float *peer_up;            // the upper neighbour's ghost-below layer
    float *peer_down;          // the lower neighbour's ghost-above layer
};

template <class Math>
__global__ void __launch_bounds__(256, 3)
sweep3d_kernel(const Sweep3DParams p, const Math math_in)
{
    if (*p.ctrl_done) {
        return;
    }
    __shared__ MathTables tables;
    __shared__ float s_red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_math_tables(&tables, tid, 256);
    __syncthreads();
    Math math = math_in;
    math.bind(&tables);

    float dmax = 0.0f;
    const uint64_t per_layer = (uint64_t)p.row_blocks * p.segs;
    const uint64_t total = (uint64_t)(p.own_hi - p.own_lo) * per_layer;
    for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const uint32_t b0 = p.own_lo + (uint32_t)(w / per_layer);
        const uint32_t rem = (uint32_t)(w % per_layer);
        const uint32_t x1 = (rem / p.segs) * 8u + (uint32_t)warp;
        const uint32_t x2 = (rem % p.segs) * 128u + (uint32_t)lane * 4u;
        const int64_t x0 = p.grow0 + (int64_t)b0;
        if (x0 <= 0 || x0 >= (int64_t)p.m0 - 1 || x1 == 0 || x1 >= p.m1 - 1) {
            continue;  // warp-uniform: a whole pencil on the global border
        }
        // free bits of this lane's four cells, border columns removed (lanes past the row: none)
        const uint64_t row = (uint64_t)b0 * p.m1 + x1;
        uint32_t nib = 0u;
        if (x2 < p.pitch) {
            nib = (__ldg(p.freemask + row * p.mask_wpr + (x2 >> 5)) >> (x2 & 31u)) & 0xFu;
            if (x2 == 0) {
                nib &= ~1u;
            }
            if (x2 + 4 > p.m2 - 1) {  // some of x2..x2+3 are >= m2-1
                const uint32_t keep = (p.m2 - 1 > x2) ? (p.m2 - 1 - x2) : 0u;  // cells below m2-1
                nib &= (1u << keep) - 1u;
            }
        }
        // active colour: (it + x0 + x1 + x2) even
        const bool even_cols = (((uint32_t)(p.it + (uint32_t)x0 + x1)) & 1u) == 0u;
        const uint32_t active = even_cols ? (nib & 0x5u) : (nib & 0xAu);
        if (!__any_sync(0xffffffffu, active != 0u)) {
            continue;
        }
        const bool in_row = x2 < p.pitch;
        float *c = p.u + (uint64_t)b0 * p.layer_floats + (uint64_t)x1 * p.pitch + (in_row ? x2 : 0u);
        const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float4 cur = in_row ? *reinterpret_cast<const float4 *>(c) : zero4;
        const float4 a0 = in_row ? *reinterpret_cast<const float4 *>(c - p.layer_floats) : zero4;
        const float4 a1 = in_row ? *reinterpret_cast<const float4 *>(c + p.layer_floats) : zero4;
        const float4 b0v = in_row ? *reinterpret_cast<const float4 *>(c - p.pitch) : zero4;
        const float4 b1v = in_row ? *reinterpret_cast<const float4 *>(c + p.pitch) : zero4;
        float4 nw = cur;
        if (even_cols) {
            float left = __shfl_up_sync(0xffffffffu, cur.w, 1);
            if (lane == 0 && x2 > 0) {
                left = c[-1];
            }
            const float nx = math.update6(a0.x, a1.x, b0v.x, b1v.x, left, cur.y);
            const float nz = math.update6(a0.z, a1.z, b0v.z, b1v.z, cur.y, cur.w);
            if (active & 1u) nw.x = nx;
            if (active & 4u) nw.z = nz;
        } else {
            float right = __shfl_down_sync(0xffffffffu, cur.x, 1);
            if (lane == 31 && x2 + 4 < p.pitch) {
                right = c[4];
            }
            const float ny = math.update6(a0.y, a1.y, b0v.y, b1v.y, cur.x, cur.z);
            const float nq = math.update6(a0.w, a1.w, b0v.w, b1v.w, cur.z, right);
            if (active & 2u) nw.y = ny;
            if (active & 8u) nw.w = nq;
        }
        if (active != 0u) {
            if (p.check) {
                float d = fabsf(__fsub_rn(cur.x, nw.x));
                if (d > dmax) dmax = d;
                d = fabsf(__fsub_rn(cur.y, nw.y));
                if (d > dmax) dmax = d;
                d = fabsf(__fsub_rn(cur.z, nw.z));
                if (d > dmax) dmax = d;
                d = fabsf(__fsub_rn(cur.w, nw.w));
                if (d > dmax) dmax = d;
            }
            // Only the active colour's words are stored: the other colour's words in this float4 are
            // being read by neighbouring warps and must not be rewritten with possibly stale copies
            // (they are unchanged here, but a partial-word store keeps the sweep formally race-free).
            if (active & 1u) c[0] = nw.x;
            if (active & 2u) c[1] = nw.y;
            if (active & 4u) c[2] = nw.z;
            if (active & 8u) c[3] = nw.w;
            float *q = nullptr;
            if (p.peer_up != nullptr && b0 == p.own_lo) {
                q = p.peer_up + (uint64_t)x1 * p.pitch + x2;
            } else if (p.peer_down != nullptr && b0 + 1 == p.own_hi) {
                q = p.peer_down + (uint64_t)x1 * p.pitch + x2;
            }
            if (q != nullptr) {
                if (active & 1u) q[0] = nw.x;
                if (active & 2u) q[1] = nw.y;
                if (active & 4u) q[2] = nw.z;
                if (active & 8u) q[3] = nw.w;
            }
            if (p.peer_up != nullptr && p.peer_down != nullptr && b0 == p.own_lo && b0 + 1 == p.own_hi) {
                q = p.peer_down + (uint64_t)x1 * p.pitch + x2;   // a one-layer slab feeds both neighbours
                if (active & 1u) q[0] = nw.x;
                if (active & 2u) q[1] = nw.y;
                if (active & 4u) q[2] = nw.z;
                if (active & 8u) q[3] = nw.w;
            }
        }
    }

    if (p.check) {
        for (int o = 16; o > 0; o >>= 1) {
            dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        }
        if (lane == 0) {
            s_red[warp] = dmax;
        }
        __syncthreads();
        if (tid == 0) {
            float m = 0.0f;
            for (int i = 0; i < 8; ++i) {
                m = fmaxf(m, s_red[i]);
            }
            if (m > 0.0f) {
                atomicMax(p.delta_bits, __float_as_uint(m));
            }
        }
    }
}

}  // namespace epic_b200
