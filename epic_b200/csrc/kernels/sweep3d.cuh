// sweep3d.cuh -- temporally blocked red-black log-sum-exp sweep, 3-D (6 neighbours).
//
// The reference has no GPU kernel for n = 3 (harmonic_gpu.cu:334-336, :367-369 are empty branches);
// the semantics are those of the CPU path, harmonic_update_3d_cpu (harmonic_cpu.cpp:81-133):
// iteration `it` updates the interior cells with (it + x0 + x1 + x2) even, locked cells are
// skipped, the six neighbours enter the max and the sum in the order x0-1, x0+1, x1-1, x1+1,
// x2-1, x2+1, and delta is the max |u_prev - u_new| over the cells of the check sweep's colour.
//
// One pass = up to T = 2 consecutive half-sweeps (one of each colour) from the source buffer into the
// destination buffer.  A CTA owns a column of the grid: an (x1, x2) tile of BH x 128 cells (halo: 2
// rows and 4 columns on each side) and walks along x0 through a chunk of layers, keeping a ring of six
// layer-tiles in shared memory.  Each tile arrives by one TMA box load (the field is presented to the
// TMA unit as a 2-D tensor of pitch x (layers * m1), so rows outside the grid come from the
// neighbouring layer or are zero-filled; they only ever feed cells whose results are masked).  At
// step s the CTA
//     sweeps colour A (iteration it0) on layer s+1     -- reads layers s, s+1, s+2,
//     sweeps colour B (iteration it0 + 1) on layer s-1 -- reads layers s-2, s-1, s, all of which had their
//                                                         colour-A sweep in an earlier step,
//     writes layer s-1 (now two iterations ahead) to the destination buffer, and
//     has the load of layer s+3 in flight (and layer s+5 on its way into L2).
// DRAM sees every cell once for reading (plus the halo overlap) and once for writing per TWO
// half-sweeps; the per-launch work is the same arithmetic as the 2-D kernel with two more
// neighbours.  With count == 1 only the colour-A sweep runs and layer s is written.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "math_policies.cuh"
#include "sweep2d.cuh"   // mbarrier / TMA helpers

namespace epic_b200 {

constexpr int k3W = 128;       // tile columns (x2): one float4 per lane
constexpr int k3G = k3W / 4;   // float4 groups per tile row
constexpr int k3Slots = 6;     // layer-tiles resident in shared memory
constexpr int k3HC = 4;        // halo columns on each side (float4 alignment; 2 would do)
constexpr int k3HR = 2;        // halo rows / layers on each side = sweeps per pass
constexpr int k3OutW = k3W - 2 * k3HC;
constexpr int k3Threads = 256;

struct Sweep3DParams {
    float *dst;                // destination buffer, buffer layer 0
    const uint32_t *freemask;  // 1 bit per cell, buffer layout (rows = layer * m1 + x1)
    const uint32_t *ctrl_done;
    uint32_t *delta_bits;
    uint64_t pitch;            // floats per x2 row
    uint64_t layer_floats;     // floats per x0 layer = m1 * pitch
    uint32_t mask_wpr;         // mask words per x2 row
    uint32_t m0, m1, m2;       // global dimensions
    int64_t grow0;             // global x0 of buffer layer 0
    uint32_t buf_layers;       // layers present in the buffer
    uint32_t own_lo, own_hi;   // buffer layers written by this slab
    uint32_t BH;               // tile rows (x1) including the halo
    uint32_t ntx, nty;         // tiles along x2 and x1
    uint32_t zchunk;           // owned layers per CTA
    uint32_t count;            // half-sweeps in this pass: 1 or 2
    uint32_t it0;              // iteration of the first one
    uint32_t check;            // the last sweep of the pass accumulates delta
    // peer-to-peer halos (sharded runs): the first / last `halo_layers` owned layers are also stored
    // into the neighbouring GPU's ghost layers (in ITS destination buffer)
    float *peer_up;            // layer 0 of the upper neighbour's ghost-below region
    float *peer_down;          // layer 0 of the lower neighbour's ghost-above region
    uint32_t halo_layers;
    // static-block skipping (TRACK kernels): one byte per CTA, as in Sweep2DParams; the neighbourhood is the
    // 3 x 3 x 3 block of CTAs around this one
    const uint8_t *chg_prev;
    uint8_t *chg_out;
    uint32_t *skipped;
    uint32_t ntz;              // chunks along x0
};

// [6 layer-tiles][6 nibble tiles][MathTables][6 mbarriers][8 warp maxima]
inline size_t sweep3d_smem_bytes(uint32_t BH)
{
    return (size_t)k3Slots * BH * k3W * sizeof(float) + (size_t)k3Slots * BH * k3G + sizeof(MathTables) +
           k3Slots * 8 + 8 * sizeof(float);
}

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <class Math, bool TRACK>
struct Rows3D {
    const Math &math;
    int lane;
    float dmax;
    bool chg;          // TRACK: some update of this lane changed a value

    // One LDS.128 per lane (conflict-free across the warp).  Written in PTX because the compiler otherwise
    // narrows a float4 load of which only two components are used into two LDS.32 with a 16-byte lane stride
    // -- four-way bank conflicts on each (profiles/r01d_sweep3d_fast_ncu.md).
    static __device__ __forceinline__ float4 ld(const float *plane, int r, int lane)
    {
        float4 v;
        asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(smem_u32(plane + r * k3W + lane * 4)));
        return v;
    }

    // One tile row: `b` = this row (updated in place), a / c = rows r-1 / r+1 of the same layer,
    // pm / pp = the layers below / above.  EVEN: the columns 0, 2 of each float4 are the active colour.
    template <bool EVEN, bool CHECK>
    __device__ __forceinline__ void row(float *pc, const float *pm, const float *pp, const uint8_t *mk, int r,
                                        const float4 &a, float4 &b, const float4 &c, bool chk)
    {
        const uint32_t nib = mk[r * k3G + lane];
        const uint32_t active = EVEN ? (nib & 0x5u) : (nib & 0xAu);
        if (!__any_sync(0xffffffffu, active != 0u)) {
            return;
        }
        const float4 m = ld(pm, r, lane), q = ld(pp, r, lane);
        float4 nw = b;
        if (EVEN) {
            // lane 0's left neighbour is outside the tile: its column 0 is never active
            const float left = __shfl_up_sync(0xffffffffu, b.w, 1);
            const float nx = math.update6(m.x, q.x, a.x, c.x, left, b.y);
            const float nz = math.update6(m.z, q.z, a.z, c.z, b.y, b.w);
            if (active & 1u) nw.x = nx;
            if (active & 4u) nw.z = nz;
            if (TRACK) {
                chg = chg || (__float_as_uint(nw.x) != __float_as_uint(b.x)) || (__float_as_uint(nw.z) != __float_as_uint(b.z));
            }
        } else {
            const float right = __shfl_down_sync(0xffffffffu, b.x, 1);
            const float ny = math.update6(m.y, q.y, a.y, c.y, b.x, b.z);
            const float nq = math.update6(m.w, q.w, a.w, c.w, b.z, right);
            if (active & 2u) nw.y = ny;
            if (active & 8u) nw.w = nq;
            if (TRACK) {
                chg = chg || (__float_as_uint(nw.y) != __float_as_uint(b.y)) || (__float_as_uint(nw.w) != __float_as_uint(b.w));
            }
        }
        if (CHECK && chk) {
            // |prev - new| is 0 for cells that were not updated
            float d = fabsf(__fsub_rn(b.x, nw.x));
            if (d > dmax) dmax = d;
            d = fabsf(__fsub_rn(b.y, nw.y));
            if (d > dmax) dmax = d;
            d = fabsf(__fsub_rn(b.z, nw.z));
            if (d > dmax) dmax = d;
            d = fabsf(__fsub_rn(b.w, nw.w));
            if (d > dmax) dmax = d;
        }
        *reinterpret_cast<float4 *>(pc + r * k3W + lane * 4) = nw;
        b = nw;
    }

    // Exactly n = 2 or 4 rows from r0, fully unrolled with the colour of each row fixed at compile time: with
    // the default tile height every warp's share of a layer is such a group, and the generic loop below
    // (peel, pair loop, tail) costs about as many instructions as a row of arithmetic.
    template <bool CHECK>
    __device__ __forceinline__ void band4(float *pc, const float *pm, const float *pp, const uint8_t *mk, int r0, int n,
                                          uint32_t par, bool chk, int chk_lo, int chk_hi)
    {
        float4 a = ld(pc, r0 - 1, lane), b = ld(pc, r0, lane), c = ld(pc, r0 + 1, lane);
        const bool k0 = chk && r0 >= chk_lo && r0 < chk_hi, k1 = chk && r0 + 1 >= chk_lo && r0 + 1 < chk_hi;
        const bool k2 = chk && r0 + 2 >= chk_lo && r0 + 2 < chk_hi, k3 = chk && r0 + 3 >= chk_lo && r0 + 3 < chk_hi;
        if (((par + (uint32_t)r0) & 1u) == 0u) {
            row<true, CHECK>(pc, pm, pp, mk, r0, a, b, c, k0);
            a = ld(pc, r0 + 2, lane);
            row<false, CHECK>(pc, pm, pp, mk, r0 + 1, b, c, a, k1);
            if (n > 2) {
                b = ld(pc, r0 + 3, lane);
                row<true, CHECK>(pc, pm, pp, mk, r0 + 2, c, a, b, k2);
                c = ld(pc, r0 + 4, lane);
                row<false, CHECK>(pc, pm, pp, mk, r0 + 3, a, b, c, k3);
            }
        } else {
            row<false, CHECK>(pc, pm, pp, mk, r0, a, b, c, k0);
            a = ld(pc, r0 + 2, lane);
            row<true, CHECK>(pc, pm, pp, mk, r0 + 1, b, c, a, k1);
            if (n > 2) {
                b = ld(pc, r0 + 3, lane);
                row<false, CHECK>(pc, pm, pp, mk, r0 + 2, c, a, b, k2);
                c = ld(pc, r0 + 4, lane);
                row<true, CHECK>(pc, pm, pp, mk, r0 + 3, a, b, c, k3);
            }
        }
    }

    // Rows [ra, rb) of one layer.  par = (it + x0 + x1 of tile row 0) & 1: cell (r, c) is active when
    // (par + r + c) is even.  chk_lo / chk_hi: rows whose deltas count (the output rows), for lanes with chk.
    template <bool CHECK>
    __device__ __forceinline__ void band(float *pc, const float *pm, const float *pp, const uint8_t *mk, int ra, int rb,
                                         uint32_t par, bool chk, int chk_lo, int chk_hi)
    {
        if (ra >= rb) {
            return;
        }
        if (rb - ra == 4 || rb - ra == 2) {
            band4<CHECK>(pc, pm, pp, mk, ra, rb - ra, par, chk, chk_lo, chk_hi);
            return;
        }
        float4 a = ld(pc, ra - 1, lane), b = ld(pc, ra, lane), c;
        int r = ra;
        if (((par + (uint32_t)r) & 1u) != 0u) {   // first row has its odd columns active: peel it
            c = ld(pc, r + 1, lane);
            row<false, CHECK>(pc, pm, pp, mk, r, a, b, c, chk && r >= chk_lo && r < chk_hi);
            a = b;
            b = c;
            ++r;
        }
        for (; r + 1 < rb; r += 2) {
            c = ld(pc, r + 1, lane);
            row<true, CHECK>(pc, pm, pp, mk, r, a, b, c, chk && r >= chk_lo && r < chk_hi);
            a = ld(pc, r + 2, lane);
            row<false, CHECK>(pc, pm, pp, mk, r + 1, b, c, a, chk && r + 1 >= chk_lo && r + 1 < chk_hi);
            b = a;
            a = c;
        }
        if (r < rb) {
            c = ld(pc, r + 1, lane);
            row<true, CHECK>(pc, pm, pp, mk, r, a, b, c, chk && r >= chk_lo && r < chk_hi);
        }
    }
};

template <class Math, bool TRACK = false>
__global__ void __launch_bounds__(k3Threads, 2)
sweep3d_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ Sweep3DParams p,
               const __grid_constant__ Math math_in)
{
    if (*p.ctrl_done) {
        return;  // a previous check sweep already met the termination rule
    }
    if (TRACK) {
        // Skip this CTA's column chunk when no update of the previous pass changed a value in any of the 27
        // chunks around it (see Sweep2DParams::chg_prev); chunks that read ghost layers always run.
        const int txy_n = (int)(p.ntx * p.nty);
        const int cz = blockIdx.x / txy_n, cxy = blockIdx.x % txy_n;
        const int cx = cxy % (int)p.ntx, cy = cxy / (int)p.ntx;
        const bool reads_ghost = (cz == 0 && p.grow0 + (int64_t)p.own_lo > 0) ||
                                 (cz == (int)p.ntz - 1 && p.grow0 + (int64_t)p.own_hi < (int64_t)p.m0);
        uint32_t any = reads_ghost ? 1u : 0u;
        for (int dz = -1; dz <= 1; ++dz) {
            for (int dy = -1; dy <= 1; ++dy) {
                for (int dx = -1; dx <= 1; ++dx) {
                    const int x = cx + dx, y = cy + dy, z = cz + dz;
                    if (x >= 0 && x < (int)p.ntx && y >= 0 && y < (int)p.nty && z >= 0 && z < (int)p.ntz) {
                        any |= p.chg_prev[(z * (int)p.nty + y) * (int)p.ntx + x];
                    }
                }
            }
        }
        if (any == 0u) {
            if (threadIdx.x == 0) {
                p.chg_out[blockIdx.x] = 0;
                atomicAdd(p.skipped, 1u);
            }
            return;
        }
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int BH = (int)p.BH;
    const int plane_floats = BH * k3W;
    float *planes = reinterpret_cast<float *>(smem_raw);
    uint8_t *masks = smem_raw + (size_t)k3Slots * plane_floats * sizeof(float);
    MathTables *tables = reinterpret_cast<MathTables *>(masks + (size_t)k3Slots * BH * k3G);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tables + 1);
    float *s_red = reinterpret_cast<float *>(bars + k3Slots);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (see sweep2d.cuh)
    const uint32_t tiles_xy = p.ntx * p.nty;
    const int tz = blockIdx.x / tiles_xy;
    const int txy = blockIdx.x % tiles_xy;
    const int tx = txy % p.ntx, ty = txy / p.ntx;
    const int gx0 = tx * k3OutW - k3HC;                        // x2 of tile column 0
    const int gy0 = ty * (BH - 2 * k3HR) - k3HR;               // x1 of tile row 0
    const int z0 = (int)p.own_lo + tz * (int)p.zchunk;         // first output layer (buffer layer)
    const int z1 = min(z0 + (int)p.zchunk, (int)p.own_hi);
    const int L = z1 - z0 + 2 * k3HR;                          // layers this CTA touches: local j = 0 .. L-1
    const int layer0 = z0 - k3HR;                              // buffer layer of j = 0

    if (tid == 0) {
        for (int i = 0; i < k3Slots; ++i) {
            mbar_init(&bars[i], 1);
        }
    }
    __syncthreads();

    auto issue_load = [&](int j) {   // thread 0 only
        const int slot = j % k3Slots;
        mbar_expect_tx(&bars[slot], (uint32_t)(plane_floats * sizeof(float)));
        tma_load_2d(planes + (size_t)slot * plane_floats, &src_map, gx0, (layer0 + j) * (int)p.m1 + gy0, &bars[slot]);
    };
    // may-update nibbles of layer j: one byte per float4 group, bit i = cell i of the group is free, inside
    // the grid's interior, and has all six neighbours in the tile.  One item (row, 32-column word) per thread;
    // mask_fetch issues the global loads, mask_commit (after the step's arithmetic, so that the load latency
    // is hidden) packs them into shared memory.
    const int w0 = gx0 >> 5;        // floor division, gx0 may be negative
    const int sh = gx0 - (w0 << 5);
    const int mrow = tid >> 2, mword = tid & 3;
    const bool mitem = tid < BH * 4;
    auto mask_fetch = [&](int j, uint32_t &lo, uint32_t &hi) {
        const int layer = layer0 + j;
        const int64_t x0 = p.grow0 + layer;
        const int x1 = gy0 + mrow;
        lo = 0u;
        hi = 0u;
        if (mitem && layer >= 0 && layer < (int)p.buf_layers && x0 > 0 && x0 < (int64_t)p.m0 - 1 && mrow > 0 &&
            mrow < BH - 1 && x1 > 0 && x1 < (int)p.m1 - 1) {
            const uint32_t *row = p.freemask + ((size_t)layer * p.m1 + (size_t)x1) * p.mask_wpr;
            const int wa = w0 + mword, wb = wa + 1;
            lo = (wa >= 0 && wa < (int)p.mask_wpr) ? __ldg(row + wa) : 0u;
            hi = (wb >= 0 && wb < (int)p.mask_wpr) ? __ldg(row + wb) : 0u;
        }
    };
    auto mask_commit = [&](int j, uint32_t lo, uint32_t hi) {
        if (!mitem) {
            return;
        }
        uint32_t bits = __funnelshift_r(lo, hi, sh);
        if (mword == 0) bits &= ~1u;            // tile column 0: no left neighbour in the tile
        if (mword == 3) bits &= ~0x80000000u;   // tile column 127
        const int x_first = gx0 + mword * 32;   // x2 of bit 0
        if (x_first <= 0 && x_first + 31 >= 0) bits &= ~(1u << (0 - x_first));
        const int x_last = (int)p.m2 - 1;       // the global border columns are never updated
        if (x_first <= x_last && x_first + 31 >= x_last) bits &= ~(1u << (x_last - x_first));
        uint2 packed;
        packed.x = (bits & 0xFu) | ((bits & 0xF0u) << 4) | ((bits & 0xF00u) << 8) | ((bits & 0xF000u) << 12);
        bits >>= 16;
        packed.y = (bits & 0xFu) | ((bits & 0xF0u) << 4) | ((bits & 0xF00u) << 8) | ((bits & 0xF000u) << 12);
        *reinterpret_cast<uint2 *>(masks + (size_t)(j % k3Slots) * BH * k3G + mrow * k3G + mword * 8) = packed;
    };

    const int first_loads = min(L, 4);
    if (tid == 0) {
        for (int j = 0; j < first_loads; ++j) {
            issue_load(j);
        }
    }
    if (Math::kUsesTables) {
        load_math_tables(tables, tid, k3Threads);
    }
    for (int j = 0; j < first_loads; ++j) {
        uint32_t lo, hi;
        mask_fetch(j, lo, hi);
        mask_commit(j, lo, hi);
    }
    __syncthreads();

    Math math = math_in;
    math.bind(tables);
    Rows3D<Math, TRACK> rows{math, lane, 0.0f, false};

    const int col = lane * 4;
    const bool lane_out = col >= k3HC && col < k3W - k3HC;      // lanes 1 .. 30
    const int perA = (BH - 2 + 7) / 8, perB = (BH - 2 * k3HR + 7) / 8;
    const int raA = 1 + warp * perA, rbA = min(raA + perA, BH - 1);
    const int raB = k3HR + warp * perB, rbB = min(raB + perB, BH - k3HR);
    const bool two = p.count >= 2;
    // colour parity of tile row 0, column 0 at iteration it0, layer j = 0 (gx0 is even)
    const uint32_t par0 = (p.it0 + (uint32_t)(p.grow0 + layer0) + (uint32_t)gy0) & 1u;

    auto slot_plane = [&](int j) { return planes + (size_t)(j % k3Slots) * plane_floats; };
    auto slot_mask = [&](int j) { return masks + (size_t)(j % k3Slots) * BH * k3G; };
    auto wait_layer = [&](int j) { mbar_wait(&bars[j % k3Slots], (uint32_t)((j / k3Slots) & 1)); };
    auto owned = [&](int j) { return j >= k3HR && j < L - k3HR; };

    // A warp stores whole rows of layer j's output region (lanes 1..30 hold its 120 columns).
    auto write_back = [&](int j) {
        const int layer = layer0 + j;
        const float *pl = slot_plane(j);
        const float *up = nullptr, *dn = nullptr;   // non-null: this layer also goes to that neighbour
        if (p.peer_up != nullptr && layer < (int)(p.own_lo + p.halo_layers)) {
            up = p.peer_up;
        }
        if (p.peer_down != nullptr && layer >= (int)(p.own_hi - p.halo_layers)) {
            dn = p.peer_down;
        }
        const int r_end = min(BH - k3HR, (int)p.m1 - gy0);
        if (!(lane_out && gx0 + col < (int)p.pitch)) {
            return;
        }
        // 32-bit offsets inside the layer (Field::create rejects layers of 2^31 floats or more)
        uint32_t off = (uint32_t)(gy0 + k3HR + warp) * (uint32_t)p.pitch + (uint32_t)(gx0 + col);
        const uint32_t step = (uint32_t)(k3Threads / 32) * (uint32_t)p.pitch;
        const float *src = pl + (k3HR + warp) * k3W + col;
        float *lay = p.dst + (size_t)layer * p.layer_floats;
        float *lay_up = up ? p.peer_up + (size_t)(layer - (int)p.own_lo) * p.layer_floats : nullptr;
        float *lay_dn = dn ? p.peer_down + (size_t)(layer - (int)(p.own_hi - p.halo_layers)) * p.layer_floats : nullptr;
        for (int r = k3HR + warp; r < r_end; r += k3Threads / 32) {
            const float4 v = *reinterpret_cast<const float4 *>(src);
            *reinterpret_cast<float4 *>(lay + off) = v;
            if (lay_up != nullptr) {
                *reinterpret_cast<float4 *>(lay_up + off) = v;
            }
            if (lay_dn != nullptr) {
                *reinterpret_cast<float4 *>(lay_dn + off) = v;
            }
            off += step;
            src += (k3Threads / 32) * k3W;
        }
    };

    for (int s = 0; s <= L - 2; ++s) {
        uint32_t mlo = 0u, mhi = 0u;
        // every thread is past the barrier that ended step s-1: the slot of layer s-3 is free
        if (s >= 1 && s + 3 < L) {
            if (tid == 0) {
                fence_proxy_async_smem();
                issue_load(s + 3);
                if (s + 5 < L) {   // warm L2 two layers further ahead
                    tma_prefetch_2d(&src_map, gx0, (layer0 + s + 5) * (int)p.m1 + gy0);
                }
            }
            mask_fetch(s + 3, mlo, mhi);
        }
        if (s == 0) {
            wait_layer(0);
            wait_layer(1);
        }
        if (s + 2 < L) {
            wait_layer(s + 2);
        }
        if (s + 1 <= L - 2) {
            // colour A (iteration it0) on layer s+1
            const int j = s + 1;
            const uint32_t par = (par0 + (uint32_t)j) & 1u;
            if (p.check && !two) {
                rows.template band<true>(slot_plane(j), slot_plane(j - 1), slot_plane(j + 1), slot_mask(j), raA, rbA, par,
                                         lane_out && owned(j), k3HR, BH - k3HR);
            } else {
                rows.template band<false>(slot_plane(j), slot_plane(j - 1), slot_plane(j + 1), slot_mask(j), raA, rbA, par,
                                          false, 0, 0);
            }
        }
        if (two && s >= 3) {
            // colour B (iteration it0 + 1) on layer s-1
            const int j = s - 1;
            const uint32_t par = (par0 + 1u + (uint32_t)j) & 1u;
            if (p.check) {
                rows.template band<true>(slot_plane(j), slot_plane(j - 1), slot_plane(j + 1), slot_mask(j), raB, rbB, par,
                                         lane_out && owned(j), k3HR, BH - k3HR);
            } else {
                rows.template band<false>(slot_plane(j), slot_plane(j - 1), slot_plane(j + 1), slot_mask(j), raB, rbB, par,
                                          false, 0, 0);
            }
        }
        if (s >= 1 && s + 3 < L) {
            mask_commit(s + 3, mlo, mhi);   // its slot (layer s-3's) has been idle since step s-2
        }
        __syncthreads();
        if (two) {
            if (s >= 3) {
                write_back(s - 1);   // layers 2 .. L-3: exactly the owned ones
            }
        } else if (s + 1 <= L - 2 && owned(s + 1)) {
            write_back(s + 1);
        }
    }

    if (TRACK) {
        const int any = __syncthreads_or(rows.chg ? 1 : 0);
        if (tid == 0) {
            p.chg_out[blockIdx.x] = any ? 1 : 0;
        }
    }
    if (p.check) {
        float dmax = rows.dmax;
        for (int o = 16; o > 0; o >>= 1) {
            dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        }
        if (lane == 0) {
            s_red[warp] = dmax;
        }
        __syncthreads();
        if (tid == 0) {
            float m = 0.0f;
            for (int i = 0; i < k3Threads / 32; ++i) {
                m = fmaxf(m, s_red[i]);
            }
            if (m > 0.0f) {
                atomicMax(p.delta_bits, __float_as_uint(m));
            }
        }
    }
}

}  // namespace epic_b200
