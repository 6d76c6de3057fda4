// sweep2d.cuh -- temporally blocked red-black log-sum-exp sweep, 2-D (4 neighbours).
//
// Replaces the reference kernels harmonic_update_2d_gpu / harmonic_update_and_check_2d_gpu /
// harmonic_compute_max_delta_gpu (libepic/src/harmonic/harmonic_gpu.cu:39-153) with one kernel
// that follows the *CPU* path's semantics (harmonic_cpu.cpp:38-78): colour phase "iteration `it`
// updates interior cells with (it + x0 + x1) odd", locked cells skipped, delta = max |u_prev -
// u_new| over the cells of the check sweep's colour.
//
// One CTA owns one tile: TH x 256 cells of shared memory, filled by a single TMA box load
// (cp.async.bulk.tensor.2d, out-of-range cells zero-filled, so edge tiles need no branches).
// It then performs up to T half-sweeps in place in shared memory.  Each half-sweep of a
// red-black ordering only reads the other colour, so a sweep is race-free within the tile; data
// that would have come from neighbouring tiles goes stale one ring per sweep, which is why the
// tile carries a halo of T rows / HC = 4*ceil(T/4) columns and only the inner (TH-2T) x (256-2HC)
// cells are written back (overlapped tiling).  Rows further than the remaining sweep count from
// the output region are skipped (trapezoid), so the redundant work is ~T/2 rows per side.
//
// Thread mapping: a warp owns a panel of 128 columns (one float4 per lane) and marches down a
// band of rows keeping rows r-1, r, r+1 in registers: per row and lane one LDS.128, one SHFL
// (the one horizontal neighbour that lives in the adjacent lane), two updates, one STS.128.
// The per-cell "may update" bits live in shared memory as one nibble per float4 group.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "math_policies.cuh"

namespace epic_b200 {

constexpr int kTileW = 256;            // columns of the shared-memory tile = one TMA box row
constexpr int kGroups = kTileW / 4;    // float4 groups per tile row

struct Sweep2DParams {
    float *dst;                // destination ping-pong buffer (buffer-row 0)
    const uint32_t *freemask;  // 1 bit per cell, buffer layout
    const uint32_t *ctrl_done; // device flag: a previous check already converged -> no-op
    uint32_t *delta_bits;      // atomicMax target for check sweeps
    uint64_t pitch;            // floats per buffer row
    uint32_t mask_wpr;         // mask words per buffer row
    uint32_t m0, m1;           // global rows / columns of the grid
    int32_t grow0;             // global row of buffer row 0 (negative for the first slab's ghost rows)
    uint32_t buf_rows;         // rows present in the buffer
    uint32_t own_lo, own_hi;   // buffer rows [own_lo, own_hi) are written by this slab
    uint32_t TH, T, HC;        // tile rows, halo rows, halo columns
    uint32_t out_h, out_w;     // TH - 2T, 256 - 2HC
    uint32_t ntx;              // tiles per row of tiles
    uint32_t count;            // half-sweeps in this pass (<= T)
    uint32_t parity0;          // (it0 + grow0) & 1
    uint32_t check;            // last sweep of the pass accumulates delta
    uint32_t prefetch_stride;  // CTAs resident at once: tile (blockIdx + stride) is prefetched to L2
    // Sharded runs with peer-to-peer halos: the first / last `halo_rows` owned rows are ALSO stored into
    // the neighbouring GPU's ghost rows (NVLink peer stores fused into this kernel's write-back).
    float *peer_up;            // row 0 of the upper neighbour's ghost-below region in ITS destination buffer
    float *peer_down;          // row 0 of the lower neighbour's ghost-above region
    uint32_t halo_rows;
    // Static-tile skipping (TRACK kernels, single-slab solves): one byte per tile, "some update of the
    // previous pass changed a value in this tile".  A tile whose 3 x 3 neighbourhood is all-zero replays a
    // computation that changed nothing, on the same input, into a buffer that already holds that data: it
    // returns at once.  Bit-identical by construction; the delta of a skipped tile is 0 because no sweep of
    // the previous pass moved any of its cells.
    // In-kernel ordering of passes between neighbouring GPUs (p2p_sync != 0).  The tile rows that touch a
    // neighbour (they read its ghost rows and store into its ghost rows) run FIRST; each of them waits until
    // that neighbour's flag word says its edge tiles finished the previous pass, and the last edge tile of a
    // side to finish publishes this pass's index in the neighbour's flag word (system-scope release after
    // every thread fenced its peer stores).  Interior tiles never wait, so the latency of the exchange hides
    // behind them, and in the steady state the edge tiles find their flag already set.
    uint32_t p2p_sync;
    uint32_t pass_index;             // 1-based index of this pass on this slab
    const uint32_t *wait_up;         // this slab's flag words: written by the upper / lower neighbour
    const uint32_t *wait_dn;
    uint32_t *signal_up;             // the neighbours' flag words for this slab
    uint32_t *signal_dn;
    uint32_t *edge_count;            // two local counters: edge tiles of this pass finished, per side
    uint32_t n_edge_up, n_edge_dn;   // edge tiles per side
    const uint8_t *chg_prev;   // flags written by the previous pass
    uint8_t *chg_out;          // flags of this pass
    uint32_t *skipped;         // statistics: tiles skipped (device counter)
    uint32_t nty;              // rows of tiles
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int x, int y)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y)
                 : "memory");
}

// Shared-memory carve-up (dynamic):
//   [tile TH*256 floats][free nibbles TH*64 bytes][MathTables][mbarrier 8 B][warp maxima NT/32 floats]
inline size_t sweep2d_smem_bytes(uint32_t TH, uint32_t NT)
{
    return (size_t)TH * kTileW * sizeof(float) + (size_t)TH * kGroups + sizeof(MathTables) + 8 + (NT / 32) * sizeof(float);
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Tile row of launch-order row r: with p2p_sync the first and the last two tile rows come first.
__device__ __forceinline__ int tile_row_of(int r, int nty, bool edge_first)
{
    if (!edge_first || nty <= 3) {
        return r;
    }
    return r == 0 ? 0 : (r == 1 ? nty - 1 : (r == 2 ? nty - 2 : r - 2));
}

// One tile row for one lane: update the active colour of `cur` (in place) from the row above (`up`,
// already holding this sweep's values of the other colour) and the row below (`dn`), store it.
template <class Math, bool TRACK>
struct RowCtx {
    const Math &math;
    float *tile;
    const uint8_t *lockt;
    int col, grp, lane, by0;
    bool checking;     // this sweep accumulates delta and this lane's columns are in the output region
    int T, TH, own_lo, own_hi;
    float dmax;
    bool chg;          // TRACK: some update of this lane changed a value

    template <bool EVEN_COLS, bool CHECK>
    __device__ __forceinline__ void row(int r, const float4 &up, float4 &cur, const float4 &dn)
    {
        const uint32_t nib = lockt[r * kGroups + grp];
        const uint32_t active = EVEN_COLS ? (nib & 0x5u) : (nib & 0xAu);
        if (!__any_sync(0xffffffffu, active != 0)) {
            return;
        }
        float4 nw = cur;
        if (EVEN_COLS) {
            float left = __shfl_up_sync(0xffffffffu, cur.w, 1);
            if (lane == 0 && col > 0) {
                left = tile[r * kTileW + col - 1];
            }
            const float nx = math.update4(up.x, dn.x, left, cur.y);
            const float nz = math.update4(up.z, dn.z, cur.y, cur.w);
            if (active & 1u) nw.x = nx;
            if (active & 4u) nw.z = nz;
            if (TRACK) {
                chg = chg || (__float_as_uint(nw.x) != __float_as_uint(cur.x)) ||
                      (__float_as_uint(nw.z) != __float_as_uint(cur.z));
            }
        } else {
            float right = __shfl_down_sync(0xffffffffu, cur.x, 1);
            if (lane == 31 && col + 4 < kTileW) {
                right = tile[r * kTileW + col + 4];
            }
            const float ny = math.update4(up.y, dn.y, cur.x, cur.z);
            const float nq = math.update4(up.w, dn.w, cur.z, right);
            if (active & 2u) nw.y = ny;
            if (active & 8u) nw.w = nq;
            if (TRACK) {
                chg = chg || (__float_as_uint(nw.y) != __float_as_uint(cur.y)) ||
                      (__float_as_uint(nw.w) != __float_as_uint(cur.w));
            }
        }
        if (CHECK && checking) {
            const int b = by0 + r;
            if (r >= T && r < TH - T && b >= own_lo && b < own_hi) {
                // |prev - new| is 0 for cells that were not updated
                float d = fabsf(__fsub_rn(cur.x, nw.x));
                if (d > dmax) dmax = d;
                d = fabsf(__fsub_rn(cur.y, nw.y));
                if (d > dmax) dmax = d;
                d = fabsf(__fsub_rn(cur.z, nw.z));
                if (d > dmax) dmax = d;
                d = fabsf(__fsub_rn(cur.w, nw.w));
                if (d > dmax) dmax = d;
            }
        }
        *reinterpret_cast<float4 *>(tile + r * kTileW + col) = nw;
        cur = nw;
    }

    __device__ __forceinline__ float4 load(int r) const
    {
        return *reinterpret_cast<const float4 *>(tile + r * kTileW + col);
    }

    // Rows [ra, rb) of one sweep.  Cell (r, c) is active when (pb + r + c) is odd; rows alternate between
    // "even columns active" and "odd columns active", so the loop body is a pair of rows with the colour
    // fixed at compile time.  a / b / c rotate roles (row above / row / row below) by renaming instead of
    // by moves.  The check sweep (1 in 100) uses the compact loop only.
    template <bool CHECK>
    __device__ __forceinline__ void band(int ra, int rb, int pb)
    {
        float4 a = load(ra - 1), b = load(ra), c;
        int r = ra;
        if (((r + pb) & 1) == 0) {  // first row updates the odd columns: peel it
            c = load(r + 1);
            row<false, CHECK>(r, a, b, c);
            a = b;
            b = c;
            ++r;
        }
        if (!CHECK) {
            for (; r + 5 < rb; r += 6) {
                c = load(r + 1);
                row<true, CHECK>(r, a, b, c);
                a = load(r + 2);
                row<false, CHECK>(r + 1, b, c, a);
                b = load(r + 3);
                row<true, CHECK>(r + 2, c, a, b);
                c = load(r + 4);
                row<false, CHECK>(r + 3, a, b, c);
                a = load(r + 5);
                row<true, CHECK>(r + 4, b, c, a);
                b = load(r + 6);
                row<false, CHECK>(r + 5, c, a, b);
            }
        }
        for (; r + 1 < rb; r += 2) {
            c = load(r + 1);
            row<true, CHECK>(r, a, b, c);
            a = load(r + 2);
            row<false, CHECK>(r + 1, b, c, a);
            b = a;
            a = c;
        }
        if (r < rb) {
            c = load(r + 1);
            row<true, CHECK>(r, a, b, c);
        }
    }
};

// MODE: bit 0 = static-tile skipping (kTrack), bit 1 = peer stores into the neighbouring slabs (kP2P), with the
// passes ordered between the GPUs inside the kernel when p.p2p_sync is set.
// Compile-time, so that the plain single-slab kernel carries none of their code (the strict variant runs at
// its 64-register cap, and the write-back loop without peer stores is a third of the size).
enum { kPlain = 0, kTrack = 1, kP2P = 2 };

template <class Math, int NT, int MODE = kPlain>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 3 : 2))
sweep2d_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ Sweep2DParams p,
               const __grid_constant__ Math math_in)
{
    constexpr bool TRACK = (MODE & kTrack) != 0;
    constexpr bool P2P = (MODE & kP2P) != 0;
    if (*p.ctrl_done) {
        // A previous check sweep already met the termination rule: the pass retires as a no-op.  The neighbouring
        // slabs retire the same pass the same way (the decision is global), but the host has counted it, so the
        // pass index still has to reach the neighbours' flag words or the first pass after the solve would wait
        // for it forever.
        if (P2P && p.p2p_sync != 0 && blockIdx.x == 0 && threadIdx.x == 0) {
            if (p.signal_up != nullptr) st_release_sys(p.signal_up, p.pass_index);
            if (p.signal_dn != nullptr) st_release_sys(p.signal_dn, p.pass_index);
        }
        return;
    }
    const bool p2p = P2P && p.p2p_sync != 0;    // edge tile rows first only when they synchronise in the kernel
    const int tx = blockIdx.x % p.ntx, ty = tile_row_of(blockIdx.x / p.ntx, (int)p.nty, p2p);
    const int by0 = (int)p.own_lo + ty * (int)p.out_h - (int)p.T;  // buffer row of tile row 0
    if (TRACK) {
        // Tiles that read ghost rows (their input is also written by a neighbouring slab) always run.  For the
        // others every thread reads the same nine bytes (broadcast, L2 hits): the decision is CTA-uniform.
        const bool reads_ghost = (by0 < (int)p.own_lo && p.grow0 + (int)p.own_lo > 0) ||
                                 (by0 + (int)p.TH > (int)p.own_hi && p.grow0 + (int)p.own_hi < (int)p.m0);
        uint32_t any = reads_ghost ? 1u : 0u;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = tx + dx, y = ty + dy;
                if (x >= 0 && x < (int)p.ntx && y >= 0 && y < (int)p.nty) {
                    any |= p.chg_prev[y * (int)p.ntx + x];
                }
            }
        }
        if (any == 0) {
            if (threadIdx.x == 0) {
                p.chg_out[ty * (int)p.ntx + tx] = 0;
                atomicAdd(p.skipped, 1u);
            }
            return;
        }
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);
    uint8_t *lockt = smem_raw + (size_t)p.TH * kTileW * sizeof(float);
    MathTables *tables = reinterpret_cast<MathTables *>(lockt + (size_t)p.TH * kGroups);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tables + 1);
    float *s_red = reinterpret_cast<float *>(bar + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    // read through a shuffle so that the compiler knows the warp index is warp-uniform: band limits and loop
    // trip counts then live in uniform registers and the warp-synchronous operations in the row loop need
    // no convergence checks
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int gx0 = tx * (int)p.out_w - (int)p.HC;                 // grid column of tile column 0
    // does this tile read the ghost rows of / store into a neighbour?
    const bool edge_up = P2P && p.p2p_sync != 0 && p.signal_up != nullptr && by0 < (int)p.own_lo;
    const bool edge_dn = P2P && p.p2p_sync != 0 && p.signal_dn != nullptr && by0 + (int)p.TH > (int)p.own_hi;

    if (tid == 0) {
        mbar_init(bar, 1);
        if (P2P && edge_up) {
            while (ld_acquire_sys(p.wait_up) + 1u < p.pass_index) {
                __nanosleep(64);
            }
        }
        if (P2P && edge_dn) {
            while (ld_acquire_sys(p.wait_dn) + 1u < p.pass_index) {
                __nanosleep(64);
            }
        }
        if (P2P && (edge_up || edge_dn)) {
            // the ghost rows were written by another GPU's generic-proxy stores; the tile load below reads
            // them through the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
        }
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, p.TH * kTileW * (uint32_t)sizeof(float));
        tma_load_2d(tile, &src_map, gx0, by0, bar);
        // Warm L2 with the tile the CTA that follows this one on the SM will ask for (never a tile that is
        // still waiting for a neighbour: edge tiles are at the front of the launch order).
        const uint32_t nb = blockIdx.x + p.prefetch_stride;
        if (nb < gridDim.x) {
            const int nty_row = tile_row_of((int)(nb / p.ntx), (int)p.nty, p2p);
            tma_prefetch_2d(&src_map, (int)(nb % p.ntx) * (int)p.out_w - (int)p.HC,
                            (int)p.own_lo + nty_row * (int)p.out_h - (int)p.T);
        }
    }

    // While the TMA is in flight: libm tables (strict arithmetic only) and the tile's may-update nibbles.
    if (Math::kUsesTables) {
        load_math_tables(tables, tid, NT);
    }
    {
        const int w0 = gx0 >> 5;        // floor division, gx0 may be negative
        const int sh = gx0 - (w0 << 5); // 0, 4, ..., 28
        for (int item = tid; item < (int)p.TH * 8; item += NT) {
            const int r = item >> 3, j = item & 7;
            const int b = by0 + r;
            const int g = p.grow0 + b;              // global row
            uint32_t bits = 0;
            if (b >= 0 && b < (int)p.buf_rows && r > 0 && r < (int)p.TH - 1 && g > 0 && g < (int)p.m0 - 1) {
                const uint32_t *row = p.freemask + (size_t)b * p.mask_wpr;
                const int wa = w0 + j, wb = wa + 1;
                const uint32_t a = (wa >= 0 && wa < (int)p.mask_wpr) ? __ldg(row + wa) : 0u;
                const uint32_t c = (wb >= 0 && wb < (int)p.mask_wpr) ? __ldg(row + wb) : 0u;
                bits = __funnelshift_r(a, c, sh);
                if (j == 0) bits &= ~1u;            // tile column 0: no left neighbour in the tile
                if (j == 7) bits &= ~0x80000000u;   // tile column 255
                // the global border columns 0 and m1-1 are never updated (harmonic_cpu.cpp:46-51)
                const int x_first = gx0 + j * 32;   // grid column of bit 0
                if (x_first <= 0 && x_first + 31 >= 0) bits &= ~(1u << (0 - x_first));
                const int x_last = (int)p.m1 - 1;
                if (x_first <= x_last && x_first + 31 >= x_last) bits &= ~(1u << (x_last - x_first));
            }
            uint2 packed;
            packed.x = (bits & 0xFu) | ((bits & 0xF0u) << 4) | ((bits & 0xF00u) << 8) | ((bits & 0xF000u) << 12);
            bits >>= 16;
            packed.y = (bits & 0xFu) | ((bits & 0xF0u) << 4) | ((bits & 0xF00u) << 8) | ((bits & 0xF000u) << 12);
            *reinterpret_cast<uint2 *>(lockt + r * kGroups + j * 8) = packed;
        }
    }
    mbar_wait(bar, 0);
    __syncthreads();

    Math math = math_in;
    math.bind(tables);

    const int panel = warp & 1;
    const int band = warp >> 1;
    constexpr int kBands = NT / 64;
    const int col = panel * 128 + lane * 4;
    const int grp = panel * 32 + lane;
    const bool col_out = (col >= (int)p.HC) && (col < kTileW - (int)p.HC) && (gx0 + col < (int)p.m1);
    float dmax = 0.0f;
    bool chg = false;

    for (uint32_t t = 0; t < p.count; ++t) {
        // rows that still matter for the output region after the remaining sweeps
        const int reach = (int)(p.count - 1 - t);
        const int lo = (int)p.T - reach, hi = (int)p.TH - (int)p.T + reach;  // [lo, hi)
        const int per = (hi - lo + kBands - 1) / kBands;
        const int ra = lo + band * per;
        const int rb = min(ra + per, hi);
        const int pb = (int)((p.parity0 + t + (uint32_t)(by0 & 1) + (uint32_t)(gx0 & 1)) & 1u);
        const bool checking = p.check && (t + 1 == p.count);

        if (ra < rb) {
            RowCtx<Math, TRACK> cx{math, tile, lockt, col, grp, lane, by0, checking && col_out, (int)p.T, (int)p.TH,
                                   (int)p.own_lo, (int)p.own_hi, dmax, chg};
            if (checking) {
                cx.template band<true>(ra, rb, pb);
            } else {
                cx.template band<false>(ra, rb, pb);
            }
            dmax = cx.dmax;
            chg = cx.chg;
        }
        __syncthreads();
    }
    if (TRACK) {
        const int any = __syncthreads_or(chg ? 1 : 0);
        if (tid == 0) {
            p.chg_out[ty * (int)p.ntx + tx] = any ? 1 : 0;
        }
    }

    // Write the output region (all of it, locked cells included: dst is a different buffer).
    // A warp stores whole rows (lane = float4 group, two groups per lane): the per-lane source and
    // destination advance by a constant per row, no per-item index arithmetic.  Only kernels built for peer
    // stores (kP2P) carry the code that mirrors the slab's edge rows into the neighbours.
    {
        const int r_end = min((int)p.TH - (int)p.T, (int)p.own_hi - by0);
        const int c_end = min(kTileW - (int)p.HC, (int)p.pitch - gx0);
        const int r0 = max((int)p.T, (int)p.own_lo - by0) + warp;
        const int c0 = (int)p.HC + lane * 4;
        const bool in0 = c0 < c_end, in1 = c0 + 128 < c_end;
        const float *trow = tile + r0 * kTileW + c0;
        float *drow = p.dst + (size_t)(by0 + r0) * p.pitch + (gx0 + c0);
        const size_t dstep = (size_t)(NT / 32) * p.pitch;
        for (int r = r0; r < r_end; r += NT / 32) {
            float4 v0, v1;
            if (in0) v0 = *reinterpret_cast<const float4 *>(trow);
            if (in1) v1 = *reinterpret_cast<const float4 *>(trow + 128);
            if (in0) *reinterpret_cast<float4 *>(drow) = v0;
            if (in1) *reinterpret_cast<float4 *>(drow + 128) = v1;
            if (P2P) {
                const int b = by0 + r;
                if (p.peer_up != nullptr && b < (int)(p.own_lo + p.halo_rows)) {
                    float *urow = p.peer_up + (size_t)(b - (int)p.own_lo) * p.pitch + (gx0 + c0);
                    if (in0) *reinterpret_cast<float4 *>(urow) = v0;
                    if (in1) *reinterpret_cast<float4 *>(urow + 128) = v1;
                }
                if (p.peer_down != nullptr && b >= (int)(p.own_hi - p.halo_rows)) {
                    float *lrow = p.peer_down + (size_t)(b - (int)(p.own_hi - p.halo_rows)) * p.pitch + (gx0 + c0);
                    if (in0) *reinterpret_cast<float4 *>(lrow) = v0;
                    if (in1) *reinterpret_cast<float4 *>(lrow + 128) = v1;
                }
            }
            trow += (NT / 32) * kTileW;
            drow += dstep;
        }
    }

    if (P2P && (edge_up || edge_dn)) {
        __threadfence_system();     // this thread's stores into the neighbour are visible system-wide ...
        __syncthreads();            // ... for every thread of the CTA before thread 0 counts the tile as done
        if (tid == 0) {
            if (edge_up && atomicAdd(p.edge_count + 0, 1u) + 1u == p.n_edge_up) {
                p.edge_count[0] = 0u;   // the next pass starts after this kernel has ended
                __threadfence_system();
                st_release_sys(p.signal_up, p.pass_index);
            }
            if (edge_dn && atomicAdd(p.edge_count + 1, 1u) + 1u == p.n_edge_dn) {
                p.edge_count[1] = 0u;
                __threadfence_system();
                st_release_sys(p.signal_dn, p.pass_index);
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
            for (int i = 0; i < NT / 32; ++i) {
                m = fmaxf(m, s_red[i]);
            }
            if (m > 0.0f) {
                atomicMax(p.delta_bits, __float_as_uint(m));
            }
        }
    }
}

}  // namespace epic_b200
