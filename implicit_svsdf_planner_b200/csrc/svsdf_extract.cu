// svsdf_extract.cu — K3: query-point construction on the device from the bit-packed map kernel, one z layer per launch.
//
// Restates what the reference does on the host before every optimisation
//   PlannerManager::generateTraj                         src/plan_manager/src/plan_manager.cpp:156-175
//   PCSmapManager::getPointsInAABBOutOfLastOne           src/map_manager/include/map_manager/PCSmap_manager.h:184-219
//   projInMap / unifiedID                                PCSmap_manager.h:118-135
//   GridMap3D::getGridIndex / getGridCubeCenter          src/map_manager/src/Gridmap3D.cpp:137-195
// on the map representation the reference broadcasts to its front end, the byte-packed "map kernel"
//   PCSmapManager::generateMapKernel2D                   PCSmap_manager.h:81-108  ((X+2h) x ceil((Y+2h)/8) bytes, MSB first).
// A cell is a query point iff it is occupied and lies in the AABB of some waypoint w but outside the AABB of waypoint
// w-1 (the reference visits, per waypoint, only the cells outside the previous box and de-duplicates by cell id; for the first
// waypoint the "previous" box is the one around tmp_pos = (999, 999, 999), the map's far corner cell).  A 3-D map is handled
// layer by layer (svsdf_extract_points3d): the host marks, per layer, which waypoint boxes contain it; occupied cells of several
// layers above the same (x, y) come out as several points with that (x, y) — the cost loop zeroes z.  The
// reference enumerates an unordered_map (unspecified order); here points come out in ascending (i * Y + j) order, which
// is the memory order of the packed map, so loads are coalesced and the result is deterministic.
//
// Integer/byte work, HBM-bound in principle (reads the packed rows of the boxes' bounding rectangle, writes 16 B per
// point) — three small kernels: count (popcount per 32-cell word) -> scan of CTA totals -> ordered write.
#include <cuda_runtime.h>
#include <stdint.h>

#include "svsdf_types.h"

namespace svsdf {

namespace {

__device__ __forceinline__ unsigned interval_mask(int lo, int hi) {  // bits lo..hi (inclusive) of a 32-bit word
    lo = lo < 0 ? 0 : lo;
    hi = hi > 31 ? 31 : hi;
    if (lo > hi) return 0u;
    return (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
}

// 32 cells of row x starting at padded y index 32*wy, bit k <-> padded y = 32*wy + k
__device__ __forceinline__ unsigned load_cells(const ExtractArgs &E, int x, int wy) {
    const unsigned char *row = E.map + (int64_t)(x + E.h) * E.row_bytes;
    const int b0 = 4 * wy;
    unsigned w = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (b0 + q < E.row_bytes) w |= (unsigned)row[b0 + q] << (8 * q);
    // MSB-first within each byte (or_mask = {0x80 .. 0x01}): reverse the bits of every byte
    return __byte_perm(__brev(w), 0, 0x0123);
}

__device__ __forceinline__ unsigned selected_cells(const ExtractArgs &E, int x, int wy) {
    const int ybase = 32 * wy - E.h;  // real y index of bit 0
    unsigned member = 0u;
    for (int w = 0; w < E.W; ++w) {
        if (!E.act[w] || x < E.bx1[w] || x > E.bx2[w]) continue;
        unsigned m = interval_mask(E.by1[w] - ybase, E.by2[w] - ybase);
        if (E.excl[w]) {  // "OutOfLastOne": skip the cells of the previous waypoint's box (w = 0: the box around tmp_pos)
            const int lx1 = w ? E.bx1[w - 1] : E.px1, lx2 = w ? E.bx2[w - 1] : E.px2;
            const int ly1 = w ? E.by1[w - 1] : E.py1, ly2 = w ? E.by2[w - 1] : E.py2;
            if (x >= lx1 && x <= lx2) m &= ~interval_mask(ly1 - ybase, ly2 - ybase);
        }
        member |= m;
    }
    if (member == 0u) return 0u;
    unsigned sel = load_cells(E, x, wy) & member;
    if (sel != 0u && E.n_keepout > 0) {
        // synthetic-scene option (not in the reference): drop cells within `clearance` of a keep-out polyline, the
        // stand-in for the A* front end's guarantee that the nominal path itself is collision free
        unsigned keep = 0u;
        const double c2 = E.clearance * E.clearance;
        const double cx = ((double)x + 0.5) * E.res + E.ox;
        for (unsigned bits = sel; bits; bits &= bits - 1) {
            const int k = __ffs(bits) - 1;
            const double cy = ((double)(ybase + k) + 0.5) * E.res + E.oy;
            bool ok = true;
            for (int q = 0; q < E.n_keepout && ok; ++q) {
                const double dx = cx - E.keepout[2 * q], dy = cy - E.keepout[2 * q + 1];
                ok = dx * dx + dy * dy > c2;
            }
            if (ok) keep |= 1u << k;
        }
        sel = keep;
    }
    return sel;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int s = (lane < nw) ? wsum[lane] : 0, si = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, si, off);
            if (lane >= off) si += u;
        }
        wsum[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    return wsum[warp] + inc - v;
}

__global__ void __launch_bounds__(256) k_extract_count(const __grid_constant__ ExtractArgs E, int *block_counts) {
    __shared__ int total;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (item < E.n_items) {
        const int x = E.rx1 + (int)(item / E.nW), wy = E.wy1 + (int)(item % E.nW);
        cnt = __popc(selected_cells(E, x, wy));
    }
    block_exclusive_scan(cnt, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_extract_scan(int *block_counts, int n_blocks, int64_t *n_total) {
    // exclusive scan of the CTA totals in place (single CTA, chunked)
    __shared__ int total;
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < n_blocks) ? block_counts[i] : 0;
        const int ex = block_exclusive_scan(v, &total);
        if (i < n_blocks) block_counts[i] = (int)(carry + ex);
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_total = carry;
}

__global__ void __launch_bounds__(256) k_extract_write(const __grid_constant__ ExtractArgs E, const int *block_offsets,
                                                       double *out_xy, int64_t cap) {
    __shared__ int total;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned sel = 0u;
    int x = 0, wy = 0;
    if (item < E.n_items) {
        x = E.rx1 + (int)(item / E.nW);
        wy = E.wy1 + (int)(item % E.nW);
        sel = selected_cells(E, x, wy);
    }
    int64_t pos = (int64_t)block_offsets[blockIdx.x] + block_exclusive_scan(__popc(sel), &total);
    const int ybase = 32 * wy - E.h;
    const double cx = ((double)x + 0.5) * E.res + E.ox;  // getGridCubeCenter: (idx + 0.5) * res + min
    for (; sel; sel &= sel - 1) {
        const int k = __ffs(sel) - 1;
        if (pos < cap) {
            out_xy[2 * pos] = cx;
            out_xy[2 * pos + 1] = ((double)(ybase + k) + 0.5) * E.res + E.oy;
        }
        ++pos;
    }
}

}  // namespace

cudaError_t launch_extract_count(const ExtractArgs &E, int *block_counts, int n_blocks, int64_t *n_total,
                                 cudaStream_t stream) {
    if (n_blocks > 0) k_extract_count<<<n_blocks, 256, 0, stream>>>(E, block_counts);
    k_extract_scan<<<1, 1024, 0, stream>>>(block_counts, n_blocks, n_total);
    return cudaGetLastError();
}

cudaError_t launch_extract_write(const ExtractArgs &E, const int *block_offsets, int n_blocks, double *out_xy,
                                 int64_t cap, cudaStream_t stream) {
    if (n_blocks > 0) k_extract_write<<<n_blocks, 256, 0, stream>>>(E, block_offsets, out_xy, cap);
    return cudaGetLastError();
}

}  // namespace svsdf
