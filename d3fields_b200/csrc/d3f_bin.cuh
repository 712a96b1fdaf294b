// d3f_bin.cuh — binned traversal order for query points without spatial order.
//
// The wide walk (d3f_tile.cuh) keeps the four corner texels of every view in registers and reloads them only when a
// point's texel cell differs from its predecessor's.  A voxel grid in z-fastest order (create_init_grid,
// fusion.py:79-88) has that locality for free; keypoints and mesh vertices (vis_tracking.py:92-130, vis_repr.py:103,
// fusion.py:1449/1650) do not: every point reloads 4 views x 4 corners x 4 KB through L2 and the kernel runs at a third
// of its grid speed.  d3f_bin_order produces a visiting order that restores the locality for EVERY view at once:
// points are grouped by the cell of a cubic lattice they fall in (edge ~ one texel footprint), cells in Morton order,
// so consecutive points are a few millimetres apart in space and therefore project into the same or a neighbouring
// texel cell of all cameras.  A counting sort: bounding box -> Morton keys + histogram -> exclusive scan -> scatter.
// The order inside a cell is whatever the atomics give; the field of a point does not depend on its position in the
// sequence, so results are bit-identical to an unordered launch.
#pragma once
#include "d3f_common.cuh"

namespace d3f {

constexpr int BIN_BITS = 7;                          // lattice cells per axis: 2^7
constexpr int BIN_COUNT = 1 << (3 * BIN_BITS);       // 2 097 152 bins
constexpr int BIN_THREADS = 256;
constexpr int BIN_SCAN_THREADS = 1024;
constexpr int BIN_SCAN_PER_THREAD = 8;               // one scan block covers 8192 bins
constexpr int BIN_SCAN_BLOCKS = BIN_COUNT / (BIN_SCAN_THREADS * BIN_SCAN_PER_THREAD);   // 256

// workspace layout (bytes): [0,32) bounding box as ordered ints | hist[BIN_COUNT] | block_base[BIN_SCAN_BLOCKS] | keys[n]
__host__ __device__ inline size_t bin_workspace_bytes(long long n) {
    return 256 + sizeof(uint32_t) * ((size_t)BIN_COUNT + BIN_SCAN_BLOCKS) + sizeof(uint32_t) * (size_t)(n > 0 ? n : 0);
}

// float <-> int32 whose signed order is the float order (finite values)
__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void bin_init_kernel(int* bbox) {
    if (threadIdx.x < 3) { bbox[threadIdx.x] = 0x7fffffff; bbox[3 + threadIdx.x] = (int)0x80000000; }
}

__global__ void __launch_bounds__(BIN_THREADS)
bin_bbox_kernel(const float* __restrict__ pts, int64_t n, int* __restrict__ bbox) {
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    int hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int64_t i = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * BIN_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float f = __ldg(pts + i * 3 + a);
            if (isfinite(f)) { const int o = float_to_ordered(f); lo[a] = min(lo[a], o); hi[a] = max(hi[a], o); }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(bbox + a, lo[a]); atomicMax(bbox + 3 + a, hi[a]); }
    }
}

__device__ __forceinline__ unsigned spread3(unsigned v) {      // 7 bits -> every third bit
    v &= 0x7fu;
    v = (v | (v << 8)) & 0x0000700fu;
    v = (v | (v << 4)) & 0x000430c3u;
    v = (v | (v << 2)) & 0x00049249u;
    return v;
}

__global__ void __launch_bounds__(BIN_THREADS)
bin_key_kernel(const float* __restrict__ pts, int64_t n, float cell, const int* __restrict__ bbox,
               uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    const int64_t i = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x;
    if (i >= n) return;
    const int side = 1 << BIN_BITS;
    unsigned q[3];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ordered_to_float(bbox[a]), hi = ordered_to_float(bbox[3 + a]);
        // never more than `side` cells per axis: a point cloud wider than side*cell gets proportionally larger cells
        const float c = fmaxf(cell, (hi - lo) / (float)side * 1.0001f);
        const float f = __ldg(pts + i * 3 + a);
        ok = ok && isfinite(f);
        q[a] = (unsigned)min(max((int)floorf((f - lo) / c), 0), side - 1);
    }
    // x is the most significant axis, z the least: inside a cell column the walk still runs along z like the grid's
    const unsigned key = ok ? ((spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2])) : (unsigned)(BIN_COUNT - 1);
    keys[i] = key;
    atomicAdd(hist + key, 1u);
}

// exclusive scan of hist inside blocks of 8192 bins; block totals to block_base (scanned by bin_scan_base_kernel)
__global__ void __launch_bounds__(BIN_SCAN_THREADS)
bin_scan_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ block_base) {
    __shared__ uint32_t warp_sum[BIN_SCAN_THREADS / 32];
    const int t = threadIdx.x;
    uint4* h4 = reinterpret_cast<uint4*>(hist + ((size_t)blockIdx.x * BIN_SCAN_THREADS + t) * BIN_SCAN_PER_THREAD);
    uint4 a = h4[0], b = h4[1];
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t e[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { e[j] = s; s += v[j]; }
    uint32_t inc = s;                                   // inclusive scan of the thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += u;
    }
    if ((t & 31) == 31) warp_sum[t >> 5] = inc;
    __syncthreads();
    if (t < 32) {
        uint32_t w = warp_sum[t];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (t >= o) w += u;
        }
        warp_sum[t] = w;
    }
    __syncthreads();
    const uint32_t base = inc - s + ((t >> 5) ? warp_sum[(t >> 5) - 1] : 0u);
    h4[0] = make_uint4(base + e[0], base + e[1], base + e[2], base + e[3]);
    h4[1] = make_uint4(base + e[4], base + e[5], base + e[6], base + e[7]);
    if (t == BIN_SCAN_THREADS - 1) block_base[blockIdx.x] = warp_sum[31];
}

__global__ void __launch_bounds__(BIN_SCAN_BLOCKS)
bin_scan_base_kernel(uint32_t* __restrict__ block_base) {
    __shared__ uint32_t warp_sum[BIN_SCAN_BLOCKS / 32];
    const int t = threadIdx.x;
    const uint32_t v = block_base[t];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += u;
    }
    if ((t & 31) == 31) warp_sum[t >> 5] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < (t >> 5); ++w) base += warp_sum[w];
    block_base[t] = base + inc - v;
}

__device__ __forceinline__ unsigned spread3_10(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Second level: inside every aligned block of 256 positions of `order` (= one tile of the ordered field kernel) sort
// by a Morton key eight times finer per axis than the lattice (30 bits).  The lattice bins hold ~10 points in no
// particular order; after this pass consecutive points of a tile are nearest neighbours along the z-curve, so the walk
// re-enters a texel cell far less often (measured: L2->L1 read sectors of the walk, profiles/r02_binned*_summary.txt).
// One CTA per block, bitonic sort in shared memory.
constexpr int BIN_FINE_BITS = BIN_BITS + 3;
__global__ void __launch_bounds__(256)
bin_refine_kernel(const float* __restrict__ pts, int64_t n, float cell, const int* __restrict__ bbox,
                  int32_t* __restrict__ order) {
    __shared__ unsigned long long kv[256];
    const int t = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * 256 + t;
    unsigned long long e = ~0ull;                       // padding sorts to the end
    if (i < n) {
        const int row = order[i];
        const int side = 1 << BIN_BITS, fine = 1 << BIN_FINE_BITS;
        unsigned q[3];
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float lo = ordered_to_float(bbox[a]), hi = ordered_to_float(bbox[3 + a]);
            const float c = fmaxf(cell, (hi - lo) / (float)side * 1.0001f) * 0.125f;
            const float f = __ldg(pts + (size_t)row * 3 + a);
            ok = ok && isfinite(f);
            q[a] = (unsigned)min(max((int)floorf((f - lo) / c), 0), fine - 1);
        }
        const unsigned key = ok ? ((spread3_10(q[0]) << 2) | (spread3_10(q[1]) << 1) | spread3_10(q[2])) : 0x3fffffffu;
        e = ((unsigned long long)key << 32) | (unsigned)row;
    }
    kv[t] = e;
    __syncthreads();
    for (int k = 2; k <= 256; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int partner = t ^ j;
            if (partner > t) {
                const unsigned long long a = kv[t], b = kv[partner];
                const bool up = (t & k) == 0;
                if ((a > b) == up) { kv[t] = b; kv[partner] = a; }
            }
            __syncthreads();
        }
    }
    if (i < n) order[i] = (int32_t)(unsigned)(kv[t] & 0xffffffffu);
}

__global__ void __launch_bounds__(BIN_THREADS)
bin_scatter_kernel(const uint32_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ cursor,
                   const uint32_t* __restrict__ block_base, int32_t* __restrict__ order) {
    const int64_t i = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = keys[i];
    const uint32_t pos = atomicAdd(cursor + key, 1u) + block_base[key / (BIN_SCAN_THREADS * BIN_SCAN_PER_THREAD)];
    order[pos] = (int32_t)i;
}

}  // namespace d3f
