// d3f_aux.cuh — the two small kernels either side of the field query.
#pragma once
#include "d3f_common.cuh"

namespace d3f {

constexpr int PCA_MAX_COMP = 8;

// y[i, j] = sum_c (x[i,c] - mean[c]) * comp[j,c] (mean may be null: plain x @ comp^T)  — sklearn PCA.transform as the reference applies it
// to eval()'s descriptors (reference fusion.py:1386-1392).  One warp per row: lanes stride the
// channels (coalesced 128-bit loads when C % 4 == 0), one accumulator per component, butterfly
// reduction.  HBM-bound on reading x once (4*C bytes per row).
__global__ void __launch_bounds__(256)
pca_project_kernel(const float* __restrict__ x, int64_t n, int C, const float* __restrict__ mean,
                   const float* __restrict__ comp, int n_comp, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    float acc[PCA_MAX_COMP];
#pragma unroll
    for (int j = 0; j < PCA_MAX_COMP; ++j) acc[j] = 0.f;
    const float* xr = x + (size_t)row * C;
    const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(mean) |
                                       reinterpret_cast<uintptr_t>(comp)) % 16 == 0);
    if (vec) {
        for (int c = lane * 4; c < C; c += 128) {
            const float4 xv = __ldcs(reinterpret_cast<const float4*>(xr + c));
            const float4 mv = mean ? __ldg(reinterpret_cast<const float4*>(mean + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 d = make_float4(xv.x - mv.x, xv.y - mv.y, xv.z - mv.z, xv.w - mv.w);
#pragma unroll
            for (int j = 0; j < PCA_MAX_COMP; ++j) {
                if (j < n_comp) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(comp + (size_t)j * C + c));
                    acc[j] = fmaf(d.x, w.x, acc[j]); acc[j] = fmaf(d.y, w.y, acc[j]);
                    acc[j] = fmaf(d.z, w.z, acc[j]); acc[j] = fmaf(d.w, w.w, acc[j]);
                }
            }
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            const float d = xr[c] - (mean ? __ldg(mean + c) : 0.f);
#pragma unroll
            for (int j = 0; j < PCA_MAX_COMP; ++j)
                if (j < n_comp) acc[j] = fmaf(d, __ldg(comp + (size_t)j * C + c), acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < PCA_MAX_COMP; ++j) {
        if (j < n_comp) {
            float a = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) y[(size_t)row * n_comp + j] = a;
        }
    }
}

// Voxel-centre grid, z fastest (reference fusion.py:79-88): coordinate i of an axis is
// float(lower + step*i) + float(step/2) — torch.arange's float32 values plus the half step.
__global__ void __launch_bounds__(256)
create_grid_kernel(double x_lower, double y_lower, double z_lower, double step,
                   int nx, int ny, int nz, float* __restrict__ pts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)nx * ny * nz;
    if (i >= n) return;
    const int iz = (int)(i % nz);
    const int iy = (int)((i / nz) % ny);
    const int ix = (int)(i / ((int64_t)nz * ny));
    const float half = (float)(step / 2.0);
    pts[i * 3 + 0] = __fadd_rn((float)(x_lower + step * ix), half);
    pts[i * 3 + 1] = __fadd_rn((float)(y_lower + step * iy), half);
    pts[i * 3 + 2] = __fadd_rn((float)(z_lower + step * iz), half);
}

}  // namespace d3f
