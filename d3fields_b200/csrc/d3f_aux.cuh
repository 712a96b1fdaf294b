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

// ---------------------------------------------------------------------------------------------------------------------
// rigid_tracking (reference fusion.py:1608-1685): the two small kernels that, together with the field query and its
// backward, make one Adam iteration four launches (d3fields_b200/tracking.py, fused=True).
//
//   track_loss_grad_kernel   d loss / d feat and d loss / d dist of
//                                loss = mean_p(|feat_p - src_p|_2 * valid_p) + dist_w * mean_p(max(dist_p * valid_p, 0)) + reg
//                            (fusion.py:1651-1662); one warp per point.  torch's norm backward is diff/|diff| (0 at 0).
//   track_update_kernel      per instance i: chain d loss / d pts (from d3f_eval_backward) through
//                                pts = last_pts @ R(log_r) + t,  R = so3_exp_map(log_r)   (pytorch3d: I + sin(a)/a K + (1-cos a)/a^2 K^2,
//                                a = sqrt(max(|log_r|^2, 1e-4)))
//                            to t and log_r, add the gradient of reg_w * (|t|_F + |log_r|_F) (norms over ALL instances, 0 at 0),
//                            take one Adam step (torch.optim.Adam defaults: betas, eps 1e-8, no weight decay, bias
//                            correction as torch computes it) and write the NEXT iteration's points.
constexpr int TRACK_THREADS = 128;

__global__ void __launch_bounds__(256)
track_loss_grad_kernel(const float* __restrict__ feat, const float* __restrict__ src, const float* __restrict__ dist,
                       const uint8_t* __restrict__ valid, int n, int C, float dist_w,
                       float* __restrict__ g_feat, float* __restrict__ g_dist, float* __restrict__ loss_terms) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const float* f = feat + (size_t)p * C;
    const float* s = src + (size_t)p * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = f[c] - s[c]; ss = fmaf(d, d, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const float v = valid[p] ? 1.f : 0.f;
    const float scale = (nrm > 0.f) ? v / (nrm * (float)n) : 0.f;
    float* g = g_feat + (size_t)p * C;
    for (int c = lane; c < C; c += 32) g[c] = (f[c] - s[c]) * scale;
    if (lane == 0) {
        const float dv = dist[p] * v;
        g_dist[p] = (dv >= 0.f) ? dist_w * v / (float)n : 0.f;                // torch's clamp(min=0) passes the gradient where x >= 0
        if (loss_terms) loss_terms[p] = (nrm * v + dist_w * fmaxf(dv, 0.f)) / (float)n;
    }
}

struct TrackParams {
    const float* t_in;        // (I,3) translation / axis-angle of this iteration ...
    const float* r_in;
    float* t_out;             // ... and of the next (ping-pong buffers: every block reads ALL instances' old parameters
    float* r_out;             //     for the regulariser's norms, so new ones must not land in the same array)
    float* m_t; float* v_t;   // Adam moments of t      (I,3), updated in place
    float* m_r; float* v_r;   // Adam moments of log_r  (I,3)
    const float* last_pts;    // (I,P,3)
    const float* grad_pts;    // (I*P,3) d loss / d pts of this iteration, or nullptr: no update, only transform with *_in
    float* pts;               // (I*P,3) out: points of the next iteration (nullptr: leave them, e.g. after the last step)
    int I, P;
    float step;               // Adam step number of this update (1, 2, ...)
    float bc1, bc2s;          // 1 - beta1^step, sqrt(1 - beta2^step)
    float lr, beta1, beta2, eps, reg_w;
};

__device__ __forceinline__ void so3_exp(const float w[3], float R[9], float& a2_clamped, float& f1, float& f2) {
    const float n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    a2_clamped = fmaxf(n2, 1e-4f);
    const float a = sqrtf(a2_clamped), inv = 1.f / a;
    f1 = inv * sinf(a);
    f2 = inv * inv * (1.f - cosf(a));
    const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float kk = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) kk += K[i * 3 + k] * K[k * 3 + j];
            R[i * 3 + j] = f1 * K[i * 3 + j] + f2 * kk + (i == j ? 1.f : 0.f);
        }
}

// One instance's update: reduce d loss / d pts over its points, chain to (t, log_r), regulariser, Adam step into
// (t_out, r_out); thread 0 leaves the NEW rotation / translation in sR / sT.  Called by all THREADS threads of a CTA.
// L2_LOADS: grad_pts was written by other CTAs of the SAME launch (track_step_kernel) — read it at the L2.
// Latency is all that matters here (one CTA, a few hundred flops): thread 0's eighteen scalars are requested before the
// reduction so their round trip overlaps it (loaded where they are used, each load would queue behind the store of the
// previous moment — the compiler cannot prove the buffers distinct), the regulariser's norms are reduced by the CTA with
// the gradient sums, the bias corrections come from the host, and the 3x3 algebra is fully unrolled into registers.
constexpr int TRACK_RED = 14;
template <int THREADS, bool L2_LOADS>
__device__ __forceinline__ void track_update_instance(const TrackParams& tp, int i, float (*red)[TRACK_RED], float* sR, float* sT) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* last = tp.last_pts + (size_t)i * tp.P * 3;
    float w3[3], t3[3], mt[3], vt[3], mr[3], vr[3];
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            w3[j] = tp.r_in[i * 3 + j]; t3[j] = tp.t_in[i * 3 + j];
            mt[j] = tp.m_t[i * 3 + j]; vt[j] = tp.v_t[i * 3 + j]; mr[j] = tp.m_r[i * 3 + j]; vr[j] = tp.v_r[i * 3 + j];
        }
    }
    // g_t = sum_p g_p ;  M[a][b] = sum_p last_p[a] * g_p[b]   (d loss / d R for pts = last @ R) ;  |t|_F^2, |log_r|_F^2
    float acc[TRACK_RED];
#pragma unroll
    for (int k = 0; k < TRACK_RED; ++k) acc[k] = 0.f;
    const float* g = tp.grad_pts + (size_t)i * tp.P * 3;
    for (int p = tid; p < tp.P; p += THREADS) {
        const float gx = L2_LOADS ? __ldcg(g + p * 3) : g[p * 3], gy = L2_LOADS ? __ldcg(g + p * 3 + 1) : g[p * 3 + 1];
        const float gz = L2_LOADS ? __ldcg(g + p * 3 + 2) : g[p * 3 + 2];
        const float lx = last[p * 3], ly = last[p * 3 + 1], lz = last[p * 3 + 2];
        acc[0] += gx; acc[1] += gy; acc[2] += gz;
        acc[3] += lx * gx; acc[4] += lx * gy; acc[5] += lx * gz;
        acc[6] += ly * gx; acc[7] += ly * gy; acc[8] += ly * gz;
        acc[9] += lz * gx; acc[10] += lz * gy; acc[11] += lz * gz;
    }
    for (int k = tid; k < tp.I * 3; k += THREADS) {       // regulariser: Frobenius norms over ALL instances (fusion.py:1654)
        const float a = tp.t_in[k], b = tp.r_in[k];
        acc[12] = fmaf(a, a, acc[12]); acc[13] = fmaf(b, b, acc[13]);
    }
#pragma unroll
    for (int k = 0; k < TRACK_RED; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0) red[warp][k] = acc[k];
    }
    __syncthreads();
    if (tid == 0) {
        float G[TRACK_RED];
#pragma unroll
        for (int k = 0; k < TRACK_RED; ++k) {
            G[k] = 0.f;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) G[k] += red[w][k];
        }
        const float* M = G + 3;
        const float nt = sqrtf(G[12]), nr = sqrtf(G[13]);
        // d R / d w_j  (K = hat(w);  dK_j = hat(e_j))
        float R[9], a2, f1, f2;
        so3_exp(w3, R, a2, f1, f2);
        const float n2 = w3[0] * w3[0] + w3[1] * w3[1] + w3[2] * w3[2];
        const float a = sqrtf(a2);
        float df1 = 0.f, df2 = 0.f;                         // d f / d a (zero while the angle is clamped)
        if (n2 > 1e-4f) {
            df1 = (a * cosf(a) - sinf(a)) / (a * a);
            df2 = (a * sinf(a) - 2.f * (1.f - cosf(a))) / (a * a * a);
        }
        const float K[9] = {0.f, -w3[2], w3[1], w3[2], 0.f, -w3[0], -w3[1], w3[0], 0.f};
        float KK[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * K[k * 3 + c];
                KK[r * 3 + c] = s;
            }
        float g_w[3], g_t[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float dK[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (j == 0) { dK[5] = -1.f; dK[7] = 1.f; } else if (j == 1) { dK[2] = 1.f; dK[6] = -1.f; } else { dK[1] = -1.f; dK[3] = 1.f; }
            const float da = (n2 > 1e-4f) ? w3[j] / a : 0.f;
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float dKK = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) dKK += dK[r * 3 + k] * K[k * 3 + c] + K[r * 3 + k] * dK[k * 3 + c];
                    const float dR = f1 * dK[r * 3 + c] + f2 * dKK + df1 * da * K[r * 3 + c] + df2 * da * KK[r * 3 + c];
                    s += M[r * 3 + c] * dR;
                }
            g_w[j] = s + (nr > 0.f ? tp.reg_w * w3[j] / nr : 0.f);
            g_t[j] = G[j] + (nt > 0.f ? tp.reg_w * t3[j] / nt : 0.f);
        }
        // Adam (torch.optim.Adam, single-tensor path): step_size = lr / (1 - b1^k), denom = sqrt(v)/sqrt(1 - b2^k) + eps;
        // bc1 = 1 - b1^k and bc2s = sqrt(1 - b2^k) are computed on the host (d3f_abi.cu)
        const float step_size = tp.lr / tp.bc1;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            mt[j] = tp.beta1 * mt[j] + (1.f - tp.beta1) * g_t[j];
            vt[j] = tp.beta2 * vt[j] + (1.f - tp.beta2) * g_t[j] * g_t[j];
            t3[j] -= step_size * mt[j] / (sqrtf(vt[j]) / tp.bc2s + tp.eps);
            mr[j] = tp.beta1 * mr[j] + (1.f - tp.beta1) * g_w[j];
            vr[j] = tp.beta2 * vr[j] + (1.f - tp.beta2) * g_w[j] * g_w[j];
            w3[j] -= step_size * mr[j] / (sqrtf(vr[j]) / tp.bc2s + tp.eps);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            tp.m_t[i * 3 + j] = mt[j]; tp.v_t[i * 3 + j] = vt[j]; tp.m_r[i * 3 + j] = mr[j]; tp.v_r[i * 3 + j] = vr[j];
            tp.t_out[i * 3 + j] = t3[j]; tp.r_out[i * 3 + j] = w3[j]; sT[j] = t3[j];
        }
        float a2n, f1n, f2n;
        so3_exp(w3, sR, a2n, f1n, f2n);
    }
}

__global__ void __launch_bounds__(TRACK_THREADS)
track_update_kernel(const TrackParams tp) {
    __shared__ float red[TRACK_THREADS / 32][TRACK_RED];
    __shared__ float sR[9], sT[3];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float* last = tp.last_pts + (size_t)i * tp.P * 3;
    if (tp.grad_pts) {
        track_update_instance<TRACK_THREADS, false>(tp, i, red, sR, sT);
    } else if (tid == 0) {
        float w3[3] = {tp.r_in[i * 3], tp.r_in[i * 3 + 1], tp.r_in[i * 3 + 2]};
        float a2, f1, f2;
        so3_exp(w3, sR, a2, f1, f2);
        sT[0] = tp.t_in[i * 3]; sT[1] = tp.t_in[i * 3 + 1]; sT[2] = tp.t_in[i * 3 + 2];
    }
    __syncthreads();
    if (!tp.pts) return;
    // pts = last @ R + t   (row vectors, pytorch3d Transform3d().rotate(R).translate(t))
    for (int p = tid; p < tp.P; p += TRACK_THREADS) {
        const float lx = last[p * 3], ly = last[p * 3 + 1], lz = last[p * 3 + 2];
        float* o = tp.pts + ((size_t)i * tp.P + p) * 3;
        o[0] = lx * sR[0] + ly * sR[3] + lz * sR[6] + sT[0];
        o[1] = lx * sR[1] + ly * sR[4] + lz * sR[7] + sT[1];
        o[2] = lx * sR[2] + ly * sR[5] + lz * sR[8] + sT[2];
    }
}

}  // namespace d3f
