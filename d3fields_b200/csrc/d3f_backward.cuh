// d3f_backward.cuh — gradient of the field query with respect to the query points.
//
// The reference's rigid_tracking (fusion.py:1608-1685) runs Adam through Fusion.eval with autograd:
//   loss(feat(pts), dist(pts)) -> d loss / d pts.
// What is differentiable in fusion.py:305-394 (everything else is a hard mask or a nearest-neighbour
// lookup, whose gradient torch defines as zero):
//   * the bilinear samples, through their pixel coordinates (grid_sample backward, align_corners=True,
//     zero padding: out-of-range corners contribute nothing),
//   * the distance weight exp(min(mu-|d|,0)/mu), through d = depth_nearest - z  (only where |d| > mu),
//   * dist = sum_v clamp(d_v,-mu,mu) vis_v / (count+1e-6), through z (only where -mu <= d_v <= mu and at
//     least one view sees the point: the 1e3 fill of fusion.py:367 is a constant),
//   * pixel coordinates and z through the projection H_v [x y z 1]^T (the |z|<1e-4 patch makes the view
//     invisible, so it carries no gradient).
//
// One warp per point (tracking evaluates a few hundred points; latency, not throughput, matters):
// lanes 0..V-1 redo the per-view forward in parallel, then for every visible view and key the lanes
// stride the channels and accumulate  sum_c g_c dr_c/dix,  sum_c g_c dr_c/diy,  sum_c g_c r_c,
// reduced with a butterfly; lane 0 chains them to the camera frame and to the point.
#pragma once
#include "d3f_common.cuh"
#include "d3f_generic.cuh"

namespace d3f {

constexpr int BWD_WARPS = 4;

struct BwdKeySet {
    const void* data[D3F_MAX_KEYS];
    const float* grad[D3F_MAX_KEYS];     // (n,C) upstream gradient of the key's output, or nullptr
    int32_t dtype[D3F_MAX_KEYS];
    int32_t h[D3F_MAX_KEYS], w[D3F_MAX_KEYS], C[D3F_MAX_KEYS];
    int64_t sv[D3F_MAX_KEYS];            // element strides of the view / row / texel axes (channel stride 1)
    int32_t sy[D3F_MAX_KEYS], sx[D3F_MAX_KEYS];
    int32_t vec4[D3F_MAX_KEYS];          // 1: float32 map, C % 4 == 0, 16-byte aligned map and gradient -> 128-bit loads
    int32_t n_keys;
};

// d loss / d pts contributed by ONE visible view v of point i (all 32 lanes of a warp take part: they stride the
// channels); the result is valid in every lane.  sv = {px, py, cz, d, weight}, inv = 1/(count+1e-6), gd = d loss/d dist.
template <bool RECIP>
__device__ __forceinline__ float3 view_gradient(const EvalParams& ep, const BwdKeySet& ks, int64_t i, int v,
                                                const float* sv, float inv, float gd, const float* sH, int lane) {
        const float px = sv[0], py = sv[1], cz = sv[2], d = sv[3], weight = sv[4];
        const float fac = weight * inv;
        float G_px = 0.f, G_py = 0.f, G_w = 0.f;
        for (int k = 0; k < ks.n_keys; ++k) {
            if (!ks.grad[k]) continue;
            const int h = ks.h[k], w = ks.w[k], C = ks.C[k];
            // footprint with explicit in-range flags (a zero weight can also be a cell-border weight)
            const float ix = to_map_index<RECIP>(px, ep.W, w), iy = to_map_index<RECIP>(py, ep.H, h);
            const float x0 = floorf(ix), y0 = floorf(iy);
            const float wx = ix - x0, wy = iy - y0;
            const float x1 = x0 + 1.f, y1 = y0 + 1.f, xm = (float)(w - 1), ym = (float)(h - 1);
            const bool x0ok = x0 >= 0.f && x0 <= xm, x1ok = x1 >= 0.f && x1 <= xm;
            const bool y0ok = y0 >= 0.f && y0 <= ym, y1ok = y1 >= 0.f && y1 <= ym;
            const int x0c = (int)fminf(fmaxf(x0, 0.f), xm), x1c = (int)fminf(fmaxf(x1, 0.f), xm);
            const int y0c = (int)fminf(fmaxf(y0, 0.f), ym), y1c = (int)fminf(fmaxf(y1, 0.f), ym);
            const float m00 = (x0ok && y0ok) ? 1.f : 0.f, m01 = (x1ok && y0ok) ? 1.f : 0.f;
            const float m10 = (x0ok && y1ok) ? 1.f : 0.f, m11 = (x1ok && y1ok) ? 1.f : 0.f;
            const size_t vbase = (size_t)v * (size_t)ks.sv[k];
            const size_t sy = (size_t)ks.sy[k], sx = (size_t)ks.sx[k];
            const size_t o00 = vbase + (size_t)y0c * sy + (size_t)x0c * sx, o01 = vbase + (size_t)y0c * sy + (size_t)x1c * sx;
            const size_t o10 = vbase + (size_t)y1c * sy + (size_t)x0c * sx, o11 = vbase + (size_t)y1c * sy + (size_t)x1c * sx;
            const float* g = ks.grad[k] + (size_t)i * C;
            const float ex = 1.f - wx, ey = 1.f - wy;
            float s_ix = 0.f, s_iy = 0.f, s_r = 0.f;
            if (ks.vec4[k]) {
                // wide float32 maps (the 1024-channel descriptors): one 128-bit load per corner and lane
                const float* vol = static_cast<const float*>(ks.data[k]);
                // unrolled by two: the five 128-bit loads of two 128-channel steps are issued together (half the L2 round
                // trips — a tracking launch is a few hundred warps, nothing else hides the latency)
#pragma unroll 2
                for (int c = lane * 4; c < C; c += 128) {
                    float4 f00 = __ldg(reinterpret_cast<const float4*>(vol + o00 + c));
                    float4 f01 = __ldg(reinterpret_cast<const float4*>(vol + o01 + c));
                    float4 f10 = __ldg(reinterpret_cast<const float4*>(vol + o10 + c));
                    float4 f11 = __ldg(reinterpret_cast<const float4*>(vol + o11 + c));
                    const float4 gc = __ldg(reinterpret_cast<const float4*>(g + c));
#define D3F_BWD_LANE(m)                                                                                     \
                    {                                                                                       \
                        const float a = f00.m * m00, b = f01.m * m01, cc = f10.m * m10, d2 = f11.m * m11;   \
                        s_ix = fmaf(gc.m, (b - a) * ey + (d2 - cc) * wy, s_ix);                             \
                        s_iy = fmaf(gc.m, (cc - a) * ex + (d2 - b) * wx, s_iy);                             \
                        s_r = fmaf(gc.m, (a * ex + b * wx) * ey + (cc * ex + d2 * wx) * wy, s_r);           \
                    }
                    D3F_BWD_LANE(x) D3F_BWD_LANE(y) D3F_BWD_LANE(z) D3F_BWD_LANE(w)
#undef D3F_BWD_LANE
                }
            } else {
                for (int c = lane; c < C; c += 32) {
                    float f00, f01, f10, f11;
                    if (ks.dtype[k] == D3F_F32) {
                        const float* vol = static_cast<const float*>(ks.data[k]);
                        f00 = __ldg(vol + o00 + c); f01 = __ldg(vol + o01 + c); f10 = __ldg(vol + o10 + c); f11 = __ldg(vol + o11 + c);
                    } else {
                        const uint8_t* vol = static_cast<const uint8_t*>(ks.data[k]);
                        f00 = (float)__ldg(vol + o00 + c); f01 = (float)__ldg(vol + o01 + c);
                        f10 = (float)__ldg(vol + o10 + c); f11 = (float)__ldg(vol + o11 + c);
                    }
                    f00 *= m00; f01 *= m01; f10 *= m10; f11 *= m11;
                    const float gc = __ldg(g + c);
                    s_ix = fmaf(gc, (f01 - f00) * ey + (f11 - f10) * wy, s_ix);
                    s_iy = fmaf(gc, (f10 - f00) * ex + (f11 - f01) * wx, s_iy);
                    s_r = fmaf(gc, (f00 * ex + f01 * wx) * ey + (f10 * ex + f11 * wx) * wy, s_r);
                }
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) {
                s_ix += __shfl_xor_sync(0xffffffffu, s_ix, sh);
                s_iy += __shfl_xor_sync(0xffffffffu, s_iy, sh);
                s_r += __shfl_xor_sync(0xffffffffu, s_r, sh);
            }
            // d ix / d px = (w-1)/(W-1): x_norm = px/(W-1)*2-1, ix = (x_norm+1)*(w-1)/2
            G_px += fac * s_ix * ((float)(w - 1) / (float)(ep.W - 1));
            G_py += fac * s_iy * ((float)(h - 1) / (float)(ep.H - 1));
            G_w += inv * s_r;
        }
        // weight = exp(min(mu-|d|,0)/mu): d weight / d d = -sign(d) weight / mu where |d| >= mu (torch's clamp passes
        // the gradient at equality), else 0
        float G_d = (fabsf(d) >= ep.mu) ? G_w * weight * (d > 0.f ? -1.f : 1.f) / ep.mu : 0.f;
        // dist term: clamp passes the gradient where -mu <= d <= mu
        if (d >= -ep.mu && d <= ep.mu) G_d += gd * inv;
        // d = depth_nearest - cz ; px = cx/cz ; py = cy/cz
        const float G_cx = G_px / cz, G_cy = G_py / cz;
        const float G_cz = -G_d - (G_px * px + G_py * py) / cz;
        const float* Hm = sH + v * 12;
        return make_float3(Hm[0] * G_cx + Hm[4] * G_cy + Hm[8] * G_cz,
                           Hm[1] * G_cx + Hm[5] * G_cy + Hm[9] * G_cz,
                           Hm[2] * G_cx + Hm[6] * G_cy + Hm[10] * G_cz);
}

template <bool RECIP>
__global__ void __launch_bounds__(BWD_WARPS * 32)
field_backward_kernel(const EvalParams ep, const BwdKeySet ks, const float* __restrict__ grad_dist,
                      float* __restrict__ grad_pts) {
    __shared__ float sH[D3F_MAX_VIEWS * 12];
    __shared__ float s_view[BWD_WARPS][D3F_MAX_VIEWS][8];   // px, py, cz, d, weight, vis, (unused x2)
    const int V = ep.V;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = threadIdx.x; r < V * 3; r += BWD_WARPS * 32) {
        const int v = r / 3, i = r - v * 3;
        float row[4];
        krt_row(ep.K + v * 9, ep.pose + v * 12, i, row);
        sH[v * 12 + i * 4 + 0] = row[0]; sH[v * 12 + i * 4 + 1] = row[1];
        sH[v * 12 + i * 4 + 2] = row[2]; sH[v * 12 + i * 4 + 3] = row[3];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * BWD_WARPS + warp;
    if (i >= ep.n) return;
    const float x = __ldg(ep.pts + i * 3), y = __ldg(ep.pts + i * 3 + 1), z = __ldg(ep.pts + i * 3 + 2);

    // per-view forward, one lane per view
    bool vis = false;
    if (lane < V) {
        float Hm[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) Hm[j] = sH[lane * 12 + j];
        const ViewSample sm = view_sample<RECIP>(Hm, x, y, z, ep.depth + (size_t)lane * ep.H * ep.W, ep.H, ep.W, ep.mu, false);
        vis = sm.vis;
        float* sv = s_view[warp][lane];
        sv[0] = sm.px; sv[1] = sm.py;
        sv[2] = hdot(Hm + 8, x, y, z);           // camera z (a visible view never has the |z|<1e-4 patch)
        sv[3] = sm.d; sv[4] = sm.weight; sv[5] = vis ? 1.f : 0.f;
    }
    const unsigned vis_mask = __ballot_sync(0xffffffffu, vis);
    const float cnt = (float)__popc(vis_mask);
    __syncwarp();
    if (cnt == 0.f) {                            // no view sees the point: every output is a constant
        if (lane < 3) grad_pts[i * 3 + lane] = 0.f;
        return;
    }
    const float inv = __fdiv_rn(1.f, __fadd_rn(cnt, 1e-6f));
    const float gd = grad_dist ? __ldg(grad_dist + i) : 0.f;

    float gx = 0.f, gy = 0.f, gz = 0.f;          // accumulated by lane 0
    for (int v = 0; v < V; ++v) {
        if (!(vis_mask & (1u << v))) continue;
        const float3 c = view_gradient<RECIP>(ep, ks, i, v, s_view[warp][v], inv, gd, sH, lane);
        gx += c.x; gy += c.y; gz += c.z;
    }
    if (lane == 0) {
        grad_pts[i * 3 + 0] = gx; grad_pts[i * 3 + 1] = gy; grad_pts[i * 3 + 2] = gz;
    }
}

// The same gradient with one CTA per point and the views dealt to its warps: a tracking launch is a few hundred points
// (reference fusion.py:1650), and with one warp per point each warp walks its views — and their L2 round trips — one
// after the other.  Here the four (or more) views of a point are in flight at once; thread 0 adds the per-view
// contributions in view order, like the kernel above.  Used for launches of up to BWD_SPLIT_MAX points.
constexpr int BWD_SPLIT_MAX = 8192;

template <bool RECIP>
__global__ void __launch_bounds__(BWD_WARPS * 32)
field_backward_split_kernel(const EvalParams ep, const BwdKeySet ks, const float* __restrict__ grad_dist,
                            float* __restrict__ grad_pts) {
    __shared__ float sH[D3F_MAX_VIEWS * 12];
    __shared__ float s_view[D3F_MAX_VIEWS][8];
    __shared__ float s_c[D3F_MAX_VIEWS][3];
    const int V = ep.V;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = threadIdx.x; r < V * 3; r += BWD_WARPS * 32) {
        const int v = r / 3, k = r - v * 3;
        float row[4];
        krt_row(ep.K + v * 9, ep.pose + v * 12, k, row);
        sH[v * 12 + k * 4 + 0] = row[0]; sH[v * 12 + k * 4 + 1] = row[1];
        sH[v * 12 + k * 4 + 2] = row[2]; sH[v * 12 + k * 4 + 3] = row[3];
    }
    __syncthreads();
    const int64_t i = blockIdx.x;
    const float x = __ldg(ep.pts + i * 3), y = __ldg(ep.pts + i * 3 + 1), z = __ldg(ep.pts + i * 3 + 2);
    if (threadIdx.x < V) {                       // per-view forward, one thread per view
        const int v = threadIdx.x;
        float Hm[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) Hm[j] = sH[v * 12 + j];
        const ViewSample sm = view_sample<RECIP>(Hm, x, y, z, ep.depth + (size_t)v * ep.H * ep.W, ep.H, ep.W, ep.mu, false);
        float* sv = s_view[v];
        sv[0] = sm.px; sv[1] = sm.py;
        sv[2] = hdot(Hm + 8, x, y, z);
        sv[3] = sm.d; sv[4] = sm.weight; sv[5] = sm.vis ? 1.f : 0.f;
    }
    __syncthreads();
    float cnt = 0.f;
    for (int v = 0; v < V; ++v) cnt += s_view[v][5];
    if (cnt == 0.f) {                            // no view sees the point: every output is a constant
        if (threadIdx.x < 3) grad_pts[i * 3 + threadIdx.x] = 0.f;
        return;
    }
    const float inv = __fdiv_rn(1.f, __fadd_rn(cnt, 1e-6f));
    const float gd = grad_dist ? __ldg(grad_dist + i) : 0.f;
    for (int v = warp; v < V; v += BWD_WARPS) {
        float3 c = make_float3(0.f, 0.f, 0.f);
        if (s_view[v][5] != 0.f) c = view_gradient<RECIP>(ep, ks, i, v, s_view[v], inv, gd, sH, lane);
        if (lane == 0) { s_c[v][0] = c.x; s_c[v][1] = c.y; s_c[v][2] = c.z; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float g = 0.f;
        for (int v = 0; v < V; ++v)
            if (s_view[v][5] != 0.f) g += s_c[v][threadIdx.x];
        grad_pts[i * 3 + threadIdx.x] = g;
    }
}

}  // namespace d3f
