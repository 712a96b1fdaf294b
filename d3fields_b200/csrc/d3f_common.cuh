// d3f_common.cuh — shared device helpers for the field-query kernels (sm_100a).
//
// The index arithmetic below decides integer results (which depth pixel a point hits, which
// four texels a bilinear sample reads, whether a view sees the point), so it replays the
// reference's float32 operation sequence exactly: one IEEE rounding per operation, no fused
// multiply-add (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn are never contracted by nvcc).
// Reference: fusion.py:32-55 (project_points_coords), fusion.py:57-77 (interpolate_feats),
// torch CPU grid_sample un-normalisation (align_corners=True): ix = (x_norm + 1) * ((size-1)/2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/d3f.h"

namespace d3f {

// Per-launch constants, passed by value (lives in the kernel parameter / constant bank).
struct EvalParams {
    const float* __restrict__ pts;      // (n,3)
    const float* __restrict__ depth;    // (V,H,W)
    const float* __restrict__ pose;     // (V,3,4)
    const float* __restrict__ K;        // (V,3,3)
    float* __restrict__ dist;           // (n)
    uint8_t* __restrict__ valid;        // (n)
    int64_t n;
    int32_t V, H, W;
    float mu;
    uint32_t flags;
};

struct KeyParams {
    const void* __restrict__ data;      // (V,h,w,C)
    float* __restrict__ out;            // (n,C)
    float* __restrict__ inter;          // (V,n,C) or nullptr
    const float* __restrict__ bias;     // nullptr, or (C): subtracted from every output row (narrow keys only)
    int32_t h, w, C;
};

// Row i of H = [K@Rt ; 0 0 0 1] for one view (reference fusion.py:45-48): the small-matrix
// product accumulates k = 0,1,2 sequentially from 0 with separate multiply and add.
__device__ __forceinline__ void krt_row(const float* __restrict__ K, const float* __restrict__ Rt,
                                        int i, float r[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) a = __fadd_rn(a, __fmul_rn(K[i * 3 + k], Rt[k * 4 + j]));
        r[j] = a;
    }
}

// pts_cam component = H[i,:] . [x y z 1] (reference fusion.py:49): k = 0..3 sequential from 0.
__device__ __forceinline__ float hdot(const float r[4], float x, float y, float z) {
    float a = __fadd_rn(0.f, __fmul_rn(r[0], x));
    a = __fadd_rn(a, __fmul_rn(r[1], y));
    a = __fadd_rn(a, __fmul_rn(r[2], z));
    a = __fadd_rn(a, __fmul_rn(r[3], 1.f));
    return a;
}

// Pixel coordinate -> continuous index on a map of `map_size` texels when the image is
// `img_size` pixels (reference fusion.py:72-73 then grid_sample's un-normalise).
//   CPU rounding (default):  ((p / (img-1)) * 2 - 1 + 1) * ((map-1)/2)
//   D3F_FLAG_RECIP_NORM:     ((p * (1/(img-1))) * 2 - 1 + 1) / 2 * (map-1)     (torch CUDA kernels)
template <bool RECIP>
__device__ __forceinline__ float to_map_index(float p, int img_size, int map_size) {
    float n;
    if (RECIP) n = __fmul_rn(p, __fdiv_rn(1.f, (float)(img_size - 1)));
    else       n = __fdiv_rn(p, (float)(img_size - 1));
    n = __fsub_rn(__fmul_rn(n, 2.f), 1.f);
    if (RECIP) return __fmul_rn(__fdiv_rn(__fadd_rn(n, 1.f), 2.f), (float)(map_size - 1));
    return __fmul_rn(__fadd_rn(n, 1.f), __fdiv_rn((float)(map_size - 1), 2.f));
}

// One view of one point: everything Fusion.eval derives before it touches a feature map
// (reference fusion.py:323-347).  Returns the pixel coords, the raw signed distance, the
// visibility bit and the distance weight.
struct ViewSample {
    float px, py;     // pixel coordinates (pts_2d)
    float d;          // inter_depth - z, unclamped
    float weight;     // exp(min(mu - |d|, 0) / mu)
    bool vis;         // dist_valid
};

template <bool RECIP>
__device__ __forceinline__ ViewSample view_sample(const float Hm[12], float x, float y, float z,
                                                  const float* __restrict__ depth_v, int H, int W,
                                                  float mu, bool eval_dist) {
    ViewSample s;
    float cx = hdot(Hm + 0, x, y, z);
    float cy = hdot(Hm + 4, x, y, z);
    float cz = hdot(Hm + 8, x, y, z);
    bool ok = !(fabsf(cz) < 1e-4f);                 // fusion.py:52
    if (!ok) cz = 1e-3f;                            // fusion.py:53
    s.px = __fdiv_rn(cx, cz);                       // fusion.py:54
    s.py = __fdiv_rn(cy, cz);
    // nearest-neighbour depth, zero padding (fusion.py:327-333): round-half-to-even, value 0 outside
    float xr = rintf(to_map_index<RECIP>(s.px, W, W));
    float yr = rintf(to_map_index<RECIP>(s.py, H, H));
    float dep = 0.f;
    if (xr >= 0.f && xr <= (float)(W - 1) && yr >= 0.f && yr <= (float)(H - 1))
        dep = __ldg(depth_v + (size_t)(int)yr * W + (int)xr);
    s.d = __fsub_rn(dep, cz);                       // fusion.py:343
    if (eval_dist) {
        s.vis = (dep > 0.f) && ok;                  // fusion.py:423
        s.weight = 1.f;
    } else {
        s.vis = (dep > 0.f) && ok && (s.d > -mu);   // fusion.py:344
        float a = fminf(__fsub_rn(mu, fabsf(s.d)), 0.f);
        s.weight = expf(__fdiv_rn(a, mu));          // fusion.py:347
    }
    return s;
}

// Bilinear footprint of one (point, view) on an (h,w) map: grid_sample 'bilinear', zero padding,
// align_corners=True.  Corner order nw, ne, sw, se.  Out-of-range corners get weight 0 and a
// clamped (always loadable) address.
struct Footprint {
    float w[4];       // corner weights
    int32_t off;      // texel offset of the (clamped) north-west corner: y0c*w + x0c
    int32_t dx, dy;   // texel steps to the east / south corners after clamping (0 or 1 / 0 or w)
};

template <bool RECIP>
__device__ __forceinline__ Footprint footprint(float px, float py, int H, int W, int h, int w) {
    Footprint f;
    float ix = to_map_index<RECIP>(px, W, w);
    float iy = to_map_index<RECIP>(py, H, h);
    float x0 = floorf(ix), y0 = floorf(iy);
    float wx = __fsub_rn(ix, x0), ex = __fsub_rn(1.f, wx);
    float wy = __fsub_rn(iy, y0), sy = __fsub_rn(1.f, wy);
    float x1 = x0 + 1.f, y1 = y0 + 1.f;
    const float xm = (float)(w - 1), ym = (float)(h - 1);
    bool x0ok = (x0 >= 0.f) && (x0 <= xm), x1ok = (x1 >= 0.f) && (x1 <= xm);
    bool y0ok = (y0 >= 0.f) && (y0 <= ym), y1ok = (y1 >= 0.f) && (y1 <= ym);
    f.w[0] = (x0ok && y0ok) ? __fmul_rn(sy, ex) : 0.f;
    f.w[1] = (x1ok && y0ok) ? __fmul_rn(sy, wx) : 0.f;
    f.w[2] = (x0ok && y1ok) ? __fmul_rn(wy, ex) : 0.f;
    f.w[3] = (x1ok && y1ok) ? __fmul_rn(wy, wx) : 0.f;
    // clamp in float first: NaN and |x| >= 2^31 must not reach the int conversion
    int x0c = (int)fminf(fmaxf(x0, 0.f), xm), x1c = (int)fminf(fmaxf(x1, 0.f), xm);
    int y0c = (int)fminf(fmaxf(y0, 0.f), ym), y1c = (int)fminf(fmaxf(y1, 0.f), ym);
    f.off = y0c * w + x0c;
    f.dx = x1c - x0c;
    f.dy = (y1c - y0c) * w;
    return f;
}

}  // namespace d3f
