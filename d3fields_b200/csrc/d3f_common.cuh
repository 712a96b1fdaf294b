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

// In-kernel all-gather of the compact field (d3f_eval_allgather): where every peer's gathered arrays and
// epoch flags live.  world == 0 means a single-GPU launch (dist / valid go to EvalParams::dist / valid).
struct GatherParams {
    float* dist[D3F_MAX_PEERS];          // gathered dist array in rank r's segment (r == own rank: local memory)
    uint8_t* valid[D3F_MAX_PEERS];       // gathered valid_mask array in rank r's segment
    uint32_t* flag[D3F_MAX_PEERS];       // &epoch_flags[own rank] in rank r's segment
    const uint32_t* my_flags;            // epoch_flags[world] of the own segment, written by the peers
    uint32_t* counter;                   // CTAs of this launch that have finished (own segment)
    uint32_t* error;                     // set to 1 when a wait times out (own segment)
    int64_t base, block, stride;         // local point i -> gathered index base + (i / block) * stride + i % block
    uint32_t epoch;
    int32_t world;
};

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
    const int32_t* __restrict__ order;  // nullptr, or (n): step i evaluates point order[i] and writes row order[i]
    GatherParams g;
};

struct KeyParams {
    const void* __restrict__ data;      // (V,h,w,C), channel stride 1
    float* __restrict__ out;            // (n,C)
    float* __restrict__ inter;          // (V,n,C) or nullptr
    const float* __restrict__ bias;     // nullptr, or (C): subtracted from every output row (narrow keys only)
    int64_t sv;                         // element stride between views
    int32_t sy, sx;                     // element strides between rows / texels (a view spans < 2^31 elements)
    int32_t h, w, C;
};

// ---- system-scope flag traffic for the in-kernel gather -------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long COMM_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // a peer that is 20 s late is gone

// Spin until *flag has reached `epoch` (epochs only grow; a peer may already be one ahead).  Returns false on timeout.
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch) {
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        if (globaltimer_ns() - t0 > COMM_TIMEOUT_NS) return false;
        __nanosleep(200);
    }
    return true;
}

// dist / valid_mask of point i of this launch: to the caller's arrays, or — gathering — into the gathered arrays of
// every rank (remote stores over NVLink; 5 B/point/peer).
__device__ __forceinline__ void store_compact(const EvalParams& ep, int64_t i, float dist, uint8_t valid) {
    if (ep.g.world == 0) {
        ep.dist[i] = dist;
        ep.valid[i] = valid;
        return;
    }
    const int64_t q = i / ep.g.block;
    const int64_t gi = ep.g.base + q * ep.g.stride + (i - q * ep.g.block);
    for (int r = 0; r < ep.g.world; ++r) {
        ep.g.dist[r][gi] = dist;
        ep.g.valid[r][gi] = valid;
    }
}

// End of a gathering launch, called by every thread of every CTA.
//
// Ordering (PTX memory model, fences are cumulative): a CTA's remote stores are ordered before its arrival on the
// launch counter by bar.sync + a GPU-scope fence of thread 0 + a GPU-scope atomic; the last CTA to arrive has thereby
// observed every CTA's stores, and ITS system-scope fence followed by release stores of the epoch flags orders all of
// them before the flags for the peers.  Only one system-scope fence per launch: a fence.sys costs a round trip through
// the NVLink fabric (~4 us measured), and one per CTA cost 0.10-0.13 ms per step on 2M points (8 % — profiles/
// r02_bench_n8_fence_sys_per_cta.json); the GPU-scope fence is satisfied at the local L2.
// The last CTA then waits until every peer has published the same epoch, so when the launch completes the local
// gathered arrays hold all ranks' results.
__device__ __forceinline__ void gather_epilogue(const EvalParams& ep) {
    if (ep.g.world == 0) return;
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    const unsigned arrived = atomicAdd(ep.g.counter, 1u);
    if (arrived != gridDim.x - 1) return;
    *ep.g.counter = 0;                               // the next launch on this stream starts from zero
    __threadfence_system();
    for (int r = 0; r < ep.g.world; ++r) st_release_sys(ep.g.flag[r], ep.g.epoch);
    for (int r = 0; r < ep.g.world; ++r)
        if (!wait_flag(ep.g.my_flags + r, ep.g.epoch)) { *ep.g.error = 1u; return; }
}

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
    // the two halvings are multiplications by 0.5f: exact, like the division by 2 they stand for, without its ~10
    // instructions (a quarter of the divisions of a point-view; the sweep and dist-only kernels are issue-bound on them)
    float n;
    if (RECIP) n = __fmul_rn(p, __fdiv_rn(1.f, (float)(img_size - 1)));
    else       n = __fdiv_rn(p, (float)(img_size - 1));
    n = __fsub_rn(__fmul_rn(n, 2.f), 1.f);
    if (RECIP) return __fmul_rn(__fmul_rn(__fadd_rn(n, 1.f), 0.5f), (float)(map_size - 1));
    return __fmul_rn(__fadd_rn(n, 1.f), __fmul_rn((float)(map_size - 1), 0.5f));
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
                                                  float mu, bool eval_dist, bool need_weight = true) {
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
        s.weight = 1.f;
        if (need_weight) {                          // launches without keys (dist / valid_mask only) never read it
            float a = fminf(__fsub_rn(mu, fabsf(s.d)), 0.f);
            s.weight = expf(__fdiv_rn(a, mu));      // fusion.py:347
        }
    }
    return s;
}

// Bilinear footprint of one (point, view) on an (h,w) map: grid_sample 'bilinear', zero padding,
// align_corners=True.  Corner order nw, ne, sw, se.  Out-of-range corners get weight 0 and a
// clamped (always loadable) address.
struct Footprint {
    float w[4];       // corner weights
    int32_t x0, y0;   // (clamped) north-west corner texel
    int32_t dx, dy;   // 1 when the east / south corner is a different texel after clamping, else 0
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
    f.x0 = x0c; f.y0 = y0c;
    f.dx = x1c - x0c;
    f.dy = y1c - y0c;
    return f;
}

}  // namespace d3f
