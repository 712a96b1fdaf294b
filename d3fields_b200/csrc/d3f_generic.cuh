// d3f_generic.cuh — the general fused field-query kernel: any V, any map size, any C, f32 or u8
// maps, optional per-view outputs.  One CTA evaluates a tile of TILE_PTS consecutive points:
//
//   phase 0  H = [K@Rt; 0 0 0 1] for every view -> shared memory
//   phase 1  one thread per (point, view): projection, nearest depth, visibility, distance
//            weight (reference fusion.py:323-347) -> shared memory
//   phase 1r one thread per point: reduce over views in order 0..V-1 -> dist, valid_mask
//            (fusion.py:358-370) and the per-view factor weight/(count+1e-6)
//   per key  one thread per (point, view): bilinear footprint on that key's map size;
//            then threads sweep (point, channel-group) pairs, channel group fastest, so a warp
//            reads contiguous channels of one texel and writes contiguous channels of one
//            output row (fusion.py:372-386)
//
// Nothing of size (V,n,C) is ever materialised unless return_inter asks for it, which is why
// batch_eval's chunking (fusion.py:526-545) disappears.
#pragma once
#include "d3f_common.cuh"

namespace d3f {

constexpr int GEN_THREADS = 256;
constexpr int GEN_TILE_PTS = 64;

struct KeySet {
    KeyParams k[D3F_MAX_KEYS];
    int32_t dtype[D3F_MAX_KEYS];
    int32_t n_keys;
};

// shared-memory carve-up (floats/ints, all 4-byte): per (point, view) slot
//   px, py, d, wt(fac), vis | fw0..fw3, off, dx, dy
__host__ __device__ inline size_t generic_smem_bytes(int V) {
    return (size_t)(V * 12 + GEN_TILE_PTS * V * 12) * 4;
}

template <typename T> struct Load4;
template <> struct Load4<float> {
    static __device__ __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
};
template <> struct Load4<uint8_t> {
    static __device__ __forceinline__ float4 ld(const uint8_t* p) {
        uchar4 u = __ldg(reinterpret_cast<const uchar4*>(p));
        return make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
    }
    static __device__ __forceinline__ float ld1(const uint8_t* p) { return (float)__ldg(p); }
};

__device__ __forceinline__ void fma4(float4& a, float w, const float4& f) {
    a.x = fmaf(w, f.x, a.x); a.y = fmaf(w, f.y, a.y); a.z = fmaf(w, f.z, a.z); a.w = fmaf(w, f.w, a.w);
}

// Sweep (point, channel-group) pairs of the tile for one key.
template <typename T, int VEC, bool INTER>
__device__ __forceinline__ void generic_accumulate(const KeyParams& kp, int V, int64_t n, int64_t tile0, int npts,
                                                   const float* s_fac, const float* s_fw,
                                                   const int* s_off, const int* s_dx, const int* s_dy,
                                                   const int32_t* __restrict__ order) {
    const int C = kp.C;
    const int G = C / VEC;
    const T* __restrict__ vol = static_cast<const T*>(kp.data);
    const size_t view_stride = (size_t)kp.sv;
    for (int item = threadIdx.x; item < npts * G; item += GEN_THREADS) {
        const int p = item / G;
        const int c = (item - p * G) * VEC;
        const int64_t row = order ? (int64_t)__ldg(order + tile0 + p) : tile0 + p;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int v = 0; v < V; ++v) {
            const int s = p * V + v;
            const float fac = s_fac[s];
            if (!INTER && fac == 0.f) continue;          // invisible view: contributes exactly 0 (finite maps)
            const T* base = vol + v * view_stride + (size_t)s_off[s] + c;
            const int dx = s_dx[s], dy = s_dy[s];
            const float w0 = s_fw[s * 4 + 0], w1 = s_fw[s * 4 + 1], w2 = s_fw[s * 4 + 2], w3 = s_fw[s * 4 + 3];
            if (VEC == 4) {
                float4 f0 = Load4<T>::ld(base), f1 = Load4<T>::ld(base + dx);
                float4 f2 = Load4<T>::ld(base + dy), f3 = Load4<T>::ld(base + dy + dx);
                if (INTER) {
                    // per-view sample in the reference's order nw, ne, sw, se (fusion.py:373-379)
                    float4 r;
                    r.x = f0.x * w0 + f1.x * w1 + f2.x * w2 + f3.x * w3;
                    r.y = f0.y * w0 + f1.y * w1 + f2.y * w2 + f3.y * w3;
                    r.z = f0.z * w0 + f1.z * w1 + f2.z * w2 + f3.z * w3;
                    r.w = f0.w * w0 + f1.w * w1 + f2.w * w2 + f3.w * w3;
                    if (kp.inter) __stcs(reinterpret_cast<float4*>(kp.inter + ((size_t)v * n + row) * C + c), r);
                    fma4(acc, fac, r);
                } else {
                    fma4(acc, w0 * fac, f0); fma4(acc, w1 * fac, f1);
                    fma4(acc, w2 * fac, f2); fma4(acc, w3 * fac, f3);
                }
            } else {
                float f0 = Load4<T>::ld1(base), f1 = Load4<T>::ld1(base + dx);
                float f2 = Load4<T>::ld1(base + dy), f3 = Load4<T>::ld1(base + dy + dx);
                float r = f0 * w0 + f1 * w1 + f2 * w2 + f3 * w3;
                if (INTER && kp.inter) __stcs(kp.inter + ((size_t)v * n + row) * C + c, r);
                acc.x = fmaf(fac, r, acc.x);
            }
        }
        if (kp.bias) {                                   // affine epilogue: out = field - bias
            if (VEC == 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(kp.bias + c));
                acc.x -= b.x; acc.y -= b.y; acc.z -= b.z; acc.w -= b.w;
            } else {
                acc.x -= __ldg(kp.bias + c);
            }
        }
        float* o = kp.out + (size_t)row * C + c;
        if (VEC == 4) __stcs(reinterpret_cast<float4*>(o), acc);
        else          __stcs(o, acc.x);
    }
}

template <bool INTER, bool RECIP>
__global__ void __launch_bounds__(GEN_THREADS)
field_generic_kernel(const EvalParams ep, const KeySet ks) {
    extern __shared__ __align__(16) float smem[];
    const int V = ep.V;
    float* sH    = smem;                                   // V*12
    float* s_px  = sH + V * 12;                            // TILE*V each from here on
    float* s_py  = s_px + GEN_TILE_PTS * V;
    float* s_d   = s_py + GEN_TILE_PTS * V;
    float* s_fac = s_d + GEN_TILE_PTS * V;
    int*   s_vis = reinterpret_cast<int*>(s_fac + GEN_TILE_PTS * V);
    float* s_fw  = reinterpret_cast<float*>(s_vis + GEN_TILE_PTS * V);   // 4 per slot
    int*   s_off = reinterpret_cast<int*>(s_fw + GEN_TILE_PTS * V * 4);
    int*   s_dx  = s_off + GEN_TILE_PTS * V;
    int*   s_dy  = s_dx + GEN_TILE_PTS * V;

    const bool eval_dist = (ep.flags & D3F_FLAG_EVAL_DIST) != 0;
    const int64_t tile0 = (int64_t)blockIdx.x * GEN_TILE_PTS;
    const int npts = (int)min((int64_t)GEN_TILE_PTS, ep.n - tile0);

    // phase 0
    for (int r = threadIdx.x; r < V * 3; r += GEN_THREADS) {
        const int v = r / 3, i = r - v * 3;
        float row[4];
        krt_row(ep.K + v * 9, ep.pose + v * 12, i, row);
        sH[v * 12 + i * 4 + 0] = row[0]; sH[v * 12 + i * 4 + 1] = row[1];
        sH[v * 12 + i * 4 + 2] = row[2]; sH[v * 12 + i * 4 + 3] = row[3];
    }
    __syncthreads();

    // phase 1: slot s = p*V + v, threads view-major so a warp walks consecutive points of one view
    for (int item = threadIdx.x; item < npts * V; item += GEN_THREADS) {
        const int v = item / npts, p = item - v * npts;
        const int64_t row = ep.order ? (int64_t)__ldg(ep.order + tile0 + p) : tile0 + p;
        const float* q = ep.pts + (size_t)row * 3;
        const float x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
        float Hm[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) Hm[j] = sH[v * 12 + j];
        ViewSample sm = view_sample<RECIP>(Hm, x, y, z, ep.depth + (size_t)v * ep.H * ep.W, ep.H, ep.W, ep.mu, eval_dist);
        const int s = p * V + v;
        s_px[s] = sm.px; s_py[s] = sm.py;
        s_d[s] = eval_dist ? sm.d : fminf(fmaxf(sm.d, -ep.mu), ep.mu);      // fusion.py:358
        s_fac[s] = sm.weight;
        s_vis[s] = sm.vis ? 1 : 0;
    }
    __syncthreads();

    // phase 1r: views summed in order (fusion.py:364-370)
    if (threadIdx.x < npts) {
        const int p = threadIdx.x;
        float acc = 0.f, cnt = 0.f;
        for (int v = 0; v < V; ++v) {
            if (s_vis[p * V + v]) { acc = __fadd_rn(acc, s_d[p * V + v]); cnt = __fadd_rn(cnt, 1.f); }
        }
        const float denom = __fadd_rn(cnt, 1e-6f);
        float dist = __fdiv_rn(acc, denom);
        if (!eval_dist && cnt == 0.f) dist = 1e3f;                            // fusion.py:367
        store_compact(ep, ep.order ? (int64_t)__ldg(ep.order + tile0 + p) : tile0 + p, dist, cnt != 0.f ? 1 : 0);
        const float inv = __fdiv_rn(1.f, denom);
        for (int v = 0; v < V; ++v) {
            const int s = p * V + v;
            s_fac[s] = s_vis[s] ? __fmul_rn(s_fac[s], inv) : 0.f;             // weight/(count+1e-6), fusion.py:385
        }
    }
    if (eval_dist || ks.n_keys == 0) {
        gather_epilogue(ep);
        return;
    }
    __syncthreads();

    for (int k = 0; k < ks.n_keys; ++k) {
        const KeyParams& kp = ks.k[k];
        for (int s = threadIdx.x; s < npts * V; s += GEN_THREADS) {
            Footprint f = footprint<RECIP>(s_px[s], s_py[s], ep.H, ep.W, kp.h, kp.w);
            s_fw[s * 4 + 0] = f.w[0]; s_fw[s * 4 + 1] = f.w[1]; s_fw[s * 4 + 2] = f.w[2]; s_fw[s * 4 + 3] = f.w[3];
            s_off[s] = f.y0 * kp.sy + f.x0 * kp.sx;
            s_dx[s] = f.dx ? kp.sx : 0; s_dy[s] = f.dy ? kp.sy : 0;
        }
        __syncthreads();
        const bool vec4 = (kp.C % 4 == 0) && ((kp.sv | kp.sy | kp.sx) & 3) == 0;
        if (ks.dtype[k] == D3F_F32) {
            if (vec4) generic_accumulate<float, 4, INTER>(kp, V, ep.n, tile0, npts, s_fac, s_fw, s_off, s_dx, s_dy, ep.order);
            else      generic_accumulate<float, 1, INTER>(kp, V, ep.n, tile0, npts, s_fac, s_fw, s_off, s_dx, s_dy, ep.order);
        } else {
            if (vec4) generic_accumulate<uint8_t, 4, INTER>(kp, V, ep.n, tile0, npts, s_fac, s_fw, s_off, s_dx, s_dy, ep.order);
            else      generic_accumulate<uint8_t, 1, INTER>(kp, V, ep.n, tile0, npts, s_fac, s_fw, s_off, s_dx, s_dy, ep.order);
        }
        __syncthreads();
    }
    gather_epilogue(ep);
}

}  // namespace d3f
