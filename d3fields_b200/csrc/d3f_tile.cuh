// d3f_tile.cuh — the production field-query kernel (V <= 8, no per-view outputs).
//
// One CTA (256 threads) evaluates a tile of consecutive query points (256 for up to 4 views, 128 for 5-8 views):
//
//   phase 0/1/1r  as in d3f_generic.cuh: H = [K@Rt;0001], one thread per (point, view) for
//                 projection / nearest depth / visibility / distance weight, then one thread per
//                 point reduces the views in order -> dist, valid_mask, per-view factor
//   per key       one thread per (point, view) turns the pixel coordinate into a bilinear footprint
//                 on that key's map: four corner weights already multiplied by the view factor
//                 weight/(count+1e-6), and one packed code (north-west texel offset, east/south
//                 step bits; -1 when the view does not see the point)
//   wide maps     (float32, C a multiple of the slice width, e.g. the 1024-channel DINOv2 volume): a warp owns a
//                 channel slice (128 channels = one float4 per lane for NV = 4; 64 channels = one float2 per lane for
//                 NV = 8, so that the cache below stays 64 registers) and marches over the tile's points in
//                 order.  The four corner texels of each view stay in registers and are reloaded
//                 only when the packed code changes — neighbouring grid points project into the
//                 same texel cell most of the time, so the 16 KB-per-point gather of the reference
//                 (4 views x 4 corners x C floats) collapses to a reload every few points, served
//                 by L1/L2 (the whole volume is L2-resident).  Per visible view the inner loop is
//                 one 128-bit shared-memory read of the folded weights and 4 FFMA per channel; the
//                 output row is written once with streaming stores (st.global.cs), a contiguous slice per warp,
//                 so the 4 KB/point output stream never evicts the feature volume from L2.
//   narrow maps   (instance masks, colours; u8 or f32, any C): threads sweep (point, channel-group).
//
// HBM traffic is the algorithmic minimum: points and depth pixels in, each output byte out once; the
// feature volume is read from HBM once and then lives in L2.
//
// Experiments on this walk that were measured and rejected (packed FFMA2 accumulation, strided tile order, switch
// dispatch, unconditional weight reads, TMEM-resident cache, TMA stores, ...) are listed with their numbers in DESIGN.md
// 4.7 / 4.8; their sources are under profiles/patches/.
#pragma once
#include "d3f_common.cuh"
#include "d3f_generic.cuh"

namespace d3f {

constexpr int TILE_THREADS = 256;
constexpr int TILE_MAX_V = 8;            // views the tile kernel handles (more: field_generic_kernel)

// Geometry of a tile for NV view slots (4 or 8; a launch with V views uses the smallest NV >= V).
// NV = 0 is the latency geometry: 4 view slots and 32-point tiles.  A warp walks its tile's points one after the other,
// so a launch of a few hundred points (rigid_tracking evaluates num_inst x 100, fusion.py:1650; select_features_* a few
// hundred samples, fusion.py:1449) would put ~10 us of serial walk on two CTAs with 256-point tiles; 32-point tiles
// spread it over 8x as many CTAs.
// NVX = 1: the same with 8-point tiles, for launches of at most a few thousand points (one Adam iteration of tracking).
template <int NVX>
struct TileGeom {
    static_assert(NVX == 0 || NVX == 1 || NVX == 4 || NVX == 8, "view slots per point: 4 or 8 (0 / 1: 4 slots, small / tiny tiles)");
    static constexpr int NV = NVX <= 1 ? 4 : NVX;
    static constexpr int PTS = NVX == 1 ? 8 : NVX == 0 ? 32 : (NV == 4 ? 256 : 128);   // 256-point tiles beat 128 by 4 % at NV = 4 (fewer barriers and cold starts)
    static constexpr int LOG_NV = NV == 4 ? 2 : 3;
    static constexpr int LANE_CH = NV == 4 ? 4 : 2;      // channels per lane: the register cache is NV x 4 corners x LANE_CH = 64
    static constexpr int SLICE = 32 * LANE_CH;           // channels a warp owns
    static constexpr int CHG = NV;                       // mask bits [0,NV): view sees the point; [NV,2NV): its cell changed
    static constexpr int AHEAD = NV == 4 ? 12 : 16;      // [AHEAD, AHEAD+NV): cell change WIDE_LOOKAHEAD points further on
    static constexpr unsigned VIEWS = (1u << NV) - 1u;
};
constexpr int TILE_PTS = TileGeom<4>::PTS;

template <bool WIDE, int NVX>
struct TileSmemT {
    static constexpr int PTS = TileGeom<NVX>::PTS;
    static constexpr int NV = TileGeom<NVX>::NV;
    float H[NV * 12];
    float px[PTS * NV];
    float py[PTS * NV];
    float d[PTS * NV];
    float fac[PTS * NV];
    int vis[PTS * NV];
    float4 w4[PTS * NV];              // folded corner weights per (point, view): (w_nw, w_ne, w_sw, w_se)
    int code[(PTS + 4) * NV];         // packed footprint codes of the NV views of a point: element offset of the
                                      // north-west texel inside the view (wide maps: a multiple of 4, used as is;
                                      // narrow maps: shifted left by 2) | east-step bit | south-step bit << 1; -1 = unseen
    int64_t row[PTS];                 // output row of each point of the tile (ordered launches only)
    int mask[PTS + 4];                // see TileGeom: visibility bits, cell-change bits, lookahead bits
};

// The lane's slice of one texel / one output row: float4 (NV = 4) or float2 (NV = 8).
template <int LANE_CH> struct LaneVec;
template <> struct LaneVec<4> {
    using type = float4;
    static __device__ __forceinline__ type zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    // Texel loads: read-only path, kept in L1 with evict-last priority — the next cell a walk enters shares two of its
    // four corners with the current one, and the output stream must not push them out (measured -1 %).
    static __device__ __forceinline__ type ldg(const float* p) {
        float4 r;
        asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ void fma(type& a, float w, const type& f) {
        a.x = fmaf(w, f.x, a.x); a.y = fmaf(w, f.y, a.y); a.z = fmaf(w, f.z, a.z); a.w = fmaf(w, f.w, a.w);
    }
    static __device__ __forceinline__ void stcs(float* p, const type& v) { __stcs(reinterpret_cast<float4*>(p), v); }
};
template <> struct LaneVec<2> {
    using type = float2;
    static __device__ __forceinline__ type zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ type ldg(const float* p) {
        float2 r;
        asm volatile("ld.global.nc.L1::evict_last.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ void fma(type& a, float w, const type& f) {
        a.x = fmaf(w, f.x, a.x); a.y = fmaf(w, f.y, a.y);
    }
    static __device__ __forceinline__ void stcs(float* p, const type& v) { __stcs(reinterpret_cast<float2*>(p), v); }
};

// Shared-memory reads of the walk go through explicit 32-bit shared addresses that are computed ONCE per walk and
// pinned in a register (the opaque mov): left to itself the compiler re-derives the CTA's shared-window base
// (S2UR SR_CgaCtaId + ULEA) in front of every point's weight reads, which puts a special-register read at the head of
// each point's dependency chain (measured: cfg2a 0.781 -> 0.760 ms, fully visible grid 1.075 -> 1.022 ms).
__device__ __forceinline__ unsigned smem_addr(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 r;
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ int4 lds_i4(unsigned a) {
    int4 r;
    asm("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ int lds_i(unsigned a) {
    int r;
    asm("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}

// The NV packed codes of point p (shared address a_code), as registers.
template <int NV>
__device__ __forceinline__ void load_codes(unsigned a_code, int p, int (&cv)[NV]) {
    const int4 c0 = lds_i4(a_code + p * (NV * 4));
    cv[0] = c0.x; cv[1] = c0.y; cv[2] = c0.z; cv[3] = c0.w;
    if constexpr (NV == 8) {
        const int4 c1 = lds_i4(a_code + p * (NV * 4) + 16);
        cv[4] = c1.x; cv[5] = c1.y; cv[6] = c1.z; cv[7] = c1.w;
    }
}

// Number of point runs a tile is split into for a map of S channel slices (8 warps per CTA).
__host__ __device__ inline int wide_runs(int S) { return S >= TILE_THREADS / 32 ? 1 : (TILE_THREADS / 32) / S; }
__host__ __device__ inline int wide_run_len(int S, int pts) { const int R = wide_runs(S); return (pts + R - 1) / R; }

constexpr int WIDE_LOOKAHEAD = 4;      // points between the L1 prefetch of a cell change and its reload

// PREFETCH: the lookahead bits of a point's mask word say which views change cell WIDE_LOOKAHEAD points later; the warp
// then prefetches its own slice of those corner texels into L1, so the reload that follows is an L1 hit
// instead of an L2 round trip with all eight warps of the CTA stalled on the same point.
template <bool PREFETCH, bool ORDERED, int NVX>
__device__ __forceinline__ void wide_accumulate(const KeyParams& kp, int64_t tile0, int npts, const TileSmemT<true, NVX>& sm) {
    using G = TileGeom<NVX>;
    constexpr int NV = G::NV;
    using LV = LaneVec<G::LANE_CH>;
    using vec = typename LV::type;
    const int C = kp.C;
    const int S = C / G::SLICE;                             // channel slices
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = TILE_THREADS / 32;
    int s0, sstep, p_begin, p_end;
    if (S >= nwarps) {                                      // every warp walks the whole tile on its slices
        s0 = warp; sstep = nwarps; p_begin = 0; p_end = npts;
    } else {                                                // fewer slices than warps: split the tile into runs
        const int r = warp / S;
        if (r >= wide_runs(S)) return;
        const int run = wide_run_len(S, G::PTS);
        s0 = warp - r * S; sstep = S;
        p_begin = r * run; p_end = min(p_begin + run, npts);
    }
    if (p_end <= p_begin) return;
    const float* __restrict__ vol = static_cast<const float*>(kp.data);
    const unsigned a_mask = smem_addr(sm.mask), a_code = smem_addr(sm.code), a_w4 = smem_addr(sm.w4);
    for (int s = s0; s < S; s += sstep) {
        const float* vbase[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) vbase[v] = vol + (size_t)v * (size_t)kp.sv + s * G::SLICE + lane * G::LANE_CH;
        float* o = kp.out + (size_t)(tile0 + p_begin) * C + s * G::SLICE + lane * G::LANE_CH;
        float* const o_col = kp.out + s * G::SLICE + lane * G::LANE_CH;
        vec cc[NV][4];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int q = 0; q < 4; ++q) cc[v][q] = LV::zero();
        int m_next = lds_i(a_mask + p_begin * 4);
        for (int p = p_begin; p < p_end; ++p, o += C) {
            // the mask is the same in every lane; the OR-reduction moves it to a uniform register so the
            // tests below are uniform branches (no divergence bookkeeping)
            const unsigned m = __reduce_or_sync(0xffffffffu, (unsigned)m_next);
            m_next = lds_i(a_mask + (p + 1) * 4);                     // the array is padded by one
            if (PREFETCH && ((m >> G::AHEAD) & G::VIEWS)) {
                int cv[NV];
                load_codes<NV>(a_code, p + WIDE_LOOKAHEAD, cv);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (m & (1u << (G::AHEAD + v))) {
                        const float* b_ = vbase[v] + (unsigned)(cv[v] & ~3);
                        const unsigned dx_ = (cv[v] & 1) ? (unsigned)kp.sx : 0u;
                        const unsigned dy_ = (cv[v] & 2) ? (unsigned)kp.sy : 0u;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_ + dx_));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_ + dy_));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_ + (dy_ + dx_)));
                    }
                }
            }
            vec acc = LV::zero();
            if (m & G::VIEWS) {
                if ((m >> G::CHG) & G::VIEWS) {
                    // reload the four corner texels of every view whose cell changed (32-bit element offsets: a view's
                    // map spans fewer than 2^31 elements, checked on the host)
                    int cv[NV];
                    load_codes<NV>(a_code, p, cv);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        if (m & (1u << (G::CHG + v))) {
                            const float* b_ = vbase[v] + (unsigned)(cv[v] & ~3);
                            const unsigned dx_ = (cv[v] & 1) ? (unsigned)kp.sx : 0u;
                            const unsigned dy_ = (cv[v] & 2) ? (unsigned)kp.sy : 0u;
                            cc[v][0] = LV::ldg(b_); cc[v][1] = LV::ldg(b_ + dx_);
                            cc[v][2] = LV::ldg(b_ + dy_); cc[v][3] = LV::ldg(b_ + (dy_ + dx_));
                        }
                    }
                }
                // the weight reads of every visible view are issued before the first FMA.  (Dispatching on the
                // visibility bits with a switch instead of this if-chain compiles to a compare tree, not an
                // indexed branch, and measured 3 % slower: profiles/r02_experiment_switch_dispatch.jsonl.)
                float4 w[NV];
                const unsigned aw = a_w4 + p * (NV * 16);
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (m & (1u << v)) w[v] = lds_f4(aw + 16 * v);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (m & (1u << v)) {                                    // corner order nw, ne, sw, se; views in order
                        LV::fma(acc, w[v].x, cc[v][0]); LV::fma(acc, w[v].y, cc[v][1]);
                        LV::fma(acc, w[v].z, cc[v][2]); LV::fma(acc, w[v].w, cc[v][3]);
                    }
                }
            }
            if (ORDERED) LV::stcs(o_col + (size_t)sm.row[p] * C, acc);
            else         LV::stcs(o, acc);
        }
    }
}

template <typename T, int VEC, bool ORDERED, bool WIDE, int NVX>
__device__ __forceinline__ void narrow_accumulate(const KeyParams& kp, int64_t tile0, int npts, const TileSmemT<WIDE, NVX>& sm) {
    constexpr int NV = TileGeom<NVX>::NV;
    const int C = kp.C;
    const int G = C / VEC;
    const T* __restrict__ vol = static_cast<const T*>(kp.data);
    for (int item = threadIdx.x; item < npts * G; item += TILE_THREADS) {
        const int p = item / G;
        const int c = (item - p * G) * VEC;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int cv = sm.code[p * NV + v];
            if (cv < 0) continue;
            const T* b = vol + (size_t)v * (size_t)kp.sv + (size_t)(cv >> 2) + c;
            const int dx = (cv & 1) ? kp.sx : 0, dy = (cv & 2) ? kp.sy : 0;
            const float4 w = sm.w4[p * NV + v];
            if (VEC == 4) {
                fma4(acc, w.x, Load4<T>::ld(b)); fma4(acc, w.y, Load4<T>::ld(b + dx));
                fma4(acc, w.z, Load4<T>::ld(b + dy)); fma4(acc, w.w, Load4<T>::ld(b + dy + dx));
            } else {
                acc.x = fmaf(w.x, Load4<T>::ld1(b), acc.x); acc.x = fmaf(w.y, Load4<T>::ld1(b + dx), acc.x);
                acc.x = fmaf(w.z, Load4<T>::ld1(b + dy), acc.x); acc.x = fmaf(w.w, Load4<T>::ld1(b + dy + dx), acc.x);
            }
        }
        if (kp.bias) {                                   // affine epilogue: out = field - bias (PCA mean term)
            if (VEC == 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(kp.bias + c));
                acc.x -= b.x; acc.y -= b.y; acc.z -= b.z; acc.w -= b.w;
            } else {
                acc.x -= __ldg(kp.bias + c);
            }
        }
        float* o = kp.out + (size_t)(ORDERED ? sm.row[p] : tile0 + p) * C + c;
        if (VEC == 4) __stcs(reinterpret_cast<float4*>(o), acc);
        else          __stcs(o, acc.x);
    }
}

// Elements spanned by one view of a key (strided): the packed footprint codes hold offsets below this.
__host__ __device__ inline long long key_extent(int h, int w, int C, long long sy, long long sx) {
    return (long long)(h - 1) * sy + (long long)(w - 1) * sx + C;
}
// wide-path eligibility of one key for a slice of `slice` channels (host and device agree through this one function)
__host__ __device__ inline bool key_is_wide(int dtype, int C, int h, int w, long long sv, long long sy, long long sx, int slice) {
    return dtype == D3F_F32 && (C % slice) == 0 && ((sv | sy | sx) & 3) == 0 && key_extent(h, w, C, sy, sx) < (1ll << 31);
}
// the tile kernel's narrow path packs (offset << 2 | step bits) into 31 bits
__host__ __device__ inline bool key_fits_tile(int dtype, int C, int h, int w, long long sv, long long sy, long long sx, int slice) {
    return key_is_wide(dtype, C, h, w, sv, sy, sx, slice) || key_extent(h, w, C, sy, sx) < (1ll << 29);
}

// WIDE=false compiles the register-cached walk out: launches with only narrow keys (instance masks, colours,
// PCA-projected volumes) or no keys at all (dist / valid_mask sweeps) then need ~60 registers and run 4 CTAs
// per SM instead of 2.
template <bool RECIP, int VARIANT, bool WIDE, bool ORDERED, int NVX>
__global__ void __launch_bounds__(TILE_THREADS, WIDE ? 2 : 4)
field_tile_kernel(const EvalParams ep, const KeySet ks) {
    using G = TileGeom<NVX>;
    constexpr int PTS = G::PTS;
    constexpr int NV = G::NV;
    __shared__ TileSmemT<WIDE, NVX> sm;
    const int V = ep.V;
    const bool eval_dist = (ep.flags & D3F_FLAG_EVAL_DIST) != 0;
    // Tiles are taken in order.  Dealing them to the CTAs with a stride (so that the CTAs resident at one time mix
    // all-zero store-bound rows with issue-bound visible rows) was measured and is slower: 0.798 vs 0.781 ms on cfg2a,
    // 2.15 vs 1.92 ms on cfg2b (profiles/r02_experiment_tile_stride.jsonl) — neighbouring tiles share texels in L1/L2.
    const int64_t tile0 = (int64_t)blockIdx.x * PTS;
    const int npts = (int)min((int64_t)PTS, ep.n - tile0);

    for (int r = threadIdx.x; r < V * 3; r += TILE_THREADS) {
        const int v = r / 3, i = r - v * 3;
        float row[4];
        krt_row(ep.K + v * 9, ep.pose + v * 12, i, row);
        sm.H[v * 12 + i * 4 + 0] = row[0]; sm.H[v * 12 + i * 4 + 1] = row[1];
        sm.H[v * 12 + i * 4 + 2] = row[2]; sm.H[v * 12 + i * 4 + 3] = row[3];
    }
    __syncthreads();

    // phase 1: view-major so a warp walks 32 consecutive points of one view
    for (int item = threadIdx.x; item < PTS * V; item += TILE_THREADS) {
        const int v = item / PTS, p = item - v * PTS;
        if (p >= npts) continue;
        int64_t row = tile0 + p;
        if (ORDERED) {
            row = __ldg(ep.order + tile0 + p);
            if (v == 0) sm.row[p] = row;
        }
        const float* q = ep.pts + (size_t)row * 3;
        const float x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
        float Hm[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) Hm[j] = sm.H[v * 12 + j];
        ViewSample smp = view_sample<RECIP>(Hm, x, y, z, ep.depth + (size_t)v * ep.H * ep.W, ep.H, ep.W, ep.mu, eval_dist,
                                            ks.n_keys > 0);
        const int s = p * NV + v;
        sm.px[s] = smp.px; sm.py[s] = smp.py;
        sm.d[s] = eval_dist ? smp.d : fminf(fmaxf(smp.d, -ep.mu), ep.mu);    // fusion.py:358
        sm.fac[s] = smp.weight;
        sm.vis[s] = smp.vis ? 1 : 0;
    }
    __syncthreads();

    // phase 1r: views in order (fusion.py:364-370)
    if (threadIdx.x < npts) {
        const int p = threadIdx.x;
        float acc = 0.f, cnt = 0.f;
        for (int v = 0; v < V; ++v)
            if (sm.vis[p * NV + v]) { acc = __fadd_rn(acc, sm.d[p * NV + v]); cnt = __fadd_rn(cnt, 1.f); }
        const float denom = __fadd_rn(cnt, 1e-6f);
        float dist = __fdiv_rn(acc, denom);
        if (!eval_dist && cnt == 0.f) dist = 1e3f;                           // fusion.py:367
        store_compact(ep, ORDERED ? sm.row[p] : tile0 + p, dist, cnt != 0.f ? 1 : 0);
        const float inv = __fdiv_rn(1.f, denom);
        for (int v = 0; v < V; ++v) {
            const int s = p * NV + v;
            sm.fac[s] = sm.vis[s] ? __fmul_rn(sm.fac[s], inv) : 0.f;          // weight/(count+1e-6), fusion.py:385
        }
    }
    if (eval_dist || ks.n_keys == 0) {
        gather_epilogue(ep);
        return;
    }
    __syncthreads();

    for (int k = 0; k < ks.n_keys; ++k) {
        const KeyParams& kp = ks.k[k];
        const bool wide = WIDE && key_is_wide(ks.dtype[k], kp.C, kp.h, kp.w, kp.sv, kp.sy, kp.sx, G::SLICE);
        for (int s = threadIdx.x; s < PTS * NV; s += TILE_THREADS) {
            const int p = s >> G::LOG_NV, v = s & (NV - 1);
            int code = -1;
            if (p < npts && v < V && sm.vis[s]) {
                const Footprint f = footprint<RECIP>(sm.px[s], sm.py[s], ep.H, ep.W, kp.h, kp.w);
                const float fac = sm.fac[s];
                sm.w4[s] = make_float4(f.w[0] * fac, f.w[1] * fac, f.w[2] * fac, f.w[3] * fac);
                const int eo = f.y0 * kp.sy + f.x0 * kp.sx;
                code = (wide ? eo : (eo << 2)) | (f.dx ? 1 : 0) | (f.dy ? 2 : 0);
            }
            sm.code[s] = code;
        }
        __syncthreads();
        if (wide) {
            // per-point mask: which views see the point, and which of those changed texel cell since the
            // previous point of the same run (a view the previous point did not see always reloads)
            if (threadIdx.x < PTS) {
                const int p = threadIdx.x;
                const int run = wide_run_len(kp.C / G::SLICE, PTS);
                const bool first = (p % run) == 0;
                unsigned m = 0;
                if (p < npts) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int c = sm.code[p * NV + v];
                        const int q = first ? -1 : sm.code[(p - 1) * NV + v];
                        if (c >= 0) m |= (1u << v) | ((c != q) ? (1u << (G::CHG + v)) : 0u);
                    }
                }
                sm.mask[p] = (int)m;
            }
            if (threadIdx.x < 4) sm.mask[PTS + threadIdx.x] = 0;         // padding read by the walk's lookahead
            if (VARIANT & 4) {
                __syncthreads();
                int ahead = 0;
                if (threadIdx.x < PTS) {
                    const int p = threadIdx.x;
                    const int run = wide_run_len(kp.C / G::SLICE, PTS);
                    if ((p % run) + WIDE_LOOKAHEAD < run && p + WIDE_LOOKAHEAD < PTS)
                        ahead = ((sm.mask[p + WIDE_LOOKAHEAD] >> G::CHG) & G::VIEWS) << G::AHEAD;
                }
                __syncthreads();
                if (threadIdx.x < PTS) sm.mask[threadIdx.x] |= ahead;
            }
            __syncthreads();
            if constexpr (WIDE) wide_accumulate<(VARIANT & 4) != 0, ORDERED, NVX>(kp, tile0, npts, sm);
        } else {
            const bool vec4 = (kp.C % 4 == 0) && ((kp.sv | kp.sy | kp.sx) & 3) == 0;
            if (ks.dtype[k] == D3F_F32) {
                if (vec4) narrow_accumulate<float, 4, ORDERED, WIDE, NVX>(kp, tile0, npts, sm);
                else      narrow_accumulate<float, 1, ORDERED, WIDE, NVX>(kp, tile0, npts, sm);
            } else {
                if (vec4) narrow_accumulate<uint8_t, 4, ORDERED, WIDE, NVX>(kp, tile0, npts, sm);
                else      narrow_accumulate<uint8_t, 1, ORDERED, WIDE, NVX>(kp, tile0, npts, sm);
            }
        }
        __syncthreads();
    }
    gather_epilogue(ep);
}

}  // namespace d3f
