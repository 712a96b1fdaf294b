// d3f_tile.cuh — the production field-query kernel (V <= 4, no per-view outputs).
//
// One CTA (256 threads) evaluates a tile of TILE_PTS = 256 consecutive query points:
//
//   phase 0/1/1r  as in d3f_generic.cuh: H = [K@Rt;0001], one thread per (point, view) for
//                 projection / nearest depth / visibility / distance weight, then one thread per
//                 point reduces the views in order -> dist, valid_mask, per-view factor
//   per key       one thread per (point, view) turns the pixel coordinate into a bilinear footprint
//                 on that key's map: four corner weights already multiplied by the view factor
//                 weight/(count+1e-6), and one packed code (north-west texel offset, east/south
//                 step bits; -1 when the view does not see the point)
//   wide maps     (float32, C % 128 == 0, e.g. the 1024-channel DINOv2 volume): a warp owns a
//                 128-channel slice (one float4 per lane) and marches over the tile's points in
//                 order.  The four corner texels of each view stay in registers and are reloaded
//                 only when the packed code changes — neighbouring grid points project into the
//                 same texel cell most of the time, so the 16 KB-per-point gather of the reference
//                 (4 views x 4 corners x C floats) collapses to a reload every few points, served
//                 by L1/L2 (the whole volume is L2-resident).  Per visible view the inner loop is
//                 one 128-bit shared-memory read of the folded weights and 16 FFMA per lane; the
//                 output row is written once with 128-bit streaming stores (st.global.cs), 512
//                 contiguous bytes per warp, so the 4 KB/point output stream never evicts the
//                 feature volume from L2.
//   narrow maps   (instance masks, colours; u8 or f32, any C): threads sweep (point, channel-group).
//
// HBM traffic is the algorithmic minimum: points and depth pixels in, each output byte out once; the
// feature volume is read from HBM once and then lives in L2.
#pragma once
#include "d3f_common.cuh"
#include "d3f_generic.cuh"

namespace d3f {

constexpr int TILE_THREADS = 256;
constexpr int TILE_PTS = 256;            // measured: 256-point tiles beat 128 by 4 % (half the barriers and cold starts per point)
constexpr int TILE_V = 4;             // views supported by this kernel (slots are padded to 4)

// D3F_WALK_FFMA2=1 builds the wide walk with packed FFMA2 accumulation (two channels per instruction, weights stored
// twice in shared memory, 60 KB dynamic shared memory).  Measured on B200 (profiles/r02_experiment_ffma2.jsonl):
// bit-identical results, 1.5 % SLOWER on cfg2a and 4 % slower on a fully visible grid — the FP32 pipe is not what
// limits the walk, and the second LDS.128 per view costs more than the 8 saved issue slots.  Kept for A/B only.
#ifndef D3F_WALK_FFMA2
#define D3F_WALK_FFMA2 0
#endif
constexpr bool WALK_FFMA2 = D3F_WALK_FFMA2 != 0;

template <bool WIDE>
struct TileSmemT {
    float H[TILE_V * 12];
    float px[TILE_PTS * TILE_V];
    float py[TILE_PTS * TILE_V];
    float d[TILE_PTS * TILE_V];
    float fac[TILE_PTS * TILE_V];
    int vis[TILE_PTS * TILE_V];
    float4 w4[TILE_PTS * TILE_V * ((WIDE && WALK_FFMA2) ? 2 : 1)];
                                      // folded corner weights per (point, view).  Narrow keys: slot s holds (w_nw, w_ne, w_sw, w_se).
                                      // Wide keys under D3F_WALK_FFMA2: slots 2s, 2s+1 hold every weight twice,
                                      // (nw,nw,ne,ne) (sw,sw,se,se) — the operand layout of the packed FFMA2
    int4 code[TILE_PTS + 4];          // packed footprint codes of the 4 views of a point: element offset of the
                                      // north-west texel inside the view (wide maps: a multiple of 4, used as is;
                                      // narrow maps: shifted left by 2) | east-step bit | south-step bit << 1; -1 = unseen
    int64_t row[TILE_PTS];            // output row of each point of the tile (ordered launches only)
    int mask[TILE_PTS + 4];           // bits 0-3: view sees the point; bits 4-7: its corner cell differs from
                                      // the previous point's (or the previous point did not see it);
                                      // bits 12-15: bits 4-7 of the point WIDE_LOOKAHEAD further on in the same run
};

// Texel loads: read-only path, kept in L1 with evict-last priority — the next cell a walk enters shares two of its
// four corners with the current one, and the output stream must not push them out (measured -1 %).
__device__ __forceinline__ float4 ldg4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Shared-memory reads of the walk go through explicit 32-bit shared addresses that are computed ONCE per walk and
// pinned in a register (the opaque mov): left to itself the compiler re-derives the CTA's shared-window base
// (S2UR SR_CgaCtaId + ULEA) in front of every point's weight reads, which puts a special-register read at the head of
// each point's dependency chain.
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

// Number of point runs a tile is split into for a map of S 128-channel slices (8 warps per CTA).
__host__ __device__ inline int wide_runs(int S) { return S >= TILE_THREADS / 32 ? 1 : (TILE_THREADS / 32) / S; }
__host__ __device__ inline int wide_run_len(int S) { const int R = wide_runs(S); return (TILE_PTS + R - 1) / R; }

// Reload the four corner texels of view v (lane's 4 channels) from the packed footprint code.
// 32-bit element offsets: a view's map holds fewer than 2^31 elements (checked on the host).
#define D3F_WIDE_RELOAD(v, cv)                                                                     \
    {                                                                                              \
        const float* b_ = vbase[v] + (unsigned)((cv) & ~3);                                        \
        const unsigned dx_ = ((cv) & 1) ? (unsigned)kp.sx : 0u;                                    \
        const unsigned dy_ = ((cv) & 2) ? (unsigned)kp.sy : 0u;                                    \
        cc[v][0] = ldg4(b_); cc[v][1] = ldg4(b_ + dx_);                                            \
        cc[v][2] = ldg4(b_ + dy_); cc[v][3] = ldg4(b_ + (dy_ + dx_));                              \
    }
// acc (4 channels) += w_corner * texel_corner for the four corners of view v, two channels per FFMA2.
// sm_100 issues a 3-register FFMA every other cycle per scheduler; the packed form carries two FMAs per issue, which
// is what lets a fully visible tile run at the HBM rate instead of the FP32 pipe's (measured: DESIGN.md §4.8).
// Per channel the order of the additions is the same as the scalar form's: nw, ne, sw, se — results are bit-identical.
#define D3F_WIDE_FMA1(v, w_)                                                                       \
    {                                                                                              \
        fma4(acc, w_.x, cc[v][0]); fma4(acc, w_.y, cc[v][1]);                                      \
        fma4(acc, w_.z, cc[v][2]); fma4(acc, w_.w, cc[v][3]);                                      \
    }
#define D3F_WIDE_FMA(v, wa_, wb_)                                                                  \
    {                                                                                              \
        const float2 w0_ = make_float2(wa_.x, wa_.y), w1_ = make_float2(wa_.z, wa_.w);             \
        const float2 w2_ = make_float2(wb_.x, wb_.y), w3_ = make_float2(wb_.z, wb_.w);             \
        acc01 = __ffma2_rn(make_float2(cc[v][0].x, cc[v][0].y), w0_, acc01);                       \
        acc23 = __ffma2_rn(make_float2(cc[v][0].z, cc[v][0].w), w0_, acc23);                       \
        acc01 = __ffma2_rn(make_float2(cc[v][1].x, cc[v][1].y), w1_, acc01);                       \
        acc23 = __ffma2_rn(make_float2(cc[v][1].z, cc[v][1].w), w1_, acc23);                       \
        acc01 = __ffma2_rn(make_float2(cc[v][2].x, cc[v][2].y), w2_, acc01);                       \
        acc23 = __ffma2_rn(make_float2(cc[v][2].z, cc[v][2].w), w2_, acc23);                       \
        acc01 = __ffma2_rn(make_float2(cc[v][3].x, cc[v][3].y), w3_, acc01);                       \
        acc23 = __ffma2_rn(make_float2(cc[v][3].z, cc[v][3].w), w3_, acc23);                       \
    }

constexpr int WIDE_LOOKAHEAD = 4;      // points between the L1 prefetch of a cell change and its reload

// PREFETCH: bits 12-15 of a point's mask word say which views change cell WIDE_LOOKAHEAD points later; the warp
// then prefetches its own 512-byte slice of those corner texels into L1, so the reload that follows is an L1 hit
// instead of an L2 round trip with all eight warps of the CTA stalled on the same point.
template <bool PREFETCH, bool ORDERED>
__device__ __forceinline__ void wide_accumulate(const KeyParams& kp, int64_t tile0, int npts, const TileSmemT<true>& sm) {
    const int C = kp.C;
    const int S = C >> 7;                                   // 128-channel slices
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = TILE_THREADS / 32;
    int s0, sstep, p_begin, p_end;
    if (S >= nwarps) {                                      // every warp walks the whole tile on its slices
        s0 = warp; sstep = nwarps; p_begin = 0; p_end = npts;
    } else {                                                // fewer slices than warps: split the tile into runs
        const int r = warp / S;
        if (r >= wide_runs(S)) return;
        const int run = wide_run_len(S);
        s0 = warp - r * S; sstep = S;
        p_begin = r * run; p_end = min(p_begin + run, npts);
    }
    if (p_end <= p_begin) return;
    const float* __restrict__ vol = static_cast<const float*>(kp.data);
    const unsigned a_mask = smem_addr(sm.mask), a_code = smem_addr(sm.code), a_w4 = smem_addr(sm.w4);
    for (int s = s0; s < S; s += sstep) {
        const float* vbase[TILE_V];
#pragma unroll
        for (int v = 0; v < TILE_V; ++v) vbase[v] = vol + (size_t)v * (size_t)kp.sv + s * 128 + lane * 4;
        float* o = kp.out + (size_t)(tile0 + p_begin) * C + s * 128 + lane * 4;
        float* const o_col = kp.out + s * 128 + lane * 4;
        float4 cc[TILE_V][4];
#pragma unroll
        for (int v = 0; v < TILE_V; ++v)
#pragma unroll
            for (int q = 0; q < 4; ++q) cc[v][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int m_next = lds_i(a_mask + p_begin * 4);
        for (int p = p_begin; p < p_end; ++p, o += C) {
            // the mask is the same in every lane; the OR-reduction moves it to a uniform register so the
            // tests below are uniform branches (no divergence bookkeeping)
            const unsigned m = __reduce_or_sync(0xffffffffu, (unsigned)m_next);
            m_next = lds_i(a_mask + (p + 1) * 4);                     // the array is padded by one
            if (PREFETCH && (m & 0xF000u)) {
                const int4 code = lds_i4(a_code + (p + WIDE_LOOKAHEAD) * 16);
                const int cvs[4] = {code.x, code.y, code.z, code.w};
#pragma unroll
                for (int v = 0; v < TILE_V; ++v) {
                    if (m & (0x1000u << v)) {
                        const int cv = cvs[v];
                        const float* b_ = vbase[v] + (unsigned)(cv & ~3);
                        const unsigned dx_ = (cv & 1) ? (unsigned)kp.sx : 0u;
                        const unsigned dy_ = (cv & 2) ? (unsigned)kp.sy : 0u;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_ + dx_));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_ + dy_));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(b_ + (dy_ + dx_)));
                    }
                }
            }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#if D3F_WALK_FFMA2
            float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
#endif
            if (m & 0xFFu) {
                if (m & 0xF0u) {
                    const int4 code = lds_i4(a_code + p * 16);
                    if (m & 0x10u) D3F_WIDE_RELOAD(0, code.x)
                    if (m & 0x20u) D3F_WIDE_RELOAD(1, code.y)
                    if (m & 0x40u) D3F_WIDE_RELOAD(2, code.z)
                    if (m & 0x80u) D3F_WIDE_RELOAD(3, code.w)
                }
#if D3F_WALK_FFMA2
                // software-pipelined: the weight reads of view v+2 are issued behind the FMAs of view v, so two views'
                // weights (16 registers) are in flight at a time
                float4 wa0, wb0, wa1, wb1;
                const float4* wp = sm.w4 + (size_t)p * (TILE_V * 2);
                if (m & 1u) { wa0 = wp[0]; wb0 = wp[1]; }
                if (m & 2u) { wa1 = wp[2]; wb1 = wp[3]; }
                if (m & 1u) D3F_WIDE_FMA(0, wa0, wb0)
                if (m & 4u) { wa0 = wp[4]; wb0 = wp[5]; }
                if (m & 2u) D3F_WIDE_FMA(1, wa1, wb1)
                if (m & 8u) { wa1 = wp[6]; wb1 = wp[7]; }
                if (m & 4u) D3F_WIDE_FMA(2, wa0, wb0)
                if (m & 8u) D3F_WIDE_FMA(3, wa1, wb1)
                acc = make_float4(acc01.x, acc01.y, acc23.x, acc23.y);
#else
                // the weight reads of every visible view are issued before the first FMA.  (Dispatching on the four
                // visibility bits with a 16-way switch instead of this if-chain compiles to a compare tree, not an
                // indexed branch, and measured 3 % slower: profiles/r02_experiment_switch_dispatch.jsonl.)
                float4 w0, w1, w2, w3;
                const unsigned aw = a_w4 + p * (TILE_V * 16);
                if (m & 1u) w0 = lds_f4(aw);
                if (m & 2u) w1 = lds_f4(aw + 16);
                if (m & 4u) w2 = lds_f4(aw + 32);
                if (m & 8u) w3 = lds_f4(aw + 48);
                if (m & 1u) D3F_WIDE_FMA1(0, w0)
                if (m & 2u) D3F_WIDE_FMA1(1, w1)
                if (m & 4u) D3F_WIDE_FMA1(2, w2)
                if (m & 8u) D3F_WIDE_FMA1(3, w3)
#endif
            }
            if (ORDERED) __stcs(reinterpret_cast<float4*>(o_col + (size_t)sm.row[p] * C), acc);
            else         __stcs(reinterpret_cast<float4*>(o), acc);
        }
    }
}

template <typename T, int VEC, bool ORDERED, bool WIDE>
__device__ __forceinline__ void narrow_accumulate(const KeyParams& kp, int64_t tile0, int npts, const TileSmemT<WIDE>& sm) {
    const int C = kp.C;
    const int G = C / VEC;
    const T* __restrict__ vol = static_cast<const T*>(kp.data);
    const int* codes = reinterpret_cast<const int*>(sm.code);
    for (int item = threadIdx.x; item < npts * G; item += TILE_THREADS) {
        const int p = item / G;
        const int c = (item - p * G) * VEC;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int v = 0; v < TILE_V; ++v) {
            const int cv = codes[p * TILE_V + v];
            if (cv < 0) continue;
            const T* b = vol + (size_t)v * (size_t)kp.sv + (size_t)(cv >> 2) + c;
            const int dx = (cv & 1) ? kp.sx : 0, dy = (cv & 2) ? kp.sy : 0;
            const float4 w = sm.w4[p * TILE_V + v];
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
// wide-path eligibility of one key (host and device agree through this one function)
__host__ __device__ inline bool key_is_wide(int dtype, int C, int h, int w, long long sv, long long sy, long long sx) {
    return dtype == D3F_F32 && (C % 128) == 0 && ((sv | sy | sx) & 3) == 0 && key_extent(h, w, C, sy, sx) < (1ll << 31);
}
// the tile kernel's narrow path packs (offset << 2 | step bits) into 31 bits
__host__ __device__ inline bool key_fits_tile(int dtype, int C, int h, int w, long long sv, long long sy, long long sx) {
    return key_is_wide(dtype, C, h, w, sv, sy, sx) || key_extent(h, w, C, sy, sx) < (1ll << 29);
}

// WIDE=false compiles the register-cached walk out: launches with only narrow keys (instance masks, colours,
// PCA-projected volumes) or no keys at all (dist / valid_mask sweeps) then need ~60 registers and run 4 CTAs
// per SM instead of 2.
template <bool RECIP, int VARIANT, bool WIDE, bool ORDERED>
__global__ void __launch_bounds__(TILE_THREADS, WIDE ? 2 : 4)
field_tile_kernel(const EvalParams ep, const KeySet ks) {
#if D3F_WALK_FFMA2
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];       // sizeof(TileSmemT<WIDE>), above 48 KB when WIDE
    TileSmemT<WIDE>& sm = *reinterpret_cast<TileSmemT<WIDE>*>(tile_smem_raw);
#else
    __shared__ TileSmemT<WIDE> sm;
#endif
    const int V = ep.V;
    const bool eval_dist = (ep.flags & D3F_FLAG_EVAL_DIST) != 0;
    // Tiles are taken in order.  Dealing them to the CTAs with a stride (so that the CTAs resident at one time mix
    // all-zero store-bound rows with issue-bound visible rows) was measured and is slower: 0.798 vs 0.781 ms on cfg2a,
    // 2.15 vs 1.92 ms on cfg2b (profiles/r02_experiment_tile_stride.jsonl) — neighbouring tiles share texels in L1/L2.
    const int64_t tile0 = (int64_t)blockIdx.x * TILE_PTS;
    const int npts = (int)min((int64_t)TILE_PTS, ep.n - tile0);

    for (int r = threadIdx.x; r < V * 3; r += TILE_THREADS) {
        const int v = r / 3, i = r - v * 3;
        float row[4];
        krt_row(ep.K + v * 9, ep.pose + v * 12, i, row);
        sm.H[v * 12 + i * 4 + 0] = row[0]; sm.H[v * 12 + i * 4 + 1] = row[1];
        sm.H[v * 12 + i * 4 + 2] = row[2]; sm.H[v * 12 + i * 4 + 3] = row[3];
    }
    __syncthreads();

    // phase 1: view-major so a warp walks 32 consecutive points of one view
    for (int item = threadIdx.x; item < TILE_PTS * V; item += TILE_THREADS) {
        const int v = item / TILE_PTS, p = item - v * TILE_PTS;
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
        ViewSample smp = view_sample<RECIP>(Hm, x, y, z, ep.depth + (size_t)v * ep.H * ep.W, ep.H, ep.W, ep.mu, eval_dist);
        const int s = p * TILE_V + v;
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
            if (sm.vis[p * TILE_V + v]) { acc = __fadd_rn(acc, sm.d[p * TILE_V + v]); cnt = __fadd_rn(cnt, 1.f); }
        const float denom = __fadd_rn(cnt, 1e-6f);
        float dist = __fdiv_rn(acc, denom);
        if (!eval_dist && cnt == 0.f) dist = 1e3f;                           // fusion.py:367
        store_compact(ep, ORDERED ? sm.row[p] : tile0 + p, dist, cnt != 0.f ? 1 : 0);
        const float inv = __fdiv_rn(1.f, denom);
        for (int v = 0; v < V; ++v) {
            const int s = p * TILE_V + v;
            sm.fac[s] = sm.vis[s] ? __fmul_rn(sm.fac[s], inv) : 0.f;          // weight/(count+1e-6), fusion.py:385
        }
    }
    if (eval_dist || ks.n_keys == 0) {
        gather_epilogue(ep);
        return;
    }
    __syncthreads();

    int* codes = reinterpret_cast<int*>(sm.code);
    for (int k = 0; k < ks.n_keys; ++k) {
        const KeyParams& kp = ks.k[k];
        const bool wide = WIDE && key_is_wide(ks.dtype[k], kp.C, kp.h, kp.w, kp.sv, kp.sy, kp.sx);
        for (int s = threadIdx.x; s < TILE_PTS * TILE_V; s += TILE_THREADS) {
            const int p = s >> 2, v = s & 3;
            int code = -1;
            if (p < npts && v < V && sm.vis[s]) {
                const Footprint f = footprint<RECIP>(sm.px[s], sm.py[s], ep.H, ep.W, kp.h, kp.w);
                const float fac = sm.fac[s];
                const float w0 = f.w[0] * fac, w1 = f.w[1] * fac, w2 = f.w[2] * fac, w3 = f.w[3] * fac;
                if (WIDE && WALK_FFMA2 && wide) {
                    sm.w4[2 * s] = make_float4(w0, w0, w1, w1);
                    sm.w4[2 * s + 1] = make_float4(w2, w2, w3, w3);
                } else {
                    sm.w4[s] = make_float4(w0, w1, w2, w3);
                }
                const int eo = f.y0 * kp.sy + f.x0 * kp.sx;
                code = (wide ? eo : (eo << 2)) | (f.dx ? 1 : 0) | (f.dy ? 2 : 0);
            }
            codes[s] = code;
        }
        __syncthreads();
        if (wide) {
            // per-point mask: which views see the point, and which of those changed texel cell since the
            // previous point of the same run (a view the previous point did not see always reloads)
            if (threadIdx.x < TILE_PTS) {
                const int p = threadIdx.x;
                const int run = wide_run_len(kp.C >> 7);
                const int4 c = sm.code[p];
                const bool first = (p % run) == 0;
                const int4 q = first ? make_int4(-1, -1, -1, -1) : sm.code[p - 1];
                unsigned m = 0;
                if (p < npts) {
                    if (c.x >= 0) m |= 0x01u | ((c.x != q.x) ? 0x10u : 0u);
                    if (c.y >= 0) m |= 0x02u | ((c.y != q.y) ? 0x20u : 0u);
                    if (c.z >= 0) m |= 0x04u | ((c.z != q.z) ? 0x40u : 0u);
                    if (c.w >= 0) m |= 0x08u | ((c.w != q.w) ? 0x80u : 0u);
                }
                sm.mask[p] = (int)m;
            }
            if (threadIdx.x < 4) sm.mask[TILE_PTS + threadIdx.x] = 0;     // padding read by the walk's lookahead
            if (VARIANT & 4) {
                __syncthreads();
                int ahead = 0;
                if (threadIdx.x < TILE_PTS) {
                    const int p = threadIdx.x;
                    const int run = wide_run_len(kp.C >> 7);
                    if ((p % run) + WIDE_LOOKAHEAD < run && p + WIDE_LOOKAHEAD < TILE_PTS)
                        ahead = (sm.mask[p + WIDE_LOOKAHEAD] & 0xF0) << 8;
                }
                __syncthreads();
                if (threadIdx.x < TILE_PTS) sm.mask[threadIdx.x] |= ahead;
            }
            __syncthreads();
            if constexpr (WIDE) wide_accumulate<(VARIANT & 4) != 0, ORDERED>(kp, tile0, npts, sm);
        } else {
            const bool vec4 = (kp.C % 4 == 0) && ((kp.sv | kp.sy | kp.sx) & 3) == 0;
            if (ks.dtype[k] == D3F_F32) {
                if (vec4) narrow_accumulate<float, 4, ORDERED, WIDE>(kp, tile0, npts, sm);
                else      narrow_accumulate<float, 1, ORDERED, WIDE>(kp, tile0, npts, sm);
            } else {
                if (vec4) narrow_accumulate<uint8_t, 4, ORDERED, WIDE>(kp, tile0, npts, sm);
                else      narrow_accumulate<uint8_t, 1, ORDERED, WIDE>(kp, tile0, npts, sm);
            }
        }
        __syncthreads();
    }
    gather_epilogue(ep);
}

}  // namespace d3f
