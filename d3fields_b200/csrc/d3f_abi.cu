// d3f_abi.cu — the C ABI of libd3f.so (include/d3f.h): argument checking, kernel selection and
// launch.  No torch, no host-side math on tensors; every entry point is asynchronous on the
// caller's stream except d3f_eval_host.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only; ranges cost nanoseconds when no profiler is attached
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/d3f.h"
#include "d3f_common.cuh"
#include "d3f_generic.cuh"
#include "d3f_tile.cuh"
#include "d3f_aux.cuh"
#include "d3f_backward.cuh"
#include "d3f_track.cuh"
#include "d3f_sweep.cuh"
#include "d3f_bin.cuh"
#include "d3f_comm.cuh"
#include <vector>

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
thread_local const char* g_variant[D3F_MAX_KEYS] = {};

// NVTX range around an entry point: shows up as d3f_* in Nsight Systems / ncu --nvtx timelines.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define D3F_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(D3F_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int check_device() {
    static thread_local int checked_dev = -1;
    int dev = 0;
    D3F_CUDA(cudaGetDevice(&dev));
    if (dev == checked_dev) return D3F_OK;
    int major = 0;
    D3F_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        return fail(D3F_EUNSUPPORTED, "libd3f is built for sm_100a only; device %d is sm_%d*", dev, major);
    checked_dev = dev;
    return D3F_OK;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// strides of a key with the contiguous default filled in
struct KeyStrides { int64_t sv, sy, sx; };
KeyStrides key_strides(const D3FKey& q) {
    if (q.stride_v == 0 && q.stride_y == 0 && q.stride_x == 0)
        return {(int64_t)q.h * q.w * q.C, (int64_t)q.w * q.C, (int64_t)q.C};
    return {q.stride_v, q.stride_y, q.stride_x};
}

int validate_key(const D3FKey& q, int k) {
    if (!q.data) return fail(D3F_EINVAL, "keys[%d].data is NULL", k);
    if (q.dtype != D3F_F32 && q.dtype != D3F_U8) return fail(D3F_EINVAL, "keys[%d].dtype=%d unknown", k, q.dtype);
    if (q.h < 1 || q.w < 1 || q.C < 1) return fail(D3F_EINVAL, "keys[%d] shape (%d,%d,%d) invalid", k, q.h, q.w, q.C);
    if ((int64_t)q.h * q.w >= (1ll << 31)) return fail(D3F_EINVAL, "keys[%d] map too large", k);
    const KeyStrides st = key_strides(q);
    if (st.sv < 0 || st.sy < 0 || st.sx < 0) return fail(D3F_EINVAL, "keys[%d]: negative strides are not supported", k);
    if (d3f::key_extent(q.h, q.w, q.C, st.sy, st.sx) >= (1ll << 31))
        return fail(D3F_EINVAL, "keys[%d]: one view spans 2^31 elements or more", k);
    return D3F_OK;
}

int validate(const D3FObs* obs, const void* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
             const void* dist, const void* valid, float* const* out, uint32_t flags, float mu,
             bool need_compact = true) {
    if (!obs) return fail(D3F_EINVAL, "obs is NULL");
    if (obs->V < 1 || obs->V > D3F_MAX_VIEWS) return fail(D3F_EINVAL, "V=%d outside 1..%d", obs->V, D3F_MAX_VIEWS);
    if (obs->H < 2 || obs->W < 2) return fail(D3F_EINVAL, "image size %dx%d must be at least 2x2", obs->H, obs->W);
    if (!obs->pose || !obs->K || !obs->depth) return fail(D3F_EINVAL, "obs pose/K/depth pointer is NULL");
    if (n < 0) return fail(D3F_EINVAL, "n=%lld is negative", (long long)n);
    if (n > 0 && (!pts || (need_compact && (!dist || !valid)))) return fail(D3F_EINVAL, "pts/dist/valid pointer is NULL");
    if (!(mu > 0.f)) return fail(D3F_EINVAL, "mu=%g must be positive", (double)mu);
    if (flags & ~(D3F_FLAG_EVAL_DIST | D3F_FLAG_RECIP_NORM)) return fail(D3F_EINVAL, "unknown flag bits 0x%x", flags);
    if (n_keys < 0 || n_keys > D3F_MAX_KEYS) return fail(D3F_EINVAL, "n_keys=%d outside 0..%d", n_keys, D3F_MAX_KEYS);
    if (n_keys > 0 && (!keys || !out)) return fail(D3F_EINVAL, "keys/out is NULL with n_keys=%d", n_keys);
    for (int k = 0; k < n_keys; ++k) {
        const D3FKey& q = keys[k];
        const int rk = validate_key(q, k);
        if (rk) return rk;
        if (n > 0 && !out[k]) return fail(D3F_EINVAL, "out[%d] is NULL", k);
        if (q.bias && q.C >= 128) return fail(D3F_EINVAL, "keys[%d].bias is supported for C < 128 only", k);
        if (q.bias && q.C % 4 == 0 && !aligned(q.bias, 16)) return fail(D3F_EINVAL, "keys[%d].bias must be 16-byte aligned", k);
    }
    return D3F_OK;
}

// a rank with no points still has to take part in the epoch exchange of a gathering launch
__global__ void gather_only_kernel(const d3f::EvalParams ep) { d3f::gather_epilogue(ep); }

template <bool RECIP, int VARIANT, bool WIDE, int NV>
void launch_tile(bool ordered, dim3 grid, dim3 block, cudaStream_t st, const d3f::EvalParams& ep, const d3f::KeySet& ks) {
    if (ordered) d3f::field_tile_kernel<RECIP, VARIANT, WIDE, true, NV><<<grid, block, 0, st>>>(ep, ks);
    else         d3f::field_tile_kernel<RECIP, VARIANT, WIDE, false, NV><<<grid, block, 0, st>>>(ep, ks);
}

// NV = 4: up to four views (every configuration the reference runs).  NV = 8: five to eight views.  NV = 0: four view
// slots with 32-point tiles (NV = 1: 8-point tiles), for launches too small to fill the GPU with 256-point tiles
// (latency: a warp walks its tile serially).  The lookahead prefetch variant is instantiated for NV = 4 only.
template <int NV>
void launch_tile_nv(bool recip, bool wide, bool prefetch, bool ordered, dim3 grid, dim3 block, cudaStream_t st,
                    const d3f::EvalParams& ep, const d3f::KeySet& ks) {
    if (!wide) {                        // no register-cached walk needed: the light instantiation (4 CTAs/SM)
        if (recip) launch_tile<true, 0, false, NV>(ordered, grid, block, st, ep, ks);
        else       launch_tile<false, 0, false, NV>(ordered, grid, block, st, ep, ks);
    } else if (NV == 4 && prefetch) {
        if (recip) launch_tile<true, 4, true, 4>(ordered, grid, block, st, ep, ks);
        else       launch_tile<false, 4, true, 4>(ordered, grid, block, st, ep, ks);
    } else {
        if (recip) launch_tile<true, 0, true, NV>(ordered, grid, block, st, ep, ks);
        else       launch_tile<false, 0, true, NV>(ordered, grid, block, st, ep, ks);
    }
}

// Device-pointer evaluation shared by d3f_eval, d3f_eval_ordered, d3f_eval_allgather and the slabs of d3f_eval_host.
int launch_eval(const D3FObs* obs, const float* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
                float* dist, uint8_t* valid, float* const* out, float* const* out_inter,
                uint32_t flags, float mu, cudaStream_t st,
                const int32_t* order = nullptr, const d3f::GatherParams* gather = nullptr) {
    d3f::EvalParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.pts = pts; ep.depth = obs->depth; ep.pose = obs->pose; ep.K = obs->K;
    ep.dist = dist; ep.valid = valid; ep.n = n; ep.V = obs->V; ep.H = obs->H; ep.W = obs->W;
    ep.mu = mu; ep.flags = flags; ep.order = order;
    if (gather) ep.g = *gather;
    if (n == 0) {
        if (gather) {                        // a rank without points still takes part in the epoch exchange
            gather_only_kernel<<<1, 32, 0, st>>>(ep);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            D3F_CUDA(cudaGetLastError());
        }
        return D3F_OK;
    }
    const bool eval_dist = (flags & D3F_FLAG_EVAL_DIST) != 0;
    const bool recip = (flags & D3F_FLAG_RECIP_NORM) != 0;

    d3f::KeySet ks;
    memset(&ks, 0, sizeof(ks));
    ks.n_keys = eval_dist ? 0 : n_keys;
    bool any_inter = false;
    for (int k = 0; k < ks.n_keys; ++k) {
        ks.k[k].data = keys[k].data; ks.k[k].out = out[k];
        ks.k[k].inter = (out_inter && out_inter[k]) ? out_inter[k] : nullptr;
        ks.k[k].h = keys[k].h; ks.k[k].w = keys[k].w; ks.k[k].C = keys[k].C;
        const KeyStrides ksd = key_strides(keys[k]);
        ks.k[k].sv = ksd.sv; ks.k[k].sy = (int32_t)ksd.sy; ks.k[k].sx = (int32_t)ksd.sx;
        ks.dtype[k] = keys[k].dtype;
        ks.k[k].bias = keys[k].bias;
        any_inter |= ks.k[k].inter != nullptr;
        if (keys[k].C % 4 == 0 && ((ksd.sv | ksd.sy | ksd.sx) & 3) == 0) {
            const size_t a_in = keys[k].dtype == D3F_F32 ? 16 : 4;
            if (!aligned(keys[k].data, a_in) || !aligned(out[k], 16) || (ks.k[k].inter && !aligned(ks.k[k].inter, 16)))
                return fail(D3F_EINVAL, "keys[%d]: map/out pointers must be 16-byte aligned when C %% 4 == 0", k);
        }
        g_variant[k] = "generic";
    }
    // production path: point tiles, register-cached corner texels for wide float32 maps
    // small launches (tracking: a few hundred points; mesh vertices: tens of thousands) take 32-point tiles; measured
    // crossover between 100 000 points (149 vs 171 us) and 150 000 (208 vs 182 us): profiles/r02_ab_small_tiles.jsonl
    static const int64_t small_n = [] { const char* e = getenv("D3F_SMALL_TILE_N"); return e ? atoll(e) : 100000ll; }();
    static const int64_t tiny_n = [] { const char* e = getenv("D3F_TINY_TILE_N"); return e ? atoll(e) : 2048ll; }();
    const int nv = obs->V <= 4 ? ((n <= small_n && !order) ? (n <= tiny_n ? 1 : 0) : 4) : 8;
    const int slice = nv == 8 ? d3f::TileGeom<8>::SLICE : d3f::TileGeom<4>::SLICE;
    const int tile_pts = nv == 1 ? d3f::TileGeom<1>::PTS : nv == 0 ? d3f::TileGeom<0>::PTS : (nv == 4 ? d3f::TileGeom<4>::PTS : d3f::TileGeom<8>::PTS);
    auto is_wide = [&](int k) {
        return d3f::key_is_wide(keys[k].dtype, keys[k].C, keys[k].h, keys[k].w, ks.k[k].sv, ks.k[k].sy, ks.k[k].sx, slice);
    };
    static const bool force_generic = [] { const char* e = getenv("D3F_FORCE_GENERIC"); return e && atoi(e) != 0; }();   // A/B only
    bool tile_ok = obs->V <= d3f::TILE_MAX_V && !any_inter && !force_generic;
    for (int k = 0; k < ks.n_keys; ++k)
        tile_ok &= d3f::key_fits_tile(keys[k].dtype, keys[k].C, keys[k].h, keys[k].w, ks.k[k].sv, ks.k[k].sy, ks.k[k].sx, slice);
    if (any_inter && order) return fail(D3F_EINVAL, "per-view outputs are not available on an ordered launch");
    if (tile_ok) {
        const int64_t tiles = (n + tile_pts - 1) / tile_pts;
        if (tiles > 0x7fffffffll) return fail(D3F_EINVAL, "n=%lld too large for one launch", (long long)n);
        for (int k = 0; k < ks.n_keys; ++k)
            g_variant[k] = is_wide(k) ? "tile/wide" : "tile/narrow";
        dim3 grid((unsigned)tiles), block(d3f::TILE_THREADS);
        // L1 lookahead prefetch of cell changes pays only when the wide volume cannot stay L2-resident
        // (measured: cfg2b 5 GB volume 2.02 -> 1.87 ms; cfg2a 50 MB volume 0.80 -> 0.88 ms).  D3F_TILE_PREFETCH=0/1 overrides.
        static const int force = [] { const char* e = getenv("D3F_TILE_PREFETCH"); return e ? atoi(e) : -1; }();
        size_t wide_bytes = 0;
        for (int k = 0; k < ks.n_keys; ++k)
            if (is_wide(k)) wide_bytes += (size_t)obs->V * keys[k].h * keys[k].w * keys[k].C * 4;
        const bool prefetch = force >= 0 ? force != 0 : wide_bytes > (size_t)(64u << 20);
        if (nv == 4)      launch_tile_nv<4>(recip, wide_bytes != 0, prefetch, order != nullptr, grid, block, st, ep, ks);
        else if (nv == 0) launch_tile_nv<0>(recip, wide_bytes != 0, false, false, grid, block, st, ep, ks);
        else if (nv == 1) launch_tile_nv<1>(recip, wide_bytes != 0, false, false, grid, block, st, ep, ks);
        else              launch_tile_nv<8>(recip, wide_bytes != 0, prefetch, order != nullptr, grid, block, st, ep, ks);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        D3F_CUDA(cudaGetLastError());
        return D3F_OK;
    }
    const int64_t tiles = (n + d3f::GEN_TILE_PTS - 1) / d3f::GEN_TILE_PTS;
    if (tiles > 0x7fffffffll) return fail(D3F_EINVAL, "n=%lld too large for one launch", (long long)n);
    const size_t smem = d3f::generic_smem_bytes(obs->V);
    if (smem > 48 * 1024) {     // V > 15: opt in to more than the default 48 KB of dynamic shared memory
        static std::once_flag once;
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(once, [] {
            const int big = (int)d3f::generic_smem_bytes(D3F_MAX_VIEWS);
            cudaError_t e;
            if ((e = cudaFuncSetAttribute(d3f::field_generic_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big))) attr_err = e;
            if ((e = cudaFuncSetAttribute(d3f::field_generic_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big))) attr_err = e;
            if ((e = cudaFuncSetAttribute(d3f::field_generic_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big))) attr_err = e;
            if ((e = cudaFuncSetAttribute(d3f::field_generic_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big))) attr_err = e;
        });
        if (attr_err != cudaSuccess) return fail(D3F_ECUDA, "cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
    }
    dim3 grid((unsigned)tiles), block(d3f::GEN_THREADS);
    if (any_inter) {
        if (recip) d3f::field_generic_kernel<true, true><<<grid, block, smem, st>>>(ep, ks);
        else       d3f::field_generic_kernel<true, false><<<grid, block, smem, st>>>(ep, ks);
    } else {
        if (recip) d3f::field_generic_kernel<false, true><<<grid, block, smem, st>>>(ep, ks);
        else       d3f::field_generic_kernel<false, false><<<grid, block, smem, st>>>(ep, ks);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

// ---- scratch for d3f_eval_host: streams + device slabs, pooled so concurrent callers never share one -----
constexpr int HOST_STREAMS = 3;
struct HostScratch {
    int dev = -1;
    cudaStream_t st[HOST_STREAMS] = {};
    cudaEvent_t ready = nullptr;
    void* buf[HOST_STREAMS] = {};
    size_t cap[HOST_STREAMS] = {};
};
std::mutex g_scratch_mu;
std::vector<HostScratch*> g_scratch_free;

void scratch_free(HostScratch* sc) {
    for (int i = 0; i < HOST_STREAMS; ++i) {
        if (sc->buf[i]) cudaFree(sc->buf[i]);
        if (sc->st[i]) cudaStreamDestroy(sc->st[i]);
    }
    if (sc->ready) cudaEventDestroy(sc->ready);
    delete sc;
}

HostScratch* scratch_acquire(int dev) {
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        for (size_t i = 0; i < g_scratch_free.size(); ++i)
            if (g_scratch_free[i]->dev == dev) {
                HostScratch* sc = g_scratch_free[i];
                g_scratch_free.erase(g_scratch_free.begin() + i);
                return sc;
            }
    }
    HostScratch* sc = new HostScratch;
    sc->dev = dev;
    bool ok = cudaEventCreateWithFlags(&sc->ready, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < HOST_STREAMS && ok; ++i)
        ok = cudaStreamCreateWithFlags(&sc->st[i], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { scratch_free(sc); return nullptr; }
    return sc;
}

void scratch_release(HostScratch* sc) {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    g_scratch_free.push_back(sc);
}

// The slab loop of d3f_eval_host; the caller synchronises the scratch streams whatever this returns.
int eval_host_slabs(HostScratch& sc, const D3FObs* obs, const float* pts_host, int64_t n, const D3FKey* keys, int nk,
                    float* dist_host, uint8_t* valid_host, float* const* out_host, uint32_t flags, float mu,
                    int64_t slab, size_t need) {
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    for (int i = 0; i < HOST_STREAMS; ++i) {
        if (sc.cap[i] < need) {
            if (sc.buf[i]) D3F_CUDA(cudaFree(sc.buf[i]));
            sc.buf[i] = nullptr; sc.cap[i] = 0;
            D3F_CUDA(cudaMalloc(&sc.buf[i], need));
            sc.cap[i] = need;
        }
    }
    int it = 0;
    for (int64_t s0 = 0; s0 < n; s0 += slab, ++it) {
        const int64_t m = (n - s0 < slab) ? (n - s0) : slab;
        const int si = it % HOST_STREAMS;
        cudaStream_t st = sc.st[si];
        char* b = static_cast<char*>(sc.buf[si]);
        float* d_pts = reinterpret_cast<float*>(b);      b += up((size_t)slab * 12);
        float* d_dist = reinterpret_cast<float*>(b);     b += up((size_t)slab * 4);
        uint8_t* d_valid = reinterpret_cast<uint8_t*>(b); b += up((size_t)slab);
        float* d_out[D3F_MAX_KEYS] = {};
        for (int k = 0; k < nk; ++k) { d_out[k] = reinterpret_cast<float*>(b); b += up((size_t)slab * keys[k].C * 4); }
        // stream order on `st` protects the slab buffers: the previous use of this slab ended
        // with its D2H copies on the same stream
        D3F_CUDA(cudaMemcpyAsync(d_pts, pts_host + s0 * 3, (size_t)m * 12, cudaMemcpyHostToDevice, st));
        const int rc = launch_eval(obs, d_pts, m, keys, nk, d_dist, d_valid, d_out, nullptr, flags, mu, st);
        if (rc) return rc;
        D3F_CUDA(cudaMemcpyAsync(dist_host + s0, d_dist, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        D3F_CUDA(cudaMemcpyAsync(valid_host + s0, d_valid, (size_t)m, cudaMemcpyDeviceToHost, st));
        for (int k = 0; k < nk; ++k)
            D3F_CUDA(cudaMemcpyAsync(out_host[k] + (size_t)s0 * keys[k].C, d_out[k], (size_t)m * keys[k].C * 4,
                                     cudaMemcpyDeviceToHost, st));
    }
    return D3F_OK;
}

}  // namespace

extern "C" {

int d3f_abi_version(void) { return D3F_ABI_VERSION; }
const char* d3f_last_error(void) { return g_err; }
int64_t d3f_launch_count(void) { return g_launches.load(); }
const char* d3f_last_variant(int32_t k) { return (k >= 0 && k < D3F_MAX_KEYS && g_variant[k]) ? g_variant[k] : ""; }
int d3f_sizeof_key(void) { return (int)sizeof(D3FKey); }
int d3f_sizeof_obs(void) { return (int)sizeof(D3FObs); }

int d3f_eval(const D3FObs* obs, const float* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
             float* dist, uint8_t* valid, float* const* out, float* const* out_inter,
             uint32_t flags, float mu, void* stream) {
    NvtxRange nvtx_("d3f_eval");
    int rc = validate(obs, pts, n, keys, n_keys, dist, valid, out, flags, mu);
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    return launch_eval(obs, pts, n, keys, n_keys, dist, valid, out, out_inter, flags, mu,
                       static_cast<cudaStream_t>(stream));
}

int d3f_eval_host(const D3FObs* obs, const float* pts_host, int64_t n, const D3FKey* keys, int32_t n_keys,
                  float* dist_host, uint8_t* valid_host, float* const* out_host, uint32_t flags, float mu,
                  void* obs_stream) {
    NvtxRange nvtx_("d3f_eval_host");
    int rc = validate(obs, pts_host, n, keys, n_keys, dist_host, valid_host, out_host, flags, mu);
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    if (n == 0) return D3F_OK;
    const bool eval_dist = (flags & D3F_FLAG_EVAL_DIST) != 0;
    const int nk = eval_dist ? 0 : n_keys;

    // bytes per point on the device side of a slab; every sub-buffer starts 256-byte aligned
    size_t out_bytes_pp = 0;
    for (int k = 0; k < nk; ++k) out_bytes_pp += (size_t)keys[k].C * 4;
    const size_t per_pt = 12 + 4 + 1 + out_bytes_pp;
    int64_t slab = (int64_t)((32ull << 20) / per_pt);
    if (slab < 4096) slab = 4096;
    if (slab > (1 << 20)) slab = 1 << 20;
    slab = (slab / d3f::TILE_PTS) * d3f::TILE_PTS;
    if (slab > n) slab = n;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t need = up((size_t)slab * 12) + up((size_t)slab * 4) + up((size_t)slab) +
                        [&] { size_t s = 0; for (int k = 0; k < nk; ++k) s += up((size_t)slab * keys[k].C * 4); return s; }();

    int dev = 0;
    D3F_CUDA(cudaGetDevice(&dev));
    HostScratch* sc = scratch_acquire(dev);
    if (!sc) return fail(D3F_ECUDA, "d3f_eval_host: cannot create scratch streams");
    // order the internal streams after whatever wrote the observation (no device-wide synchronisation)
    cudaError_t e = cudaEventRecord(sc->ready, static_cast<cudaStream_t>(obs_stream));
    for (int i = 0; i < HOST_STREAMS && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(sc->st[i], sc->ready, 0);
    if (e != cudaSuccess) rc = fail(D3F_ECUDA, "d3f_eval_host: stream ordering failed: %s", cudaGetErrorString(e));
    else rc = eval_host_slabs(*sc, obs, pts_host, n, keys, nk, dist_host, valid_host, out_host, flags, mu, slab, need);
    // synchronous also on failure: no copy into the caller's buffers may still be in flight when we return
    for (int i = 0; i < HOST_STREAMS; ++i) {
        const cudaError_t es = cudaStreamSynchronize(sc->st[i]);
        if (es != cudaSuccess && rc == D3F_OK) rc = fail(D3F_ECUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(es));
    }
    scratch_release(sc);
    return rc;
}

int d3f_release_scratch(void) {
    int dev = 0;
    D3F_CUDA(cudaGetDevice(&dev));
    std::vector<HostScratch*> mine;
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        for (size_t i = 0; i < g_scratch_free.size();) {
            if (g_scratch_free[i]->dev == dev) { mine.push_back(g_scratch_free[i]); g_scratch_free.erase(g_scratch_free.begin() + i); }
            else ++i;
        }
    }
    for (HostScratch* sc : mine) scratch_free(sc);
    return D3F_OK;
}

int d3f_eval_ordered(const D3FObs* obs, const float* pts, int64_t n, const int32_t* order,
                     const D3FKey* keys, int32_t n_keys, float* dist, uint8_t* valid, float* const* out,
                     uint32_t flags, float mu, void* stream) {
    NvtxRange nvtx_("d3f_eval_ordered");
    int rc = validate(obs, pts, n, keys, n_keys, dist, valid, out, flags, mu);
    if (rc) return rc;
    if (n > 0 && !order) return fail(D3F_EINVAL, "order is NULL");
    if (n >= (1ll << 31)) return fail(D3F_EINVAL, "ordered launches index points with int32: n=%lld too large", (long long)n);
    if ((rc = check_device())) return rc;
    return launch_eval(obs, pts, n, keys, n_keys, dist, valid, out, nullptr, flags, mu,
                       static_cast<cudaStream_t>(stream), order, nullptr);
}

int64_t d3f_bin_workspace_bytes(int64_t n) { return (int64_t)d3f::bin_workspace_bytes(n); }

int d3f_bin_order(const float* pts, int64_t n, float cell, int32_t* order, void* workspace, int64_t workspace_bytes,
                  void* stream) {
    NvtxRange nvtx_("d3f_bin_order");
    if (n < 0 || n >= (1ll << 31)) return fail(D3F_EINVAL, "bin: n=%lld outside 0..2^31", (long long)n);
    if (n > 0 && (!pts || !order || !workspace)) return fail(D3F_EINVAL, "bin: NULL pointer");
    if (!(cell > 0.f)) return fail(D3F_EINVAL, "bin: cell=%g must be positive", (double)cell);
    if (workspace_bytes < (int64_t)d3f::bin_workspace_bytes(n))
        return fail(D3F_EINVAL, "bin: workspace of %lld bytes, %lld needed", (long long)workspace_bytes, (long long)d3f::bin_workspace_bytes(n));
    if (!aligned(workspace, 16)) return fail(D3F_EINVAL, "bin: workspace must be 16-byte aligned");
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return D3F_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    int* bbox = reinterpret_cast<int*>(ws);
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws + 256);
    uint32_t* block_base = hist + d3f::BIN_COUNT;
    uint32_t* keys = block_base + d3f::BIN_SCAN_BLOCKS;
    D3F_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * ((size_t)d3f::BIN_COUNT + d3f::BIN_SCAN_BLOCKS), st));
    const unsigned blocks = (unsigned)((n + d3f::BIN_THREADS - 1) / d3f::BIN_THREADS);
    d3f::bin_init_kernel<<<1, 32, 0, st>>>(bbox);
    d3f::bin_bbox_kernel<<<blocks < 1184u ? blocks : 1184u, d3f::BIN_THREADS, 0, st>>>(pts, n, bbox);
    d3f::bin_key_kernel<<<blocks, d3f::BIN_THREADS, 0, st>>>(pts, n, cell, bbox, keys, hist);
    d3f::bin_scan_kernel<<<d3f::BIN_SCAN_BLOCKS, d3f::BIN_SCAN_THREADS, 0, st>>>(hist, block_base);
    d3f::bin_scan_base_kernel<<<1, d3f::BIN_SCAN_BLOCKS, 0, st>>>(block_base);
    d3f::bin_scatter_kernel<<<blocks, d3f::BIN_THREADS, 0, st>>>(keys, n, hist, block_base, order);
    static const bool refine = [] { const char* e = getenv("D3F_BIN_REFINE"); return e ? atoi(e) != 0 : true; }();
    if (refine) d3f::bin_refine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, n, cell, bbox, order);
    g_launches.fetch_add(refine ? 7 : 6, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_sweep_select(const D3FObs* obs, const D3FGrid* grid, const float* pts, int64_t n,
                     const D3FKey* mask_key, float dist_threshold, float mask_threshold,
                     float* dist_out, uint8_t* valid_out,
                     int64_t capacity, int64_t* sel_count, int32_t* sel_index, int32_t* sel_inst,
                     uint32_t flags, float mu, void* stream) {
    NvtxRange nvtx_("d3f_sweep_select");
    if (!obs) return fail(D3F_EINVAL, "obs is NULL");
    if (obs->V < 1 || obs->V > D3F_MAX_VIEWS) return fail(D3F_EINVAL, "V=%d outside 1..%d", obs->V, D3F_MAX_VIEWS);
    if (obs->H < 2 || obs->W < 2) return fail(D3F_EINVAL, "image size %dx%d must be at least 2x2", obs->H, obs->W);
    if (!obs->pose || !obs->K || !obs->depth) return fail(D3F_EINVAL, "obs pose/K/depth pointer is NULL");
    if (!(mu > 0.f)) return fail(D3F_EINVAL, "mu=%g must be positive", (double)mu);
    if (flags & ~D3F_FLAG_RECIP_NORM) return fail(D3F_EINVAL, "sweep: unsupported flag bits 0x%x", flags);
    if (grid) {
        if (grid->nx < 0 || grid->ny < 0 || grid->nz < 0) return fail(D3F_EINVAL, "sweep: negative grid size");
        if ((int64_t)grid->nx * grid->ny * grid->nz != n) return fail(D3F_EINVAL, "sweep: n=%lld is not nx*ny*nz", (long long)n);
        if (n > 0 && (!grid->x || !grid->y || !grid->z)) return fail(D3F_EINVAL, "sweep: grid axis pointer is NULL");
    } else if (n > 0 && !pts) {
        return fail(D3F_EINVAL, "sweep: neither a grid nor pts");
    }
    if (n < 0 || n >= (1ll << 31)) return fail(D3F_EINVAL, "sweep: n=%lld outside 0..2^31 (survivors are int32 indices)", (long long)n);
    d3f::SweepParams sp;
    memset(&sp, 0, sizeof(sp));
    if (mask_key) {
        int rc = validate_key(*mask_key, 0);
        if (rc) return rc;
        if (mask_key->C > d3f::SWEEP_MAX_INST) return fail(D3F_EINVAL, "sweep: num_inst=%d above %d", mask_key->C, d3f::SWEEP_MAX_INST);
        if (capacity < 0 || !sel_count || (capacity > 0 && (!sel_index || !sel_inst)))
            return fail(D3F_EINVAL, "sweep: selection outputs missing");
        const KeyStrides ksd = key_strides(*mask_key);
        sp.mask = mask_key->data; sp.mdtype = mask_key->dtype; sp.mh = mask_key->h; sp.mw = mask_key->w; sp.mC = mask_key->C;
        sp.msv = ksd.sv; sp.msy = (int32_t)ksd.sy; sp.msx = (int32_t)ksd.sx;
    }
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return D3F_OK;
    if (grid) { sp.gx = grid->x; sp.gy = grid->y; sp.gz = grid->z; sp.nx = grid->nx; sp.ny = grid->ny; sp.nz = grid->nz; }
    sp.dist_thr = dist_threshold; sp.mask_thr = mask_threshold;
    sp.dist_out = dist_out; sp.valid_out = valid_out;
    sp.capacity = capacity; sp.count = reinterpret_cast<unsigned long long*>(sel_count);
    sp.sel_index = sel_index; sp.sel_inst = sel_inst;
    d3f::EvalParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.pts = pts; ep.depth = obs->depth; ep.pose = obs->pose; ep.K = obs->K; ep.n = n;
    ep.V = obs->V; ep.H = obs->H; ep.W = obs->W; ep.mu = mu; ep.flags = flags;
    const unsigned blocks = (unsigned)((n + d3f::SWEEP_THREADS - 1) / d3f::SWEEP_THREADS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (flags & D3F_FLAG_RECIP_NORM) d3f::field_sweep_kernel<true><<<blocks, d3f::SWEEP_THREADS, 0, st>>>(ep, sp);
    else                             d3f::field_sweep_kernel<false><<<blocks, d3f::SWEEP_THREADS, 0, st>>>(ep, sp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------
int d3f_comm_create(int32_t rank, int32_t world, int64_t capacity_points, int64_t staging_bytes,
                    D3FComm** comm, void* handle_out) {
    if (!comm) return fail(D3F_EINVAL, "comm is NULL");
    if (world < 1 || world > D3F_MAX_PEERS || rank < 0 || rank >= world)
        return fail(D3F_EINVAL, "comm: rank %d / world %d outside 1..%d", rank, world, D3F_MAX_PEERS);
    if (capacity_points < 0 || staging_bytes < 0) return fail(D3F_EINVAL, "comm: negative capacity");
    if (world > 1 && !handle_out) return fail(D3F_EINVAL, "comm: handle_out is NULL");
    int rc = check_device();
    if (rc) return rc;
    D3FComm* c = new D3FComm;
    c->rank = rank; c->world = world; c->capacity = capacity_points;
    D3F_CUDA(cudaGetDevice(&c->dev));
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 4096;
    for (int b = 0; b < 2; ++b) {
        c->off_dist[b] = off;  off += up((size_t)capacity_points * 4);
        c->off_valid[b] = off; off += up((size_t)capacity_points);
    }
    c->off_staging = off;
    c->staging_bytes = up((size_t)staging_bytes);
    off += c->staging_bytes;
    c->seg_bytes = off;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, c->seg_bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, 4096);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out), p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        delete c;
        return fail(D3F_ECUDA, "comm: segment of %zu bytes: %s", off, cudaGetErrorString(e));
    }
    c->seg[rank] = static_cast<char*>(p);
    c->connected = world == 1;
    *comm = c;
    return D3F_OK;
}

int d3f_comm_connect(D3FComm* c, const void* handles) {
    if (!c) return fail(D3F_EINVAL, "comm is NULL");
    if (c->connected) return D3F_OK;
    if (!handles) return fail(D3F_EINVAL, "comm: handles is NULL");
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(D3F_ECUDA, "comm: cannot map the segment of rank %d (CUDA IPC / peer access): %s", r, cudaGetErrorString(e));
        }
        c->seg[r] = static_cast<char*>(p);
    }
    c->connected = true;
    return D3F_OK;
}

int d3f_comm_destroy(D3FComm* c) {
    if (!c) return D3F_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r) {
        if (!c->seg[r]) continue;
        if (r == c->rank) cudaFree(c->seg[r]);
        else cudaIpcCloseMemHandle(c->seg[r]);
    }
    delete c;
    return D3F_OK;
}

int d3f_comm_status(D3FComm* c, void* stream) {
    if (!c) return fail(D3F_EINVAL, "comm is NULL");
    D3F_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    uint32_t err = 0;
    D3F_CUDA(cudaMemcpy(&err, &c->hdr(c->rank)->error, 4, cudaMemcpyDeviceToHost));
    if (err) return fail(D3F_ETIMEOUT, "comm: a wait on a peer of rank %d timed out (a rank did not join the collective)", c->rank);
    return D3F_OK;
}

int d3f_eval_allgather(D3FComm* c, const D3FObs* obs, const float* pts, int64_t n,
                       const D3FKey* keys, int32_t n_keys, float* const* out,
                       int64_t gather_base, int64_t gather_block, int64_t gather_stride,
                       uint32_t flags, float mu, void* stream, float** dist_all, uint8_t** valid_all) {
    NvtxRange nvtx_("d3f_eval_allgather");
    if (!c || !c->connected) return fail(D3F_EINVAL, "comm is NULL or not connected");
    int rc = validate(obs, pts, n, keys, n_keys, nullptr, nullptr, out, flags, mu, false);
    if (rc) return rc;
    if (gather_block < 1 || gather_base < 0 || gather_stride < 0) return fail(D3F_EINVAL, "gather: bad base/block/stride");
    if (n > 0) {
        const int64_t last = gather_base + ((n - 1) / gather_block) * gather_stride + (n - 1) % gather_block;
        if (last >= c->capacity) return fail(D3F_EINVAL, "gather: index %lld outside the communicator's capacity %lld", (long long)last, (long long)c->capacity);
    }
    if ((rc = check_device())) return rc;
    const uint32_t epoch = ++c->epoch;
    const int b = (int)(epoch & 1u);
    d3f::GatherParams g;
    memset(&g, 0, sizeof(g));
    for (int r = 0; r < c->world; ++r) {
        g.dist[r] = reinterpret_cast<float*>(c->seg[r] + c->off_dist[b]);
        g.valid[r] = reinterpret_cast<uint8_t*>(c->seg[r] + c->off_valid[b]);
        g.flag[r] = &c->hdr(r)->gather_flag[c->rank];
    }
    g.my_flags = c->hdr(c->rank)->gather_flag;
    g.counter = &c->hdr(c->rank)->counter;
    g.error = &c->hdr(c->rank)->error;
    g.base = gather_base; g.block = gather_block; g.stride = gather_stride;
    g.epoch = epoch; g.world = c->world;
    if (dist_all) *dist_all = g.dist[c->rank];
    if (valid_all) *valid_all = g.valid[c->rank];
    return launch_eval(obs, pts, n, keys, n_keys, nullptr, nullptr, out, nullptr, flags, mu,
                       static_cast<cudaStream_t>(stream), nullptr, &g);
}

int d3f_comm_broadcast(D3FComm* c, void* buf, int64_t bytes, int32_t root, void* stream) {
    NvtxRange nvtx_("d3f_comm_broadcast");
    if (!c || !c->connected) return fail(D3F_EINVAL, "comm is NULL or not connected");
    if (root < 0 || root >= c->world) return fail(D3F_EINVAL, "broadcast: root %d outside 0..%d", root, c->world - 1);
    if (bytes < 0 || (bytes > 0 && !buf)) return fail(D3F_EINVAL, "broadcast: bad buffer");
    if (c->world == 1 || bytes == 0) return D3F_OK;
    if (c->staging_bytes == 0) return fail(D3F_EINVAL, "broadcast: the communicator was created without staging");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    d3f::CommHeader* me = c->hdr(c->rank);
    d3f::FlagTargets everyone, ready;
    memset(&everyone, 0, sizeof(everyone));
    memset(&ready, 0, sizeof(ready));
    for (int r = 0; r < c->world; ++r) {
        everyone.p[everyone.count++] = &c->hdr(r)->bcast_taken[c->rank];
        if (r != c->rank) ready.p[ready.count++] = &c->hdr(r)->bcast_ready;
    }
    char* src = static_cast<char*>(buf);
    for (int64_t off = 0; off < bytes; off += (int64_t)c->staging_bytes) {
        const size_t m = (size_t)((bytes - off < (int64_t)c->staging_bytes) ? bytes - off : (int64_t)c->staging_bytes);
        const uint32_t seq = ++c->bcast_seq;
        if (c->rank == root) {
            // every rank must be done with the previous chunk before its staging area is overwritten
            d3f::comm_wait_kernel<<<1, 32, 0, st>>>(me->bcast_taken, c->world, seq - 1, &me->error);
            for (int r = 0; r < c->world; ++r)
                if (r != root) D3F_CUDA(cudaMemcpyAsync(c->seg[r] + c->off_staging, src + off, m, cudaMemcpyDefault, st));
            d3f::comm_signal_kernel<<<1, 32, 0, st>>>(ready, seq);
        } else {
            d3f::comm_wait_kernel<<<1, 32, 0, st>>>(&me->bcast_ready, 1, seq, &me->error);
            D3F_CUDA(cudaMemcpyAsync(src + off, c->seg[c->rank] + c->off_staging, m, cudaMemcpyDeviceToDevice, st));
        }
        d3f::comm_signal_kernel<<<1, 32, 0, st>>>(everyone, seq);
        g_launches.fetch_add(c->rank == root ? 3 : 2, std::memory_order_relaxed);
    }
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_eval_backward(const D3FObs* obs, const float* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
                      const float* const* grad_out, const float* grad_dist, float* grad_pts,
                      uint32_t flags, float mu, void* stream) {
    NvtxRange nvtx_("d3f_eval_backward");
    if (!obs) return fail(D3F_EINVAL, "obs is NULL");
    if (obs->V < 1 || obs->V > D3F_MAX_VIEWS) return fail(D3F_EINVAL, "V=%d outside 1..%d", obs->V, D3F_MAX_VIEWS);
    if (obs->H < 2 || obs->W < 2) return fail(D3F_EINVAL, "image size %dx%d must be at least 2x2", obs->H, obs->W);
    if (!obs->pose || !obs->K || !obs->depth) return fail(D3F_EINVAL, "obs pose/K/depth pointer is NULL");
    if (n < 0 || (n > 0 && (!pts || !grad_pts))) return fail(D3F_EINVAL, "backward: bad n / NULL pts or grad_pts");
    if (!(mu > 0.f)) return fail(D3F_EINVAL, "mu=%g must be positive", (double)mu);
    if (flags & ~D3F_FLAG_RECIP_NORM) return fail(D3F_EINVAL, "backward: unsupported flag bits 0x%x", flags);
    if (n_keys < 0 || n_keys > D3F_MAX_KEYS || (n_keys > 0 && (!keys || !grad_out)))
        return fail(D3F_EINVAL, "backward: bad keys / grad_out");
    d3f::BwdKeySet ks;
    memset(&ks, 0, sizeof(ks));
    ks.n_keys = n_keys;
    for (int k = 0; k < n_keys; ++k) {
        const int rk = validate_key(keys[k], k);
        if (rk) return rk;
        const KeyStrides ksd = key_strides(keys[k]);
        ks.data[k] = keys[k].data; ks.grad[k] = grad_out[k]; ks.dtype[k] = keys[k].dtype;
        ks.h[k] = keys[k].h; ks.w[k] = keys[k].w; ks.C[k] = keys[k].C;
        ks.sv[k] = ksd.sv; ks.sy[k] = (int32_t)ksd.sy; ks.sx[k] = (int32_t)ksd.sx;
        ks.vec4[k] = keys[k].dtype == D3F_F32 && keys[k].C % 4 == 0 && ((ksd.sv | ksd.sy | ksd.sx) & 3) == 0 &&
                     aligned(keys[k].data, 16) && aligned(grad_out[k], 16);
    }
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return D3F_OK;
    d3f::EvalParams ep;
    ep.pts = pts; ep.depth = obs->depth; ep.pose = obs->pose; ep.K = obs->K; ep.dist = nullptr; ep.valid = nullptr;
    ep.n = n; ep.V = obs->V; ep.H = obs->H; ep.W = obs->W; ep.mu = mu; ep.flags = flags;
    const int64_t blocks = (n + d3f::BWD_WARPS - 1) / d3f::BWD_WARPS;
    if (blocks > 0x7fffffffll) return fail(D3F_EINVAL, "backward: n too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool recip = (flags & D3F_FLAG_RECIP_NORM) != 0;
    if (n <= d3f::BWD_SPLIT_MAX) {               // latency path: one CTA per point, views dealt to its warps
        if (recip) d3f::field_backward_split_kernel<true><<<(unsigned)n, d3f::BWD_WARPS * 32, 0, st>>>(ep, ks, grad_dist, grad_pts);
        else       d3f::field_backward_split_kernel<false><<<(unsigned)n, d3f::BWD_WARPS * 32, 0, st>>>(ep, ks, grad_dist, grad_pts);
    } else if (recip) {
        d3f::field_backward_kernel<true><<<(unsigned)blocks, d3f::BWD_WARPS * 32, 0, st>>>(ep, ks, grad_dist, grad_pts);
    } else {
        d3f::field_backward_kernel<false><<<(unsigned)blocks, d3f::BWD_WARPS * 32, 0, st>>>(ep, ks, grad_dist, grad_pts);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_track_loss_grad(const float* feat, const float* src, const float* dist, const uint8_t* valid,
                        int64_t n, int32_t C, float dist_w, float* g_feat, float* g_dist, float* loss_terms, void* stream) {
    NvtxRange nvtx_("d3f_track_loss_grad");
    if (n < 0 || n >= (1ll << 31) || C < 1) return fail(D3F_EINVAL, "track_loss_grad: n=%lld C=%d", (long long)n, C);
    if (n > 0 && (!feat || !src || !dist || !valid || !g_feat || !g_dist)) return fail(D3F_EINVAL, "track_loss_grad: NULL pointer");
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return D3F_OK;
    const int warps = 8;
    d3f::track_loss_grad_kernel<<<(unsigned)((n + warps - 1) / warps), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, src, dist, valid, (int)n, C, dist_w, g_feat, g_dist, loss_terms);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_track_update(const D3FTrack* t, void* stream) {
    NvtxRange nvtx_("d3f_track_update");
    if (!t) return fail(D3F_EINVAL, "track_update: NULL");
    if (t->n_inst < 0 || t->n_pts < 0) return fail(D3F_EINVAL, "track_update: negative sizes");
    if (t->n_inst > 0 && (!t->t_in || !t->r_in || !t->last_pts)) return fail(D3F_EINVAL, "track_update: NULL pointer");
    if (t->grad_pts && (!t->t_out || !t->r_out || !t->m_t || !t->v_t || !t->m_r || !t->v_r || t->t_out == t->t_in || t->r_out == t->r_in))
        return fail(D3F_EINVAL, "track_update: an update needs moments and output buffers distinct from the inputs");
    if (t->grad_pts && !(t->step >= 1.f)) return fail(D3F_EINVAL, "track_update: step=%g must be >= 1", (double)t->step);
    int rc = check_device();
    if (rc) return rc;
    if (t->n_inst == 0) return D3F_OK;
    d3f::TrackParams tp;
    tp.t_in = t->t_in; tp.r_in = t->r_in; tp.t_out = t->t_out; tp.r_out = t->r_out;
    tp.m_t = t->m_t; tp.v_t = t->v_t; tp.m_r = t->m_r; tp.v_r = t->v_r;
    tp.last_pts = t->last_pts; tp.grad_pts = t->grad_pts; tp.pts = t->pts; tp.I = t->n_inst; tp.P = t->n_pts;
    tp.step = t->step; tp.lr = t->lr; tp.beta1 = t->beta1; tp.beta2 = t->beta2; tp.eps = t->eps; tp.reg_w = t->reg_w;
    tp.bc1 = (float)(1.0 - pow((double)t->beta1, (double)t->step)); tp.bc2s = (float)sqrt(1.0 - pow((double)t->beta2, (double)t->step));
    d3f::track_update_kernel<<<(unsigned)t->n_inst, d3f::TRACK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(tp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_track_step_supported(const D3FObs* obs, const D3FKey* key) {
    if (!obs || !key || !key->data) return 0;
    if (obs->V < 1 || obs->V > d3f::STEP_VIEWS) return 0;
    if (key->dtype != D3F_F32 || key->C < 4 || key->C % 4 != 0 || key->C > d3f::STEP_MAX_C) return 0;
    const KeyStrides ksd = key_strides(*key);
    if (((ksd.sv | ksd.sy | ksd.sx) & 3) != 0 || !aligned(key->data, 16)) return 0;
    return 1;
}

int d3f_track_step(const D3FObs* obs, const D3FKey* key, const float* src, const D3FTrack* t, float dist_w,
                   float* grad_scratch, uint32_t* arrivals, float* loss_terms, uint32_t flags, float mu, void* stream) {
    NvtxRange nvtx_("d3f_track_step");
    if (!obs || !key || !t) return fail(D3F_EINVAL, "track_step: NULL obs / key / track");
    if (obs->H < 2 || obs->W < 2) return fail(D3F_EINVAL, "image size %dx%d must be at least 2x2", obs->H, obs->W);
    if (!obs->pose || !obs->K || !obs->depth) return fail(D3F_EINVAL, "obs pose/K/depth pointer is NULL");
    if (!(mu > 0.f)) return fail(D3F_EINVAL, "mu=%g must be positive", (double)mu);
    if (flags & ~D3F_FLAG_RECIP_NORM) return fail(D3F_EINVAL, "track_step: unsupported flag bits 0x%x", flags);
    const int rk = validate_key(*key, 0);
    if (rk) return rk;
    if (!d3f_track_step_supported(obs, key))
        return fail(D3F_EINVAL, "track_step: needs V <= %d and a float32 map with C %% 4 == 0, C <= %d, strides multiples of 4, "
                    "16-byte aligned (use the four-launch iteration otherwise)", d3f::STEP_VIEWS, d3f::STEP_MAX_C);
    if (t->n_inst < 0 || t->n_pts < 0 || (int64_t)t->n_inst * t->n_pts >= (1ll << 31)) return fail(D3F_EINVAL, "track_step: bad sizes");
    if (t->n_inst > 0 && t->n_pts > 0 &&
        (!src || !grad_scratch || !arrivals || !t->t_in || !t->r_in || !t->t_out || !t->r_out || !t->m_t || !t->v_t || !t->m_r ||
         !t->v_r || !t->last_pts || t->t_out == t->t_in || t->r_out == t->r_in))
        return fail(D3F_EINVAL, "track_step: NULL pointer, or output parameters alias the inputs");
    if (!aligned(src, 16)) return fail(D3F_EINVAL, "track_step: src must be 16-byte aligned");
    if (!(t->step >= 1.f)) return fail(D3F_EINVAL, "track_step: step=%g must be >= 1", (double)t->step);
    int rc = check_device();
    if (rc) return rc;
    const int64_t n = (int64_t)t->n_inst * t->n_pts;
    if (n == 0) return D3F_OK;
    const KeyStrides ksd = key_strides(*key);
    d3f::TrackStepParams sp;
    sp.pose = obs->pose; sp.K = obs->K; sp.depth = obs->depth; sp.V = obs->V; sp.H = obs->H; sp.W = obs->W; sp.mu = mu;
    sp.map = static_cast<const float*>(key->data); sp.h = key->h; sp.w = key->w; sp.C = key->C;
    sp.sv = ksd.sv; sp.sy = (int32_t)ksd.sy; sp.sx = (int32_t)ksd.sx;
    sp.src = src; sp.dist_w = dist_w; sp.grad_pts = grad_scratch; sp.loss_terms = loss_terms; sp.arrivals = arrivals;
    d3f::TrackParams& tp = sp.tp;
    tp.t_in = t->t_in; tp.r_in = t->r_in; tp.t_out = t->t_out; tp.r_out = t->r_out;
    tp.m_t = t->m_t; tp.v_t = t->v_t; tp.m_r = t->m_r; tp.v_r = t->v_r;
    tp.last_pts = t->last_pts; tp.grad_pts = nullptr; tp.pts = t->pts; tp.I = t->n_inst; tp.P = t->n_pts;
    tp.step = t->step; tp.lr = t->lr; tp.beta1 = t->beta1; tp.beta2 = t->beta2; tp.eps = t->eps; tp.reg_w = t->reg_w;
    tp.bc1 = (float)(1.0 - pow((double)t->beta1, (double)t->step)); tp.bc2s = (float)sqrt(1.0 - pow((double)t->beta2, (double)t->step));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // one wave if it can be: two CTAs per SM keep everything in registers, three take 1.5x the points at once
    static const int minb_env = [] { const char* e = getenv("D3F_STEP_MINB"); return e ? atoi(e) : 0; }();            // A/B only
    const int minb = minb_env ? minb_env : (n <= 2 * 148 ? 2 : 3);
    const bool recip = (flags & D3F_FLAG_RECIP_NORM) != 0;
    if (minb == 2) {
        if (recip) d3f::track_step_kernel<true, 2><<<(unsigned)n, d3f::STEP_THREADS, 0, st>>>(sp);
        else       d3f::track_step_kernel<false, 2><<<(unsigned)n, d3f::STEP_THREADS, 0, st>>>(sp);
    } else {
        if (recip) d3f::track_step_kernel<true, 3><<<(unsigned)n, d3f::STEP_THREADS, 0, st>>>(sp);
        else       d3f::track_step_kernel<false, 3><<<(unsigned)n, d3f::STEP_THREADS, 0, st>>>(sp);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_pca_project(const float* x, int64_t n, int32_t C, const float* mean, const float* components,
                    int32_t n_comp, float* y, void* stream) {
    NvtxRange nvtx_("d3f_pca_project");
    if (n < 0 || C < 1 || n_comp < 1 || n_comp > d3f::PCA_MAX_COMP)
        return fail(D3F_EINVAL, "pca: n=%lld C=%d n_comp=%d (n_comp must be 1..%d)", (long long)n, C, n_comp, d3f::PCA_MAX_COMP);
    if (n > 0 && (!x || !components || !y)) return fail(D3F_EINVAL, "pca: NULL pointer");
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return D3F_OK;
    const int warps_per_block = 8;
    const int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
    if (blocks > 0x7fffffffll) return fail(D3F_EINVAL, "pca: n too large");
    d3f::pca_project_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, C, mean, components, n_comp, y);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_create_grid(double x_lower, double y_lower, double z_lower, double step,
                    int32_t nx, int32_t ny, int32_t nz, float* pts, void* stream) {
    NvtxRange nvtx_("d3f_create_grid");
    if (nx < 0 || ny < 0 || nz < 0) return fail(D3F_EINVAL, "grid: negative size");
    const int64_t n = (int64_t)nx * ny * nz;
    if (n > 0 && !pts) return fail(D3F_EINVAL, "grid: pts is NULL");
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return D3F_OK;
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    if (blocks > 0x7fffffffll) return fail(D3F_EINVAL, "grid: too many points");
    d3f::create_grid_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        x_lower, y_lower, z_lower, step, nx, ny, nz, pts);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

}  // extern "C"
