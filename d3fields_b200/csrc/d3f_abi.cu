// d3f_abi.cu — the C ABI of libd3f.so (include/d3f.h): argument checking, kernel selection and
// launch.  No torch, no host-side math on tensors; every entry point is asynchronous on the
// caller's stream except d3f_eval_host.
#include <cuda_runtime.h>
#include <atomic>
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

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
thread_local const char* g_variant[D3F_MAX_KEYS] = {};

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

int validate(const D3FObs* obs, const void* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
             const void* dist, const void* valid, float* const* out, uint32_t flags, float mu) {
    if (!obs) return fail(D3F_EINVAL, "obs is NULL");
    if (obs->V < 1 || obs->V > D3F_MAX_VIEWS) return fail(D3F_EINVAL, "V=%d outside 1..%d", obs->V, D3F_MAX_VIEWS);
    if (obs->H < 2 || obs->W < 2) return fail(D3F_EINVAL, "image size %dx%d must be at least 2x2", obs->H, obs->W);
    if (!obs->pose || !obs->K || !obs->depth) return fail(D3F_EINVAL, "obs pose/K/depth pointer is NULL");
    if (n < 0) return fail(D3F_EINVAL, "n=%lld is negative", (long long)n);
    if (n > 0 && (!pts || !dist || !valid)) return fail(D3F_EINVAL, "pts/dist/valid pointer is NULL");
    if (!(mu > 0.f)) return fail(D3F_EINVAL, "mu=%g must be positive", (double)mu);
    if (flags & ~(D3F_FLAG_EVAL_DIST | D3F_FLAG_RECIP_NORM)) return fail(D3F_EINVAL, "unknown flag bits 0x%x", flags);
    if (n_keys < 0 || n_keys > D3F_MAX_KEYS) return fail(D3F_EINVAL, "n_keys=%d outside 0..%d", n_keys, D3F_MAX_KEYS);
    if (n_keys > 0 && (!keys || !out)) return fail(D3F_EINVAL, "keys/out is NULL with n_keys=%d", n_keys);
    for (int k = 0; k < n_keys; ++k) {
        const D3FKey& q = keys[k];
        if (!q.data) return fail(D3F_EINVAL, "keys[%d].data is NULL", k);
        if (q.dtype != D3F_F32 && q.dtype != D3F_U8) return fail(D3F_EINVAL, "keys[%d].dtype=%d unknown", k, q.dtype);
        if (q.h < 1 || q.w < 1 || q.C < 1) return fail(D3F_EINVAL, "keys[%d] shape (%d,%d,%d) invalid", k, q.h, q.w, q.C);
        if ((int64_t)q.h * q.w >= (1ll << 31)) return fail(D3F_EINVAL, "keys[%d] map too large", k);
        if (n > 0 && !out[k]) return fail(D3F_EINVAL, "out[%d] is NULL", k);
        if (q.bias && q.C >= 128) return fail(D3F_EINVAL, "keys[%d].bias is supported for C < 128 only", k);
        if (q.bias && q.C % 4 == 0 && !aligned(q.bias, 16)) return fail(D3F_EINVAL, "keys[%d].bias must be 16-byte aligned", k);
    }
    return D3F_OK;
}

// Device-pointer evaluation shared by d3f_eval and the slabs of d3f_eval_host.
int launch_eval(const D3FObs* obs, const float* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
                float* dist, uint8_t* valid, float* const* out, float* const* out_inter,
                uint32_t flags, float mu, cudaStream_t st) {
    if (n == 0) return D3F_OK;
    d3f::EvalParams ep;
    ep.pts = pts; ep.depth = obs->depth; ep.pose = obs->pose; ep.K = obs->K;
    ep.dist = dist; ep.valid = valid; ep.n = n; ep.V = obs->V; ep.H = obs->H; ep.W = obs->W;
    ep.mu = mu; ep.flags = flags;
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
        ks.dtype[k] = keys[k].dtype;
        ks.k[k].bias = keys[k].bias;
        any_inter |= ks.k[k].inter != nullptr;
        if (keys[k].C % 4 == 0) {
            const size_t a_in = keys[k].dtype == D3F_F32 ? 16 : 4;
            if (!aligned(keys[k].data, a_in) || !aligned(out[k], 16) || (ks.k[k].inter && !aligned(ks.k[k].inter, 16)))
                return fail(D3F_EINVAL, "keys[%d]: map/out pointers must be 16-byte aligned when C %% 4 == 0", k);
        }
        g_variant[k] = "generic";
    }
    // production path: 128-point tiles, register-cached corner texels for wide float32 maps
    bool tile_ok = obs->V <= d3f::TILE_V && !any_inter;
    for (int k = 0; k < ks.n_keys; ++k) tile_ok &= ((int64_t)keys[k].h * keys[k].w < (1ll << 29));
    if (tile_ok) {
        const int64_t tiles = (n + d3f::TILE_PTS - 1) / d3f::TILE_PTS;
        if (tiles > 0x7fffffffll) return fail(D3F_EINVAL, "n=%lld too large for one launch", (long long)n);
        for (int k = 0; k < ks.n_keys; ++k)
            g_variant[k] = d3f::key_is_wide(keys[k].dtype, keys[k].C, keys[k].h, keys[k].w) ? "tile/wide" : "tile/narrow";
        dim3 grid((unsigned)tiles), block(d3f::TILE_THREADS);
        // L1 lookahead prefetch of cell changes pays only when the wide volume cannot stay L2-resident
        // (measured: cfg2b 5 GB volume 2.02 -> 1.87 ms; cfg2a 50 MB volume 0.80 -> 0.88 ms).  D3F_TILE_PREFETCH=0/1 overrides.
        static const int force = [] { const char* e = getenv("D3F_TILE_PREFETCH"); return e ? atoi(e) : -1; }();
        size_t wide_bytes = 0;
        for (int k = 0; k < ks.n_keys; ++k)
            if (d3f::key_is_wide(keys[k].dtype, keys[k].C, keys[k].h, keys[k].w))
                wide_bytes += (size_t)obs->V * keys[k].h * keys[k].w * keys[k].C * 4;
        const bool prefetch = force >= 0 ? force != 0 : wide_bytes > (size_t)(64u << 20);
        if (wide_bytes == 0) {            // no register-cached walk needed: the light instantiation (4 CTAs/SM)
            if (recip) d3f::field_tile_kernel<true, 0, false><<<grid, block, 0, st>>>(ep, ks);
            else       d3f::field_tile_kernel<false, 0, false><<<grid, block, 0, st>>>(ep, ks);
        } else if (recip) {
            if (prefetch) d3f::field_tile_kernel<true, 4, true><<<grid, block, 0, st>>>(ep, ks);
            else          d3f::field_tile_kernel<true, 0, true><<<grid, block, 0, st>>>(ep, ks);
        } else {
            if (prefetch) d3f::field_tile_kernel<false, 4, true><<<grid, block, 0, st>>>(ep, ks);
            else          d3f::field_tile_kernel<false, 0, true><<<grid, block, 0, st>>>(ep, ks);
        }
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

// ---- scratch for d3f_eval_host: streams + device slabs, grown on demand, kept until exit -----
constexpr int HOST_STREAMS = 3;
struct HostScratch {
    int dev = -1;
    cudaStream_t st[HOST_STREAMS] = {};
    void* buf[HOST_STREAMS] = {};
    size_t cap[HOST_STREAMS] = {};
};
std::mutex g_scratch_mu;
HostScratch g_scratch;

}  // namespace

extern "C" {

int d3f_abi_version(void) { return D3F_ABI_VERSION; }
const char* d3f_last_error(void) { return g_err; }
int64_t d3f_launch_count(void) { return g_launches.load(); }
const char* d3f_last_variant(int32_t k) { return (k >= 0 && k < D3F_MAX_KEYS && g_variant[k]) ? g_variant[k] : ""; }

int d3f_eval(const D3FObs* obs, const float* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
             float* dist, uint8_t* valid, float* const* out, float* const* out_inter,
             uint32_t flags, float mu, void* stream) {
    int rc = validate(obs, pts, n, keys, n_keys, dist, valid, out, flags, mu);
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    return launch_eval(obs, pts, n, keys, n_keys, dist, valid, out, out_inter, flags, mu,
                       static_cast<cudaStream_t>(stream));
}

int d3f_eval_host(const D3FObs* obs, const float* pts_host, int64_t n, const D3FKey* keys, int32_t n_keys,
                  float* dist_host, uint8_t* valid_host, float* const* out_host, uint32_t flags, float mu) {
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

    std::lock_guard<std::mutex> lock(g_scratch_mu);
    int dev = 0;
    D3F_CUDA(cudaGetDevice(&dev));
    HostScratch& sc = g_scratch;
    if (sc.dev != dev) {
        for (int i = 0; i < HOST_STREAMS; ++i) {
            if (sc.buf[i]) cudaFree(sc.buf[i]);
            sc.buf[i] = nullptr; sc.cap[i] = 0;
            if (sc.st[i]) cudaStreamDestroy(sc.st[i]);
            D3F_CUDA(cudaStreamCreateWithFlags(&sc.st[i], cudaStreamNonBlocking));
        }
        sc.dev = dev;
    }
    for (int i = 0; i < HOST_STREAMS; ++i) {
        if (sc.cap[i] < need) {
            if (sc.buf[i]) D3F_CUDA(cudaFree(sc.buf[i]));
            sc.buf[i] = nullptr; sc.cap[i] = 0;
            D3F_CUDA(cudaMalloc(&sc.buf[i], need));
            sc.cap[i] = need;
        }
    }
    // the observation may have been written on another stream by the caller: make it visible
    D3F_CUDA(cudaDeviceSynchronize());

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
        rc = launch_eval(obs, d_pts, m, keys, nk, d_dist, d_valid, d_out, nullptr, flags, mu, st);
        if (rc) return rc;
        D3F_CUDA(cudaMemcpyAsync(dist_host + s0, d_dist, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        D3F_CUDA(cudaMemcpyAsync(valid_host + s0, d_valid, (size_t)m, cudaMemcpyDeviceToHost, st));
        for (int k = 0; k < nk; ++k)
            D3F_CUDA(cudaMemcpyAsync(out_host[k] + (size_t)s0 * keys[k].C, d_out[k], (size_t)m * keys[k].C * 4,
                                     cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < HOST_STREAMS; ++i) D3F_CUDA(cudaStreamSynchronize(sc.st[i]));
    return D3F_OK;
}

int d3f_eval_backward(const D3FObs* obs, const float* pts, int64_t n, const D3FKey* keys, int32_t n_keys,
                      const float* const* grad_out, const float* grad_dist, float* grad_pts,
                      uint32_t flags, float mu, void* stream) {
    if (!obs) return fail(D3F_EINVAL, "obs is NULL");
    if (obs->V < 1 || obs->V > D3F_MAX_VIEWS) return fail(D3F_EINVAL, "V=%d outside 1..%d", obs->V, D3F_MAX_VIEWS);
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
        if (!keys[k].data || (keys[k].dtype != D3F_F32 && keys[k].dtype != D3F_U8) || keys[k].h < 1 || keys[k].w < 1 || keys[k].C < 1)
            return fail(D3F_EINVAL, "backward: keys[%d] invalid", k);
        ks.data[k] = keys[k].data; ks.grad[k] = grad_out[k]; ks.dtype[k] = keys[k].dtype;
        ks.h[k] = keys[k].h; ks.w[k] = keys[k].w; ks.C[k] = keys[k].C;
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
    if (flags & D3F_FLAG_RECIP_NORM)
        d3f::field_backward_kernel<true><<<(unsigned)blocks, d3f::BWD_WARPS * 32, 0, st>>>(ep, ks, grad_dist, grad_pts);
    else
        d3f::field_backward_kernel<false><<<(unsigned)blocks, d3f::BWD_WARPS * 32, 0, st>>>(ep, ks, grad_dist, grad_pts);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    D3F_CUDA(cudaGetLastError());
    return D3F_OK;
}

int d3f_pca_project(const float* x, int64_t n, int32_t C, const float* mean, const float* components,
                    int32_t n_comp, float* y, void* stream) {
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
