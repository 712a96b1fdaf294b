/*
 * d3f.h — C ABI of the B200-native d3fields field query (libd3f.so).
 *
 * The reference (WangYixuan12/d3fields) has no FFI: its boundary for this path is the
 * Python class `Fusion` in fusion.py.  This header is the boundary a native library for
 * that path exposes; each entry point names the reference code it replaces.  The Python
 * mirror of the reference interface (d3fields_b200/fusion.py) binds it with ctypes and
 * passes `tensor.data_ptr()` values — there are no torch types in any signature.
 * INTEGRATION.md shows the binding a maintainer of the reference would add to fusion.py.
 *
 * Conventions
 *   - Every pointer in D3FObs / D3FKey and every pts / output pointer of d3f_eval is a
 *     DEVICE pointer on the current CUDA device, borrowed for the duration of the call;
 *     nothing is cached across calls (Fusion.update() replaces the tensors every frame,
 *     reference fusion.py:707-712).  d3f_eval_host takes HOST pointers for pts / outputs.
 *   - Calls are asynchronous on `stream` (a cudaStream_t; NULL = legacy default stream),
 *     except d3f_eval_host, which returns after the last copy has completed.
 *   - Return value: 0 on success, a negative D3F_E* code otherwise; d3f_last_error()
 *     gives a thread-local message.  Nothing throws across the ABI.
 *   - All maps are channels-last and C-contiguous: (V, h, w, C), exactly the layout of
 *     Fusion.curr_obs_torch (reference fusion.py:618 dino_feats, :1171 mask, :709
 *     color_tensor).  Outputs are always float32.
 */
#ifndef D3F_H_
#define D3F_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3F_ABI_VERSION 1
#define D3F_MAX_VIEWS 16
#define D3F_MAX_KEYS 8

/* element type of a sampled map */
#define D3F_F32 0
#define D3F_U8  1   /* e.g. a uint8 one-hot instance mask; read as its float value */

/* flags */
#define D3F_FLAG_EVAL_DIST     1u  /* Fusion.eval_dist semantics (reference fusion.py:396-436):
                                      no clamp, no `dist > -mu` test, no 1e3 fill; keys ignored */
#define D3F_FLAG_RECIP_NORM    2u  /* pixel normalisation multiplies by 1/(size-1) and un-normalises
                                      as ((c+1)/2)*(size-1): torch's CUDA kernels' rounding instead
                                      of its CPU kernels' (the parity oracle is the CPU path) */

/* error codes */
#define D3F_OK            0
#define D3F_EINVAL       -1   /* bad argument (null pointer, V/C/n out of range, unknown dtype) */
#define D3F_ECUDA        -2   /* a CUDA runtime call failed; see d3f_last_error() */
#define D3F_EUNSUPPORTED -3   /* the device is not sm_100 */

/* The per-frame observation: Fusion.curr_obs_torch['pose'|'K'|'depth'] + Fusion.H/W
 * (reference fusion.py:710-714). */
typedef struct D3FObs {
    int32_t V;            /* number of views (Fusion.num_cam), 1..D3F_MAX_VIEWS */
    int32_t H, W;         /* image size: pixel coords are normalised by (W-1),(H-1) whatever the
                             map size (reference fusion.py:72-73, 329-330, 375-376) */
    const float* pose;    /* (V,3,4) world->camera [R|t] */
    const float* K;       /* (V,3,3) intrinsics */
    const float* depth;   /* (V,H,W) metres; 0 = hole */
} D3FObs;

/* One sampled map of return_names: Fusion.curr_obs_torch[name] (reference fusion.py:372-379). */
typedef struct D3FKey {
    const void* data;     /* (V,h,w,C) channels-last, contiguous */
    int32_t dtype;        /* D3F_F32 | D3F_U8 */
    int32_t h, w, C;
    const float* bias;    /* NULL, or (C) device floats subtracted from every output row (C < 128 only).
                             Used for PCA'd descriptor fields: because the field is linear in the sampled map,
                             (field - mean) @ W^T == field_of(map @ W^T) - mean @ W^T, so the map is projected
                             once with d3f_pca_project and queried as a C = n_comp key with bias = mean @ W^T
                             (reference fusion.py:1386-1392 does the projection on the host, after eval). */
} D3FKey;

/* Replaces Fusion.eval (reference fusion.py:305-394), and with D3F_FLAG_EVAL_DIST
 * Fusion.eval_dist (fusion.py:396-436), including the helpers they call:
 * project_points_coords (fusion.py:32-55) and interpolate_feats (fusion.py:57-77).
 *
 *   pts        (n,3) float32 world points
 *   dist       (n)   float32  out: truncated signed distance, 1e3 where no view sees the point
 *   valid      (n)   uint8    out: 1 where at least one view sees the point (torch.bool storage)
 *   out[k]     (n,C_k) float32 out: visibility-weighted mean of the bilinear samples of keys[k]
 *   out_inter  NULL, or n_keys pointers to (V,n,C_k) float32: the per-view bilinear samples
 *              (return_inter=True, reference fusion.py:389-390); entries may be NULL
 *   mu         truncation distance (Fusion.mu, reference fusion.py:208)
 *
 * Batching (Fusion.batch_eval, reference fusion.py:526-545) needs no entry point of its own:
 * the (V,n,C) temporaries that force the reference to chunk never exist here, so batch_eval
 * is one d3f_eval call over all n points.  n may be 0.
 */
int d3f_eval(const D3FObs* obs, const float* pts, int64_t n,
             const D3FKey* keys, int32_t n_keys,
             float* dist, uint8_t* valid,
             float* const* out, float* const* out_inter,
             uint32_t flags, float mu, void* stream);

/* Same computation with HOST pts / outputs (observation and maps stay device-resident, as
 * after Fusion.update()).  Points are uploaded and results downloaded in slabs, copies
 * overlapped with the kernels on internal streams; pinned host memory gives full PCIe rate.
 * This is the call bench.py's end-to-end number times.  Synchronous. */
int d3f_eval_host(const D3FObs* obs, const float* pts_host, int64_t n,
                  const D3FKey* keys, int32_t n_keys,
                  float* dist_host, uint8_t* valid_host,
                  float* const* out_host,
                  uint32_t flags, float mu);

/* Gradient of d3f_eval's outputs with respect to the query points: what torch autograd computes when the
 * reference's rigid_tracking back-propagates through Fusion.eval (reference fusion.py:1643-1665).
 *   grad_out[k]  (n,C_k) upstream gradient of out[k], or NULL for a key that does not need one
 *   grad_dist    (n) upstream gradient of dist, or NULL
 *   grad_pts     (n,3) out
 * Differentiable terms: the bilinear samples through their pixel coordinates, the distance weight and the
 * clamped dist through the camera-frame depth; visibility masks, nearest-depth lookups and the 1e3 fill are
 * constants, exactly as in torch.  D3F_FLAG_EVAL_DIST is not supported. */
int d3f_eval_backward(const D3FObs* obs, const float* pts, int64_t n,
                      const D3FKey* keys, int32_t n_keys,
                      const float* const* grad_out, const float* grad_dist,
                      float* grad_pts, uint32_t flags, float mu, void* stream);

/* Fused PCA projection of a descriptor field: y = (x - mean) @ components^T, the
 * sklearn.decomposition.PCA.transform the reference applies on the host to eval's
 * 'dino_feats' (reference fusion.py:1386-1392, weights from pca_model/*.pkl).
 *   x (n,C) device, mean (C) device or NULL (no centring), components (n_comp,C) device, y (n,n_comp) device. */
int d3f_pca_project(const float* x, int64_t n, int32_t C,
                    const float* mean, const float* components, int32_t n_comp,
                    float* y, void* stream);

/* Voxel-centre grid, z fastest (reference fusion.py:79-88 create_init_grid): writes
 * pts (nx*ny*nz, 3), coordinate i of an axis = float(lower + step*i) + float(step/2): torch.arange's
 * float32 value (computed in double, rounded once) plus the half step. */
int d3f_create_grid(double x_lower, double y_lower, double z_lower, double step,
                    int32_t nx, int32_t ny, int32_t nz, float* pts, void* stream);

/* Diagnostics */
int         d3f_abi_version(void);
const char* d3f_last_error(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
int64_t     d3f_launch_count(void);
/* name of the kernel variant the last d3f_eval used for keys[k] (static string) */
const char* d3f_last_variant(int32_t k);

#ifdef __cplusplus
}
#endif
#endif /* D3F_H_ */
