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

#define D3F_ABI_VERSION 2
#define D3F_MAX_VIEWS 16
#define D3F_MAX_KEYS 8
#define D3F_MAX_PEERS 8            /* GPUs of one box that can share a gather (d3f_comm_*) */
#define D3F_IPC_HANDLE_BYTES 64    /* sizeof(cudaIpcMemHandle_t) */

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
#define D3F_ETIMEOUT     -4   /* a peer never signalled (d3f_comm_status) */

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
    const void* data;     /* (V,h,w,C) channels-last; contiguous unless the strides below say otherwise */
    int32_t dtype;        /* D3F_F32 | D3F_U8 */
    int32_t h, w, C;
    int64_t stride_v, stride_y, stride_x;
                          /* element strides of the view / row / texel axes; all three 0 = contiguous
                             (h*w*C, w*C, C).  The channel axis always has stride 1 (what makes the gather
                             coalesced).  Lets a caller pass a crop or a padded map without a copy: the
                             reference samples a permuted *view* of its (V,h,w,C) tensor (fusion.py:373).
                             Strides that are multiples of 4 elements (and C % 4 == 0) keep the 128-bit
                             loads; otherwise the map is read with scalar loads.  One view may span at most
                             2^31 - 1 elements. */
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
 * This is the call bench.py's end-to-end number times.  Synchronous: returns after the last copy
 * has completed (also on error: no copy into the caller's buffers is left in flight).
 * `obs_stream` is the stream the observation was last written on (the call orders itself after it
 * with an event; no device-wide synchronisation).  Thread-safe: concurrent callers use separate
 * scratch (streams + device slabs) from a pool that d3f_release_scratch() frees. */
int d3f_eval_host(const D3FObs* obs, const float* pts_host, int64_t n,
                  const D3FKey* keys, int32_t n_keys,
                  float* dist_host, uint8_t* valid_host,
                  float* const* out_host,
                  uint32_t flags, float mu, void* obs_stream);

/* Free the pooled scratch of d3f_eval_host (device slabs, streams) of the current device. */
int d3f_release_scratch(void);

/* Binned traversal for query points without spatial order (keypoints, mesh vertices; the tracking
 * use of the reference, vis_tracking.py:92-130 / fusion.py:1449,1650).  d3f_bin_order writes to
 * `order` (n int32, device) a permutation that groups points by the cell of a cubic lattice of
 * edge `cell` metres they fall in, cells in Morton (z-curve) order: neighbouring points of the
 * permuted sequence project into the same texel cells of EVERY view, which is what the wide walk's
 * register cache of corner texels needs.  `workspace` is device scratch of at least
 * d3f_bin_workspace_bytes(n) bytes.  Asynchronous on `stream`. */
int64_t d3f_bin_workspace_bytes(int64_t n);
int d3f_bin_order(const float* pts, int64_t n, float cell, int32_t* order,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* d3f_eval visiting the points in the sequence order[0..n) (any permutation of 0..n-1, e.g. from
 * d3f_bin_order): point order[i] is evaluated at step i and its results are written to row
 * order[i], so every output is identical — bit for bit — to d3f_eval's.  out_inter is not
 * supported here. */
int d3f_eval_ordered(const D3FObs* obs, const float* pts, int64_t n, const int32_t* order,
                     const D3FKey* keys, int32_t n_keys,
                     float* dist, uint8_t* valid, float* const* out,
                     uint32_t flags, float mu, void* stream);

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

/* rigid_tracking (reference fusion.py:1608-1685): with d3f_eval and d3f_eval_backward these two calls make one Adam
 * iteration of the reference's loop four launches (d3fields_b200/tracking.py FusedRigidTracker captures 100 of them in
 * one CUDA graph).
 *
 * d3f_track_loss_grad: gradients of  loss = mean_p(|feat_p - src_p|_2 * valid_p) + dist_w * mean_p(max(dist_p*valid_p, 0))
 * (fusion.py:1651-1653) with respect to feat (n,C) and dist (n); loss_terms (n, nullable) receives each point's share. */
int d3f_track_loss_grad(const float* feat, const float* src, const float* dist, const uint8_t* valid,
                        int64_t n, int32_t C, float dist_w,
                        float* g_feat, float* g_dist, float* loss_terms, void* stream);

/* d3f_track_update: per instance, chain grad_pts (n_inst*n_pts,3; from d3f_eval_backward) through
 * pts = last_pts @ so3_exp_map(log_r) + t  (pytorch3d conventions, fusion.py:1646-1648) to t and log_r, add the gradient
 * of reg_w * (|t| + |log_r|) (Frobenius norms over all instances, fusion.py:1654), take Adam step number `step`
 * (torch.optim.Adam defaults) from (t_in, r_in) into (t_out, r_out) — distinct buffers — and write the next iteration's
 * points to pts.  grad_pts == NULL: no update, pts = transform with (t_in, r_in).  pts == NULL: no transform. */
typedef struct D3FTrack {
    const float* t_in;  const float* r_in;      /* (n_inst,3) */
    float* t_out;       float* r_out;           /* (n_inst,3) */
    float* m_t; float* v_t; float* m_r; float* v_r;   /* Adam moments (n_inst,3), updated in place */
    const float* last_pts;                      /* (n_inst,n_pts,3) */
    const float* grad_pts;                      /* (n_inst*n_pts,3) or NULL */
    float* pts;                                 /* (n_inst*n_pts,3) or NULL */
    int32_t n_inst, n_pts;
    float step, lr, beta1, beta2, eps, reg_w;
} D3FTrack;
int d3f_track_update(const D3FTrack* tp, void* stream);

/* d3f_track_step: one whole Adam iteration of rigid_tracking (reference fusion.py:1643-1665) in ONE launch — what the
 * four calls above do, with one CTA per point: points from (last_pts, t_in, r_in), the field query of `key` at them, the
 * loss against src (n_inst*n_pts, C), its gradient through the field back to the points (texels read once, kept in
 * registers for forward and backward), and, by the last CTA of every instance to finish, the Adam step into
 * (t_out, r_out).  t->grad_pts is ignored: grad_scratch (n_inst*n_pts,3) receives d loss / d pts.  t->pts (nullable)
 * receives the points this launch evaluated.  arrivals: n_inst uint32 counters, zero before the first call (every launch
 * leaves them zero).  loss_terms as in d3f_track_loss_grad.
 * Supported (d3f_track_step_supported() != 0): V <= 4, float32 key, C % 4 == 0, C <= 1024, strides multiples of 4,
 * 16-byte aligned map and src; D3F_EINVAL otherwise — use the four-call iteration. */
int d3f_track_step_supported(const D3FObs* obs, const D3FKey* key);
int d3f_track_step(const D3FObs* obs, const D3FKey* key, const float* src, const D3FTrack* t, float dist_w,
                   float* grad_scratch, uint32_t* arrivals, float* loss_terms,
                   uint32_t flags, float mu, void* stream);

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

/* Fused dense sweep: the candidate search of select_features_rand / select_features_from_pcd
 * (reference fusion.py:1420-1445, 1477-1501) and the dense `dist` volume extract_mesh consumes
 * (fusion.py:1321-1322) without the grid or the dense mask field ever existing in HBM.
 *
 * Points: voxel centres of a create_init_grid grid (fusion.py:79-88) taken from the three axis
 * arrays (device; exactly torch.arange(lower, upper, step) + step/2 as the reference computes them,
 * so the coordinates are the reference's bit for bit), linear index i = (ix*ny + iy)*nz + iz
 * (z fastest) — or, when grid is NULL, the n rows of `pts`.
 *
 * Per point: dist / valid_mask as d3f_eval; then, only where valid && |dist| < dist_threshold,
 * the field of the one-hot instance mask `mask_key` (u8 or f32, (V,h,w,num_inst), num_inst <= 32),
 * m = field / (sum_j field_j + 1e-7)  (fusion.py:1439), and the first instance j >= 1 with
 * m_j > mask_threshold (fusion.py:1443; instance 0 is background and is never selected).
 * Survivors are stream-compacted: sel_index[k] = i, sel_inst[k] = j, *sel_count = number found
 * (device; may exceed `capacity`, entries beyond capacity are dropped — call again with more room).
 * The order of the survivors is not deterministic; sort by index to get the reference's order.
 *
 *   dist_out / valid_out   NULL, or dense (n) outputs (what extract_mesh needs)
 *   mask_key               NULL = no selection (dense outputs only)
 *   *sel_count             must be zeroed by the caller (device int64)
 */
typedef struct D3FGrid {
    const float* x;       /* (nx) device */
    const float* y;       /* (ny) device */
    const float* z;       /* (nz) device */
    int32_t nx, ny, nz;
} D3FGrid;

int d3f_sweep_select(const D3FObs* obs, const D3FGrid* grid, const float* pts, int64_t n,
                     const D3FKey* mask_key, float dist_threshold, float mask_threshold,
                     float* dist_out, uint8_t* valid_out,
                     int64_t capacity, int64_t* sel_count, int32_t* sel_index, int32_t* sel_inst,
                     uint32_t flags, float mu, void* stream);

/* ---- Multi-GPU: one process per GPU of one box ------------------------------------------------
 * The reference is single-device (fusion.py:203).  Every query point is independent, so the path
 * shards over the points with no data-path collective; what callers need on every
 * GPU is only the compact field: dist (4 B/point) and valid_mask (1 B/point).  A D3FComm holds one
 * cudaMalloc'ed segment per rank, mapped into every peer with CUDA IPC (NVLink peer access): the
 * field kernel stores each point's dist / valid straight into the gathered arrays of ALL ranks
 * (plain remote st.global, 5 B/point/peer against 4 KB/point of local descriptor rows), and the
 * last CTA of the launch publishes a per-rank epoch flag to every peer (fence.sys + st.release.sys)
 * and waits for theirs (ld.acquire.sys) — the gather costs no second kernel, no NCCL call and no
 * extra stream.  Gathered arrays are double-buffered: the result of call k stays valid until
 * call k+2 is issued.
 *
 * Setup is collective and host-synchronised by the caller (torch.distributed in the mirror):
 *   1. every rank: d3f_comm_create(rank, world, capacity, &comm, handle)      -> 64-byte IPC handle
 *   2. exchange the handles (all-gather of D3F_IPC_HANDLE_BYTES bytes per rank, rank order)
 *   3. every rank: d3f_comm_connect(comm, all_handles); then a host barrier.
 */
typedef struct D3FComm D3FComm;

int d3f_comm_create(int32_t rank, int32_t world, int64_t capacity_points, int64_t staging_bytes,
                    D3FComm** comm, void* handle_out);
int d3f_comm_connect(D3FComm* comm, const void* handles);
int d3f_comm_destroy(D3FComm* comm);
/* 0, or D3F_ETIMEOUT if a wait inside a kernel of this communicator gave up (a peer never arrived).
 * Synchronises `stream` first. */
int d3f_comm_status(D3FComm* comm, void* stream);

/* d3f_eval over this rank's n points + in-kernel all-gather of dist / valid_mask.  Local point i
 * lands at gathered index  gather_base + (i / gather_block) * gather_stride + i % gather_block
 * (contiguous slab: base = slab start, block >= n; blocks dealt round-robin: base = rank*block,
 * stride = world*block), so the gathered arrays come out in canonical point order with no
 * re-ordering pass.  *dist_all / *valid_all receive the LOCAL device addresses of the gathered
 * arrays of this call (capacity_points elements each), complete for all ranks once the launch has
 * finished on `stream`.  Collective: every rank must call it the same number of times. */
int d3f_eval_allgather(D3FComm* comm, const D3FObs* obs, const float* pts, int64_t n,
                       const D3FKey* keys, int32_t n_keys, float* const* out,
                       int64_t gather_base, int64_t gather_block, int64_t gather_stride,
                       uint32_t flags, float mu, void* stream,
                       float** dist_all, uint8_t** valid_all);

/* Replicate `bytes` bytes at `buf` (device) from rank `root` to every rank — the observation after
 * Fusion.update() on one rank (SURVEY.md 8e "replicate once per update()").  Chunks travel through
 * the communicator's staging area with copy-engine P2P writes; flags order them.  Collective,
 * asynchronous on `stream`. */
int d3f_comm_broadcast(D3FComm* comm, void* buf, int64_t bytes, int32_t root, void* stream);

/* Diagnostics */
int         d3f_abi_version(void);
const char* d3f_last_error(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
int64_t     d3f_launch_count(void);
/* name of the kernel variant the last d3f_eval used for keys[k] (static string) */
const char* d3f_last_variant(int32_t k);
/* sizeof(D3FKey) / sizeof(D3FObs) as compiled into the library: the binding checks its struct layout */
int         d3f_sizeof_key(void);
int         d3f_sizeof_obs(void);

#ifdef __cplusplus
}
#endif
#endif /* D3F_H_ */
