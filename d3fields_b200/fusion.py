"""Drop-in mirror of the reference's field-query interface, backed by libd3f.so.

Mirrors, for the hot path only, what reference fusion.py exposes and its drivers import
(`from fusion import Fusion, create_init_grid`, reference vis_repr.py:13):

  create_init_grid(boundaries, step_size)              reference fusion.py:79-88
  project_points_coords(pts, Rt, K)                    reference fusion.py:32-55   (thin view over eval's kernel)
  Fusion(num_cam, feat_backbone, device, dtype)        reference fusion.py:203
  Fusion.update(obs)                                   reference fusion.py:686-714
  Fusion.eval(pts, return_names, return_inter)         reference fusion.py:305-394
  Fusion.eval_dist(pts)                                reference fusion.py:396-436
  Fusion.batch_eval(pts, return_names)                 reference fusion.py:526-545
  Fusion.text_queries_for_inst_mask(_no_track)(...)    reference fusion.py:1173 / :1112 (delegated, see below)
  Fusion.get_inst_num()                                reference fusion.py:1258
  Fusion.select_features_rand / _from_pcd(...)         reference fusion.py:1418 / :1477 (fused sweep + device FPS + eval)
  Fusion.rigid_tracking(...)                           reference fusion.py:1608 (tracking.FusedRigidTracker, one CUDA graph)
  Fusion.curr_obs_torch                                reference fusion.py:210-215, 707-714

Same names, argument meaning, return dict and error behaviour.  What differs:

  * eval/eval_dist/batch_eval launch hand-written sm_100a kernels through the C ABI
    (include/d3f.h); torch only owns the device memory and the stream.  There is no CPU
    fallback: without the library or a CUDA device the call raises.
  * The perception front-end (DINOv2 / GroundingDINO / SAM / XMem, reference fusion.py:223-286)
    is out of scope.  update() takes precomputed features
    (obs['dino_feats']) or a user-supplied `feature_extractor`; the mask methods delegate to a
    user-supplied `perception` object, and set_instance_masks() injects masks directly.
  * 'mask' may be stored as uint8 one-hot (4x less traffic); outputs are float32 either way.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _native


def create_init_grid(boundaries, step_size):
    """Voxel-centre grid over `boundaries`, z fastest; returns (coords (N,3) f32 on CPU, grid shape).
    Same construction as reference fusion.py:79-88 (torch.arange in float32, + step/2, ij meshgrid)."""
    axes = []
    for a in ('x', 'y', 'z'):
        lo, hi = boundaries[a + '_lower'], boundaries[a + '_upper']
        axes.append(torch.arange(lo, hi, step_size, dtype=torch.float32) + step_size / 2)
    gx, gy, gz = torch.meshgrid(*axes, indexing='ij')
    return torch.stack([gx, gy, gz], dim=-1).reshape(-1, 3), gx.shape


class _FieldQuery(torch.autograd.Function):
    """Fusion.eval as a differentiable function of the query points (the reference gets this from torch autograd;
    its rigid_tracking optimises an SE(3) pose through eval, fusion.py:1643-1665).  Backward is one launch of
    d3f_eval_backward; the observation is a constant."""

    @staticmethod
    def forward(ctx, pts, fusion, names):
        res = fusion._run(pts.detach(), names, False, False)
        ctx.fusion, ctx.names = fusion, tuple(names)
        ctx.obs = {k: fusion.curr_obs_torch[k] for k in ('pose', 'K', 'depth') + tuple(names)}   # what forward saw
        ctx.mu, ctx.flags = float(fusion.mu), fusion._flags(False)
        ctx.save_for_backward(pts.detach())
        ctx.mark_non_differentiable(res['valid_mask'])
        return (res['dist'], res['valid_mask']) + tuple(res[k] for k in names)

    @staticmethod
    def backward(ctx, g_dist, _g_valid, *g_feats):
        (pts,) = ctx.saved_tensors
        pts = pts.contiguous()
        o = ctx.obs
        depth = o['depth']
        V, H, W = (int(x) for x in depth.shape)
        dev = depth.device
        keys, grads = [], []
        for name, g in zip(ctx.names, g_feats):
            vol = o[name]
            if vol.dtype == torch.bool:
                vol = vol.view(torch.uint8)
            dt = _native.D3F_F32 if vol.dtype == torch.float32 else _native.D3F_U8
            strides = None if vol.is_contiguous() else tuple(int(x) for x in vol.stride()[:3])
            keys.append((vol.data_ptr(), dt, int(vol.shape[1]), int(vol.shape[2]), int(vol.shape[3]), None, strides))
            grads.append(None if g is None else g.contiguous().float())
        gd = None if g_dist is None else g_dist.contiguous().float()
        grad_pts = torch.empty_like(pts)
        with torch.cuda.device(dev):
            _native.eval_backward(V, H, W, o['pose'].data_ptr(), o['K'].data_ptr(), depth.data_ptr(), pts.data_ptr(),
                                  int(pts.shape[0]), keys, [None if g is None else g.data_ptr() for g in grads],
                                  None if gd is None else gd.data_ptr(), grad_pts.data_ptr(), ctx.flags, ctx.mu,
                                  torch.cuda.current_stream(dev).cuda_stream)
        return grad_pts, None, None


def create_init_grid_device(boundaries, step_size, device='cuda:0'):
    """create_init_grid (reference fusion.py:79-88) generated on the device by d3f_create_grid: the 101 M-point sweep
    of select_features_rand (fusion.py:1420-1424) then never exists on the host nor crosses PCIe.  Same point count
    and order as create_init_grid; coordinates are float(lower + step*i) + float(step/2), which can differ from
    torch.arange's vectorised CPU values by one ulp."""
    import math
    dev = torch.device(device)
    dims = []
    for a in ('x', 'y', 'z'):
        lo, hi = float(boundaries[a + '_lower']), float(boundaries[a + '_upper'])
        dims.append(max(int(math.ceil((hi - lo) / float(step_size))), 0))
    nx, ny, nz = dims
    pts = torch.empty((nx * ny * nz, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _native.create_grid(float(boundaries['x_lower']), float(boundaries['y_lower']), float(boundaries['z_lower']),
                            float(step_size), nx, ny, nz, pts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return pts, torch.Size(dims)


def _as_device(t, device, dtype=None):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device=device)


class Fusion:
    """The per-frame multi-view field (reference fusion.py:202) with a CUDA-native query path.

    index_rounding — which of torch's two roundings of the pixel normalisation the integer decisions replay:
      'cpu' (default)  x_n = p / (W-1) * 2 - 1, index = (x_n + 1) * ((size-1)/2): torch's CPU kernels.  This is the parity
                       oracle: the unmodified reference was run on CPU to produce tests/golden/, and dist / valid_mask
                       match it bit for bit.
      'cuda'           x_n = p * (1/(W-1)) * 2 - 1, index = ((x_n + 1)/2) * (size-1): torch's CUDA kernels, i.e. what a
                       user of the reference on its default device ('cuda:0', fusion.py:203) gets.  The two differ only
                       for points within an ulp of a pixel/texel boundary (a few per million: nearest-depth pixel, bilinear
                       cell, hence visibility); set it when comparing against a GPU run of the reference."""

    def __init__(self, num_cam, feat_backbone='dinov2', device='cuda:0', dtype=torch.float32,
                 feature_extractor: Optional[Callable] = None, perception=None):
        if dtype != torch.float32:
            raise ValueError('the field query computes in float32 (the reference default, fusion.py:203)')
        self.device = device
        self.dtype = dtype
        self.num_cam = num_cam
        self.feat_backbone = feat_backbone
        self.mu = 0.02                       # reference fusion.py:208
        self.curr_obs_torch: Dict[str, object] = {}
        self.H = self.W = -1
        self.feature_extractor = feature_extractor
        self.perception = perception
        self.index_rounding = 'cpu'          # 'cpu' (parity oracle) | 'cuda' (torch CUDA kernels' rounding)

    # ------------------------------------------------------------------ state ------------
    def update(self, obs):
        """obs: 'color' (V,H,W,3) uint8, 'depth' (V,H,W), 'pose' (V,3,4) world->camera, 'K' (V,3,3);
        optionally 'dino_feats' (V,h,w,C) precomputed.  Layout written: reference fusion.py:707-714."""
        depth = obs['depth']
        self.num_cam = int(depth.shape[0])
        color = obs.get('color')
        if 'dino_feats' in obs:
            feats = _as_device(obs['dino_feats'], self.device, self.dtype).contiguous()
        elif self.feature_extractor is not None:
            params = {'patch_h': color.shape[1] // 10, 'patch_w': color.shape[2] // 10}   # fusion.py:694-697
            feats = _as_device(self.feature_extractor(color, params), self.device, self.dtype).contiguous()
        else:
            feats = None
        if feats is not None:
            self.curr_obs_torch['dino_feats'] = feats
        if color is not None:
            self.curr_obs_torch['color'] = color
            self.curr_obs_torch['color_tensor'] = (_as_device(color, self.device, self.dtype) / 255.0).contiguous()
        pose = _as_device(obs['pose'], self.device, self.dtype)
        if pose.shape[-2:] == (4, 4):
            pose = pose[:, :3]
        self.curr_obs_torch['depth'] = _as_device(depth, self.device, self.dtype).contiguous()
        self.curr_obs_torch['pose'] = pose.contiguous()
        self.curr_obs_torch['K'] = _as_device(obs['K'], self.device, self.dtype).contiguous()
        _, self.H, self.W = depth.shape

    def set_instance_masks(self, masks, labels: Optional[Sequence[str]] = None, as_uint8: bool = False):
        """Write curr_obs_torch['mask'] directly: `masks` is a (V,H,W) uint8 label image or a
        (V,H,W,num_inst) one-hot (reference layout, fusion.py:1171).  Stands in for the
        Grounded-SAM/XMem front-end on synthetic or precomputed masks."""
        m = _as_device(masks, self.device)
        if m.dim() == 3:
            num = int(m.max().item()) + 1 if labels is None else len(labels)
            m = (m.unsqueeze(-1) == torch.arange(num, device=m.device, dtype=m.dtype))
        m = m.to(torch.uint8 if as_uint8 else self.dtype).contiguous()
        self.curr_obs_torch['mask'] = m
        if labels is not None:
            self.curr_obs_torch['consensus_mask_label'] = list(labels)

    def get_inst_num(self):
        return self.curr_obs_torch['mask'].shape[-1]          # reference fusion.py:1258-1259

    def text_queries_for_inst_mask_no_track(self, queries, thresholds, boundaries, use_sam=False,
                                            merge_all=False, expected_labels=None, robot_pcd=None, **kw):
        return self._perceive('text_queries_for_inst_mask_no_track', queries, thresholds, boundaries,
                              use_sam=use_sam, merge_all=merge_all, expected_labels=expected_labels,
                              robot_pcd=robot_pcd, **kw)

    def text_queries_for_inst_mask(self, queries, thresholds, boundaries, use_sam=False,
                                   merge_all=False, expected_labels=None, robot_pcd=None, **kw):
        return self._perceive('text_queries_for_inst_mask', queries, thresholds, boundaries,
                              use_sam=use_sam, merge_all=merge_all, expected_labels=expected_labels,
                              robot_pcd=robot_pcd, **kw)

    def _perceive(self, method, *a, **kw):
        if self.perception is None:
            raise RuntimeError(
                f'{method}: the Grounded-SAM / XMem front-end is outside this package (SURVEY.md §2 rows 10-12); '
                'pass perception=<object with this method> to Fusion(...), or call set_instance_masks()')
        masks, labels = getattr(self.perception, method)(self, *a, **kw)
        self.set_instance_masks(masks, labels)

    # ------------------------------------------------------------------ query ------------
    def _check_pts(self, pts):
        if len(self.curr_obs_torch) == 0:           # reference fusion.py:313-317
            print('Please call update() first!')
            exit()
        assert type(pts) == torch.Tensor            # reference fusion.py:318-320
        assert len(pts.shape) == 2
        assert pts.shape[1] == 3
        if pts.dtype != torch.float32:
            raise ValueError(f'pts must be float32, got {pts.dtype}')

    def _obs_ptrs(self):
        o = self.curr_obs_torch
        depth, pose, K = o['depth'], o['pose'], o['K']
        for name, t in (('depth', depth), ('pose', pose), ('K', K)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError(f"curr_obs_torch['{name}'] must be a contiguous float32 CUDA tensor")
        V, H, W = depth.shape
        if tuple(pose.shape) != (V, 3, 4) or tuple(K.shape) != (V, 3, 3):
            raise ValueError(f'pose/K shapes {tuple(pose.shape)}/{tuple(K.shape)} do not match depth {tuple(depth.shape)}')
        if (H, W) != (self.H, self.W):
            raise ValueError(f'depth is {H}x{W} but Fusion.H/W = {self.H}x{self.W}')
        return V, H, W, pose.data_ptr(), K.data_ptr(), depth.data_ptr()

    def _key_tuple(self, name, V):
        vol = self.curr_obs_torch[name]              # KeyError for unknown names, as in the reference (fusion.py:373)
        if not isinstance(vol, torch.Tensor) or vol.dim() != 4 or vol.shape[0] != V:
            raise ValueError(f"curr_obs_torch['{name}'] must be a (V,h,w,C) tensor")
        if vol.dtype == torch.bool:
            vol = vol.view(torch.uint8)
        if vol.dtype not in (torch.float32, torch.uint8):
            raise ValueError(f"curr_obs_torch['{name}'] must be float32 or uint8, got {vol.dtype}")
        if not vol.is_cuda:
            raise ValueError(f"curr_obs_torch['{name}'] must be a CUDA tensor")
        dt = _native.D3F_F32 if vol.dtype == torch.float32 else _native.D3F_U8
        strides = None
        if not vol.is_contiguous():
            # a crop / padded map / permuted view is sampled in place through D3FKey's strides (the reference samples
            # a permuted view of its tensor too, fusion.py:373); only the channel axis must be dense
            sv, sy, sx, sc = (int(x) for x in vol.stride())
            if (sc != 1 and vol.shape[3] != 1) or min(sv, sy, sx) < 0:
                raise ValueError(f"curr_obs_torch['{name}'] must be channels-last (V,h,w,C) with unit channel stride; "
                                 f'got strides {tuple(vol.stride())}')
            strides = (sv, sy, sx)
        return (vol.data_ptr(), dt, int(vol.shape[1]), int(vol.shape[2]), int(vol.shape[3]), None, strides), vol

    def _flags(self, eval_dist=False):
        f = _native.FLAG_EVAL_DIST if eval_dist else 0
        if self.index_rounding == 'cuda':
            f |= _native.FLAG_RECIP_NORM
        elif self.index_rounding != 'cpu':
            raise ValueError("index_rounding must be 'cpu' or 'cuda'")
        return f

    def bin_order(self, pts, cell=None):
        """Visiting order (int32 permutation, device) that groups `pts` (n,3) by lattice cell in Morton order
        (d3f_bin_order) — for keypoints / mesh vertices, which arrive without the z-fastest locality of a
        create_init_grid grid.  `cell` defaults to ~1/128 of the cloud's extent, at least 2.5 mm."""
        n = int(pts.shape[0])
        dev = pts.device
        with torch.cuda.device(dev):
            order = torch.empty(n, dtype=torch.int32, device=dev)
            if n == 0:
                return order
            if cell is None:
                cell = 0.0025
            nbytes = _native.bin_workspace_bytes(n)
            ws = getattr(self, '_bin_ws', None)
            if ws is None or ws.numel() < nbytes or ws.device != dev:
                ws = self._bin_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _native.bin_order(pts.data_ptr(), n, float(cell), order.data_ptr(), ws.data_ptr(), ws.numel(),
                              torch.cuda.current_stream(dev).cuda_stream)
        return order

    def _run(self, pts, return_names, return_inter, eval_dist, out=None, binned=False, gather=None):
        self._check_pts(pts)
        names = list(return_names)
        V, H, W, pose_p, K_p, depth_p = self._obs_ptrs()
        keys, keep = [], []
        for k in ([] if eval_dist else names):
            kt, vol = self._key_tuple(k, V)
            keys.append(kt)
            keep.append(vol)
        n = int(pts.shape[0])
        if not pts.is_cuda:
            return self._run_host(pts, names, keys, V, H, W, pose_p, K_p, depth_p, eval_dist, out=out)
        dev = self.curr_obs_torch['depth'].device
        if pts.device != dev:
            raise ValueError(f'pts is on {pts.device} but the observation is on {dev}')
        pts = pts.contiguous()
        with torch.cuda.device(dev):
            def buf(name, shape, dtype):
                if out is not None and name in out:          # caller-owned output (e.g. a slot of a gather buffer)
                    t = out[name]
                    if tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous() or t.device != dev:
                        raise ValueError(f"out['{name}'] must be a contiguous {dtype} tensor of shape {tuple(shape)} on {dev}")
                    return t
                return torch.empty(shape, dtype=dtype, device=dev)

            outs = [buf(k, (n, kt[4]), torch.float32) for k, kt in zip(names, keys)]
            stream = torch.cuda.current_stream(dev).cuda_stream
            if gather is not None:
                # in-kernel all-gather of dist / valid_mask into every rank's gathered arrays (d3f_eval_allgather)
                if return_inter or binned:
                    raise ValueError('a gathering launch supports neither return_inter nor binned')
                comm, base, block, stride = gather
                dist, valid = comm.eval_allgather(V, H, W, pose_p, K_p, depth_p, pts.data_ptr(), n, keys,
                                                  [o.data_ptr() for o in outs], base, block, stride,
                                                  self._flags(eval_dist), float(self.mu), stream)
                inters = None
            else:
                dist = buf('dist', (n,), torch.float32)
                valid = buf('valid_mask', (n,), torch.bool)
                inters = [torch.empty((V, n, kt[4]), dtype=torch.float32, device=dev) for kt in keys] if return_inter else None
                if binned is not False and binned is not None and n > 0:
                    if return_inter:
                        raise ValueError('binned=True does not support return_inter')
                    order = binned if isinstance(binned, torch.Tensor) else \
                        self.bin_order(pts, None if binned is True else float(binned))
                    _native.eval_ordered(V, H, W, pose_p, K_p, depth_p, pts.data_ptr(), n, order.data_ptr(), keys,
                                         dist.data_ptr(), valid.data_ptr(), [o.data_ptr() for o in outs],
                                         self._flags(eval_dist), float(self.mu), stream)
                else:
                    _native.eval_device(V, H, W, pose_p, K_p, depth_p, pts.data_ptr(), n, keys,
                                        dist.data_ptr(), valid.data_ptr(), [o.data_ptr() for o in outs],
                                        [t.data_ptr() for t in inters] if inters is not None else None,
                                        self._flags(eval_dist), float(self.mu), stream)
        res = {'dist': dist, 'valid_mask': valid}
        for i, k in enumerate([] if eval_dist else names):
            res[k] = outs[i]
            if return_inter:
                res[k + '_inter'] = inters[i]
        return res

    def _run_host(self, pts, names, keys, V, H, W, pose_p, K_p, depth_p, eval_dist, out=None):
        """CPU points in, CPU results out: one d3f_eval_host call (slab-pipelined H2D/kernel/D2H)."""
        n = int(pts.shape[0])
        pts = pts.contiguous()
        dev = self.curr_obs_torch['depth'].device
        pin = True

        def buf(name, shape, dtype):
            if out is not None and name in out:
                t = out[name]
                if tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous() or t.is_cuda:
                    raise ValueError(f"out['{name}'] must be a contiguous CPU {dtype} tensor of shape {tuple(shape)}")
                return t
            return torch.empty(shape, dtype=dtype, pin_memory=pin)

        dist = buf('dist', (n,), torch.float32)
        valid = buf('valid_mask', (n,), torch.bool)
        outs = [buf(k, (n, kt[4]), torch.float32) for k, kt in zip(names, keys)]
        with torch.cuda.device(dev):
            _native.eval_host(V, H, W, pose_p, K_p, depth_p, pts.data_ptr(), n, keys,
                              dist.data_ptr(), valid.data_ptr(), [o.data_ptr() for o in outs],
                              self._flags(eval_dist), float(self.mu), torch.cuda.current_stream(dev).cuda_stream)
        res = {'dist': dist, 'valid_mask': valid}
        for k, o in zip(names, outs):
            res[k] = o
        return res

    def eval(self, pts, return_names=['dino_feats', 'mask'], return_inter=False, out=None, binned=False):
        """(N,3) world points -> {'dist' (N,), 'valid_mask' (N,) bool, '<k>' (N,C_k) for k in return_names
        [, '<k>_inter' (V,N,C_k)]}.  Reference fusion.py:305-394.  CPU `pts` give CPU results through the
        host-buffer entry point (return_inter is device-only).  Extensions: `out` may hold preallocated
        tensors for 'dist' / 'valid_mask' / any name (the kernel writes them in place); `binned=True` (or a cell
        size in metres, or a precomputed bin_order() tensor) walks the points in lattice-cell order — same results
        bit for bit, about twice as fast for keypoints / mesh vertices that arrive in no spatial order."""
        if isinstance(pts, torch.Tensor) and not pts.is_cuda and return_inter:
            raise ValueError('return_inter needs device points')
        if isinstance(pts, torch.Tensor) and pts.requires_grad and torch.is_grad_enabled():
            if not pts.is_cuda or return_inter or out is not None:
                raise ValueError('differentiable eval needs device points, return_inter=False and no out=')
            self._check_pts(pts)
            names = list(return_names)
            outs = _FieldQuery.apply(pts, self, names)
            res = {'dist': outs[0], 'valid_mask': outs[1]}
            res.update({k: outs[2 + i] for i, k in enumerate(names)})
            return res
        return self._run(pts, return_names, return_inter, eval_dist=False, out=out, binned=binned)

    def eval_dist(self, pts):
        """Unclamped signed distance: {'dist', 'valid_mask'}.  Reference fusion.py:396-436."""
        return self._run(pts, [], False, eval_dist=True)

    def batch_eval(self, pts, return_names=['dino_feats', 'mask']):
        """Reference fusion.py:526-545 chunks by 60 000 points to bound its (V,n,C) temporaries; this
        path has none, so the whole batch is one launch and the result is the same dict."""
        return self.eval(pts, return_names=return_names)

    def eval_host(self, pts, return_names: Iterable[str] = (), out: Optional[dict] = None):
        """Host-buffer evaluation with optional preallocated (ideally pinned) outputs."""
        self._check_pts(pts)
        if pts.is_cuda:
            raise ValueError('eval_host takes CPU points')
        names = list(return_names)
        V, H, W, pose_p, K_p, depth_p = self._obs_ptrs()
        keys = [self._key_tuple(k, V)[0] for k in names]
        return self._run_host(pts, names, keys, V, H, W, pose_p, K_p, depth_p, False, out=out)

    # ------------------------------------------------------------------ dense sweeps -----
    def sweep_select(self, boundaries=None, res=0.001, pts=None, dist_threshold=0.005, mask_threshold=0.6,
                     mask_name='mask', capacity=None, dense=False):
        """The candidate search of select_features_rand / select_features_from_pcd (reference fusion.py:1420-1445,
        1477-1501) as ONE fused launch: for every voxel centre of create_init_grid(boundaries, res) — or every row of
        `pts` — dist / valid_mask, and where valid & |dist| < dist_threshold the normalised instance-mask field
        mask / (mask.sum(1) + 1e-7); points with an instance i >= 1 above mask_threshold are stream-compacted on the
        device.  Neither the grid (1.2 GB at the reference's 1 mm resolution) nor the dense mask field exists in HBM.

        Returns {'index' (K,) int64 ascending linear point index, 'inst' (K,) int64 instance id,
                 'pts' (K,3) the selected points, 'grid_shape' (sweeps of a grid), 'count' K
                 [, 'dist' (N,), 'valid_mask' (N,) when dense=True — what extract_mesh consumes, fusion.py:1321]}.
        The reference's masked_pts of instance i is  res['pts'][res['inst'] == i]  (same order)."""
        if len(self.curr_obs_torch) == 0:
            print('Please call update() first!')
            exit()
        V, H, W, pose_p, K_p, depth_p = self._obs_ptrs()
        dev = self.curr_obs_torch['depth'].device
        mk = None
        if mask_name is not None:
            kt, _vol = self._key_tuple(mask_name, V)
            mk = kt
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            grid = None
            if pts is None:
                # the three axis arrays exactly as the reference builds them (fusion.py:83-85); the kernel indexes them
                axes = [(torch.arange(boundaries[a + '_lower'], boundaries[a + '_upper'], res, dtype=torch.float32)
                         + res / 2).to(dev) for a in ('x', 'y', 'z')]
                shape = torch.Size(int(a.numel()) for a in axes)
                n = int(shape[0] * shape[1] * shape[2])
                grid = (axes[0].data_ptr(), axes[1].data_ptr(), axes[2].data_ptr(), shape[0], shape[1], shape[2])
                pts_p = None
            else:
                self._check_pts(pts)
                if not pts.is_cuda or pts.device != dev:
                    raise ValueError(f'pts must be on {dev}')
                pts = pts.contiguous()
                n, pts_p, shape, axes = int(pts.shape[0]), pts.data_ptr(), None, None
            dist = torch.empty(n, dtype=torch.float32, device=dev) if dense else None
            valid = torch.empty(n, dtype=torch.bool, device=dev) if dense else None
            cap = int(capacity) if capacity is not None else max(min(n, 1 << 16), n // 16)
            while True:
                count = torch.zeros(1, dtype=torch.int64, device=dev)
                sel_i = torch.empty(cap, dtype=torch.int32, device=dev)
                sel_c = torch.empty(cap, dtype=torch.int32, device=dev)
                _native.sweep_select(V, H, W, pose_p, K_p, depth_p, grid, pts_p, n, mk, float(dist_threshold),
                                     float(mask_threshold), None if dist is None else dist.data_ptr(),
                                     None if valid is None else valid.data_ptr(), cap if mk is not None else 0,
                                     count.data_ptr() if mk is not None else None,
                                     sel_i.data_ptr() if mk is not None else None,
                                     sel_c.data_ptr() if mk is not None else None, self._flags(False), float(self.mu), stream)
                k = int(count.item()) if mk is not None else 0
                if k <= cap:
                    break
                cap = k                                        # rare: the shell held more points than guessed
            idx, perm = torch.sort(sel_i[:k].long())
            inst = sel_c[:k].long()[perm]
            if pts is None:
                nz, ny = shape[2], shape[1]
                sel_pts = torch.stack([axes[0][idx // (ny * nz)], axes[1][(idx // nz) % ny], axes[2][idx % nz]], -1)
            else:
                sel_pts = pts[idx]
        res_d = {'index': idx, 'inst': inst, 'pts': sel_pts, 'count': k}
        if shape is not None:
            res_d['grid_shape'] = shape
        if dense:
            res_d['dist'], res_d['valid_mask'] = dist, valid
        return res_d

    # ------------------------------------------------------------------ callers of the path, kept as thin mirrors -----
    @staticmethod
    def farthest_point_sample(pts, n, init_idx=-1, generator=None):
        """Farthest-point sampling of `n` of the (M,3) device points (the role of utils/my_utils.py fps_np in the
        reference, fusion.py:1446): greedy max-min distance from a start point (init_idx, or random when -1).
        Runs on the device; returns (samples (n,3), indices (n,))."""
        M = int(pts.shape[0])
        assert M > 0
        n = int(n)
        start = int(init_idx) if init_idx != -1 else int(torch.randint(M, (1,), generator=generator).item())
        idx = torch.empty(n, dtype=torch.long, device=pts.device)
        idx[0] = start
        d = (pts - pts[start]).norm(dim=1)
        for k in range(1, n):
            j = torch.argmax(d)
            idx[k] = j
            d = torch.minimum(d, (pts - pts[j]).norm(dim=1))
        return pts[idx], idx

    def _select_features(self, cand, N, per_instance, init_idx, name='dino_feats'):
        """The per-instance loop of select_features_rand / select_features_from_pcd (fusion.py:1441-1451): for every
        instance i >= 1 (consecutive repeats of a label skipped unless per_instance) farthest-point-sample N of its
        candidates and evaluate their descriptors."""
        label = self.curr_obs_torch.get('consensus_mask_label')
        num_inst = int(self.curr_obs_torch['mask'].shape[-1])
        if label is None:
            label = [str(i) for i in range(num_inst)]
        feats, pts_out = [], []
        last_label = label[0]
        for i in range(1, len(label)):
            if label[i] == last_label and not per_instance:
                continue
            sel = cand['pts'][cand['inst'] == i]
            if sel.shape[0] == 0:                       # select_features_from_pcd skips empty instances (fusion.py:1504)
                last_label = label[i]
                continue
            sample, _ = self.farthest_point_sample(sel, N, init_idx)
            feats.append(self.eval(sample, return_names=[name])[name])
            pts_out.append(sample.cpu().numpy())
            last_label = label[i]
        return feats, pts_out, []                       # (src_feats_list, src_pts_list, img_list: drawing is out of scope)

    def select_features_rand(self, boundaries, N, per_instance=False, res=None, init_idx=-1):
        """Reference fusion.py:1418-1475: N descriptors per instance from a `res`-spaced sweep of `boundaries`
        (default 1 mm: 101 M voxels).  The sweep, the thresholds and the selection are ONE fused launch (sweep_select);
        FPS and the descriptor evaluation of the N samples follow on the device.  Returns (src_feats_list,
        src_pts_list, img_list) like the reference; img_list is empty (keypoint drawing is visualisation)."""
        cand = self.sweep_select(boundaries, 0.001 if res is None else res)
        return self._select_features(cand, N, per_instance, init_idx)

    def select_features_from_pcd(self, pcd, N, per_instance=False, init_idx=-1, vis=False):
        """Reference fusion.py:1477-1537: the same selection over an explicit (M,3) point cloud (numpy or tensor)."""
        dev = self.curr_obs_torch['depth'].device
        pts = _as_device(pcd, dev, torch.float32).contiguous()
        cand = self.sweep_select(pts=pts)
        return self._select_features(cand, N, per_instance, init_idx)

    def rigid_tracking(self, src_feat_info, last_match_pts_list, boundaries, rand_ptcl_num, iters=100):
        """Reference fusion.py:1608-1685: per-instance rigid pose by 100 Adam iterations through eval.  One CUDA-graph
        replay (tracking.FusedRigidTracker: 4 launches per iteration); the observation tensors are captured by address,
        so the graph is rebuilt when update() has replaced them.  Returns {'match_pts_list': [...]} like the reference."""
        from .tracking import FusedRigidTracker
        dev = self.curr_obs_torch['depth'].device
        src = torch.cat([src_feat_info[k]['src_feats'] for k in src_feat_info.keys()], dim=0).to(dev, torch.float32)
        num_inst = len(last_match_pts_list)
        last = torch.from_numpy(np.stack(last_match_pts_list, axis=0)).to(dev, torch.float32)
        assert tuple(last.shape[:2]) == (num_inst, rand_ptcl_num)
        key = (num_inst, int(rand_ptcl_num), int(src.shape[1]), int(iters),
               tuple(int(self.curr_obs_torch[k].data_ptr()) for k in ('pose', 'K', 'depth', 'dino_feats')))
        tr = getattr(self, '_tracker', None)
        if tr is None or getattr(self, '_tracker_key', None) != key:
            tr = self._tracker = FusedRigidTracker(self, num_inst, int(rand_ptcl_num), int(src.shape[1]), iters=iters)
            self._tracker_key = key
        res = tr.track(src, last)
        m = res['match_pts'].cpu().numpy()
        return {'match_pts_list': [m[i] for i in range(num_inst)]}

    def eval_pca(self, pts, name, mean, components):
        """eval(pts, [name])[name] followed by sklearn's PCA.transform, (row - mean) @ components.T (what the
        reference does on the host after eval, fusion.py:1386-1392) — without ever forming the (N,C) field.

        The field is linear in the sampled map, so the map is projected first (V*h*w rows instead of N) and the
        k-channel result is queried through the narrow path with bias = mean @ components.T:
            (field - mean) @ W^T  ==  field_of(map @ W^T) - mean @ W^T.
        Returns {'dist', 'valid_mask', name + '_pca' (N,k)}."""
        self._check_pts(pts)
        V, H, W, pose_p, K_p, depth_p = self._obs_ptrs()
        vol = self.curr_obs_torch[name]
        if not (isinstance(vol, torch.Tensor) and vol.dim() == 4 and vol.is_cuda and vol.is_contiguous()
                and vol.dtype == torch.float32 and vol.shape[0] == V):
            raise ValueError(f"curr_obs_torch['{name}'] must be a contiguous float32 CUDA (V,h,w,C) tensor")
        dev = vol.device
        if not pts.is_cuda or pts.device != dev:
            raise ValueError(f'pts must be on {dev}')
        _, h, w, C = (int(x) for x in vol.shape)
        mean = _as_device(mean, dev, torch.float32).contiguous().reshape(1, C)
        comp = _as_device(components, dev, torch.float32).contiguous()
        k = int(comp.shape[0])
        if tuple(comp.shape) != (k, C) or not (1 <= k <= 8):
            raise ValueError(f'components must be (k,{C}) with 1 <= k <= 8')
        n = int(pts.shape[0])
        pts = pts.contiguous()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            proj = torch.empty((V, h, w, k), dtype=torch.float32, device=dev)
            bias = torch.empty((1, k), dtype=torch.float32, device=dev)
            _native.pca_project(vol.data_ptr(), V * h * w, C, None, comp.data_ptr(), k, proj.data_ptr(), stream)
            _native.pca_project(mean.data_ptr(), 1, C, None, comp.data_ptr(), k, bias.data_ptr(), stream)
            dist = torch.empty(n, dtype=torch.float32, device=dev)
            valid = torch.empty(n, dtype=torch.bool, device=dev)
            y = torch.empty((n, k), dtype=torch.float32, device=dev)
            _native.eval_device(V, H, W, pose_p, K_p, depth_p, pts.data_ptr(), n,
                                [(proj.data_ptr(), _native.D3F_F32, h, w, k, bias.data_ptr())],
                                dist.data_ptr(), valid.data_ptr(), [y.data_ptr()], None,
                                self._flags(False), float(self.mu), stream)
        return {'dist': dist, 'valid_mask': valid, name + '_pca': y}

    # ------------------------------------------------------------------ descriptors -> PCA
    def pca_project(self, feats, mean, components):
        """(feats - mean) @ components.T on the device: sklearn PCA.transform as the reference applies it
        to descriptors on the host (fusion.py:1386-1392).  feats (N,C), mean (C,), components (k,C)."""
        dev = feats.device
        mean = _as_device(mean, dev, torch.float32).contiguous() if mean is not None else None
        components = _as_device(components, dev, torch.float32).contiguous()
        feats = feats.contiguous()
        n, c = feats.shape
        k = components.shape[0]
        y = torch.empty((n, k), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.pca_project(feats.data_ptr(), n, c, mean.data_ptr() if mean is not None else None,
                                components.data_ptr(), k, y.data_ptr(),
                                torch.cuda.current_stream(dev).cuda_stream)
        return y


def project_points_coords(pts, Rt, K):
    """Reference fusion.py:32-55, kept for callers that import it: (pts_2d (V,N,2), valid (V,N), depth (V,N,1)).
    Not on the hot path (eval projects inside its kernel); computed with the same float32 sequence in torch."""
    V, N = Rt.shape[0], pts.shape[0]
    KRt = torch.zeros((V, 3, 4), dtype=pts.dtype, device=pts.device)
    for k in range(3):
        KRt = KRt + K[:, :, k:k + 1] * Rt[:, k:k + 1, :]
    hp = torch.cat([pts, torch.ones((N, 1), dtype=pts.dtype, device=pts.device)], 1)
    cam = torch.zeros((V, N, 3), dtype=pts.dtype, device=pts.device)
    for k in range(4):
        cam = cam + KRt[:, None, :, k] * hp[None, :, k:k + 1]
    depth = cam[:, :, 2:].clone()
    bad = depth.abs() < 1e-4
    depth[bad] = 1e-3
    return cam[:, :, :2] / depth, ~bad[..., 0], depth


def interpolate_feats(feats, points, h=None, w=None, padding_mode='zeros', align_corners=False, inter_mode='bilinear'):
    """Reference fusion.py:57-77, kept for callers that import it: feats (b,f,h,w), points (b,n,2) in pixels
    of an (h,w) image -> (b,n,f).  Not on the hot path (eval samples inside its kernel)."""
    import torch.nn.functional as F
    b, _, ch, cw = feats.shape
    if h is None and w is None:
        h, w = ch, cw
    gx = points[:, :, 0] / (w - 1) * 2 - 1
    gy = points[:, :, 1] / (h - 1) * 2 - 1
    grid = torch.stack([gx, gy], -1).unsqueeze(1)
    out = F.grid_sample(feats, grid, mode=inter_mode, padding_mode=padding_mode, align_corners=align_corners)
    return out.squeeze(2).permute(0, 2, 1)
