"""TEST INFRASTRUCTURE — CPU restatement of the reference's field query.  Not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; d3fields_b200/ never does (the product path fails loudly when
the CUDA library is missing instead of falling back to anything in here).

What is restated (every step cites the reference line it follows, paths relative to
/root/reference):

  project()      fusion.py:32-55   project_points_coords
  sample_*()     fusion.py:57-77   interpolate_feats -> torch.nn.functional.grid_sample
                                   (align_corners=True, padding 'zeros'; nearest / bilinear)
  field_eval()   fusion.py:305-394 Fusion.eval          (eval_dist=False)
                 fusion.py:396-436 Fusion.eval_dist     (eval_dist=True)
  init_grid()    fusion.py:79-88   create_init_grid (axes computed by torch.arange, as the reference does)
  select_candidates()  fusion.py:1420-1445 / 1477-1501  the candidate search of select_features_rand /
                                   select_features_from_pcd up to (not including) farthest-point sampling

Third-party arithmetic not under /root/reference: torch (reference pins pytorch=2.1.0 in
env.yaml:11; this image has 2.11.0).  The two torch ops whose rounding decides integer
results are restated from their published CPU algorithms:

  * ``Tensor @ Tensor`` with 4x4 . 4x1 operands: ATen's small-matrix bmm kernel,
    ``r = 0; for k: r += a[k]*b[k]`` in fp32, products and sums rounded separately.
  * ``grid_sample`` (CPU, vectorised kernel, align_corners=True):
    ``ix = (x_norm + 1) * ((size-1)/2)``; nearest = round-half-to-even, in-bounds test on
    the rounded index; bilinear = floor, ``w = ix - floor``, ``e = 1 - w`` (same in y),
    corner weights ``nw=s*e, ne=s*w, sw=n*e, se=n*w``, out-of-range corners contribute 0,
    accumulation order nw, ne, sw, se.

Parity pinning: the reference has no tests or golden vectors for this path,
so this restatement is pinned against outputs of the unmodified reference itself, run in
the build container by oracle/gen_golden.py and committed under tests/golden/
(tests/test_oracle_golden.py).  Integer/boolean results (valid_mask, the sampled pixel
indices) and ``dist`` must match bit-for-bit; descriptors within 1e-4 relative.

All arithmetic is numpy float32, one rounding per operation, no fused multiply-add; the
only libm call is exp() in the distance weight, which never feeds an integer decision.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np

F32 = np.float32


def _f(x) -> np.float32:
    return np.float32(x)


def project(pts: np.ndarray, pose: np.ndarray, K: np.ndarray):
    """reference fusion.py:32-55.

    pts (N,3) f32, pose (V,3,4) f32 world->camera, K (V,3,3) f32.
    Returns pts_2d (V,N,2) f32 pixel coords, ok (V,N) bool, z (V,N) f32 (after the 1e-3 patch).
    """
    pts = np.ascontiguousarray(pts, dtype=F32)
    V = pose.shape[0]
    N = pts.shape[0]
    # KRt = K @ Rt (fusion.py:45): small-matrix kernel, sequential k, separate mul and add.
    KRt = np.zeros((V, 3, 4), dtype=F32)
    for k in range(3):
        KRt = (KRt + K[:, :, k:k + 1].astype(F32) * pose[:, k:k + 1, :].astype(F32)).astype(F32)
    # H = [KRt; 0 0 0 1] (fusion.py:46-48); pts_cam = H @ [x y z 1]^T (fusion.py:43,49-50)
    hp = np.concatenate([pts, np.ones((N, 1), dtype=F32)], 1)          # (N,4)
    cam = np.zeros((V, N, 3), dtype=F32)
    for k in range(4):
        cam = (cam + KRt[:, None, :, k] * hp[None, :, k:k + 1]).astype(F32)
    z = cam[:, :, 2].copy()                                            # fusion.py:51
    bad = np.abs(z) < _f(1e-4)                                         # fusion.py:52
    z[bad] = _f(1e-3)                                                  # fusion.py:53 (in place)
    pts_2d = (cam[:, :, :2] / z[:, :, None]).astype(F32)               # fusion.py:54
    return pts_2d, ~bad, z


def _unnormalised(p: np.ndarray, img_size: int, map_size: int) -> np.ndarray:
    """fusion.py:72-73 followed by grid_sample's align_corners=True un-normalisation.
    Normalised by the IMAGE size, sampled on the MAP size."""
    n = ((p / _f(img_size - 1)).astype(F32) * _f(2)).astype(F32) - _f(1)
    n = n.astype(F32)
    sf = _f(_f(map_size - 1) / _f(2))
    return ((n + _f(1)).astype(F32) * sf).astype(F32)


def sample_nearest(img: np.ndarray, pts_2d: np.ndarray, H: int, W: int) -> np.ndarray:
    """interpolate_feats(..., inter_mode='nearest') on a (V,h,w) map -> (V,N). fusion.py:327-333."""
    V, h, w = img.shape
    ix = _unnormalised(pts_2d[..., 0], W, w)
    iy = _unnormalised(pts_2d[..., 1], H, h)
    xr = np.rint(ix)            # round-half-to-even
    yr = np.rint(iy)
    inb = (xr >= 0) & (xr <= w - 1) & (yr >= 0) & (yr <= h - 1)       # NaN/inf compare false
    xi = np.where(inb, xr, 0).astype(np.int64)
    yi = np.where(inb, yr, 0).astype(np.int64)
    vi = np.arange(V)[:, None]
    return np.where(inb, img[vi, yi, xi], _f(0)).astype(F32)


def bilinear_setup(pts_2d: np.ndarray, H: int, W: int, h: int, w: int):
    """Corner indices, in-bounds flags and weights of grid_sample bilinear for a (h,w) map."""
    ix = _unnormalised(pts_2d[..., 0], W, w)
    iy = _unnormalised(pts_2d[..., 1], H, h)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx = (ix - x0).astype(F32)
    ex = (_f(1) - wx).astype(F32)
    wy = (iy - y0).astype(F32)
    sy = (_f(1) - wy).astype(F32)
    wts = [(sy * ex).astype(F32), (sy * wx).astype(F32), (wy * ex).astype(F32), (wy * wx).astype(F32)]
    x1 = x0 + 1
    y1 = y0 + 1
    def ok(xx, yy):
        return (xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)
    corners = [(x0, y0), (x1, y0), (x0, y1), (x1, y1)]                 # nw, ne, sw, se
    inb = [ok(xx, yy) for xx, yy in corners]
    idx = [(np.where(m, yy, 0).astype(np.int64), np.where(m, xx, 0).astype(np.int64))
           for (xx, yy), m in zip(corners, inb)]
    return idx, inb, wts


def sample_bilinear(vol: np.ndarray, pts_2d: np.ndarray, H: int, W: int) -> np.ndarray:
    """interpolate_feats(..., inter_mode='bilinear') on a (V,h,w,C) map -> (V,N,C). fusion.py:373-379."""
    V, h, w, C = vol.shape
    idx, inb, wts = bilinear_setup(pts_2d, H, W, h, w)
    vi = np.arange(V)[:, None]
    out = np.zeros(pts_2d.shape[:2] + (C,), dtype=F32)
    for (yi, xi), m, wt in zip(idx, inb, wts):
        val = vol[vi, yi, xi].astype(F32)                              # (V,N,C)
        val = np.where(m[..., None], val, _f(0))
        out = (out + (val * wt[..., None]).astype(F32)).astype(F32)
    return out


def visibility(pts: np.ndarray, pose, K, depth, H: int, W: int, mu: float, eval_dist: bool = False):
    """The per-view part of Fusion.eval that decides integer results. fusion.py:323-347 (eval),
    :412-427 (eval_dist).  Returns pts_2d, d (V,N) raw signed distance, vis (V,N) bool, weight (V,N)."""
    pts_2d, ok, z = project(pts, pose, K)
    inter_depth = sample_nearest(depth.astype(F32), pts_2d, H, W)      # fusion.py:327-333
    d = (inter_depth - z).astype(F32)                                  # fusion.py:343
    muf = _f(mu)
    if eval_dist:
        vis = (inter_depth > 0) & ok                                   # fusion.py:423
        weight = np.ones_like(d)
    else:
        vis = (inter_depth > 0) & ok & (d > -muf)                      # fusion.py:344
        a = np.minimum((muf - np.abs(d)).astype(F32), _f(0))           # clamp(max=0)
        weight = np.exp((a / muf).astype(F32)).astype(F32)             # fusion.py:347
    return pts_2d, d, vis, weight


def field_eval(pts: np.ndarray, pose: np.ndarray, K: np.ndarray, depth: np.ndarray,
               H: int, W: int, maps: Optional[Dict[str, np.ndarray]] = None,
               return_names: Iterable[str] = (), mu: float = 0.02,
               eval_dist: bool = False, return_inter: bool = False,
               chunk: int = 8192) -> Dict[str, np.ndarray]:
    """Fusion.eval (fusion.py:305-394) / Fusion.eval_dist (fusion.py:396-436) on numpy arrays.

    maps[name] is (V,h,w,C) float32 or uint8 (a uint8 map is read as its float value; the
    reference always stores float, reference fusion.py:1171).
    Returns 'dist' (N,) f32, 'valid_mask' (N,) bool and one (N,C) f32 array per name
    (plus name+'_inter' (V,N,C) when return_inter).
    """
    maps = maps or {}
    names = list(return_names)
    pts = np.ascontiguousarray(pts, dtype=F32)
    N = pts.shape[0]
    out: Dict[str, list] = {'dist': [], 'valid_mask': []}
    for k in names:
        out[k] = []
        if return_inter:
            out[k + '_inter'] = []
    muf = _f(mu)
    for s in range(0, max(N, 1), chunk):
        p = pts[s:s + chunk]
        pts_2d, d, vis, weight = visibility(p, pose, K, depth, H, W, mu, eval_dist)
        visf = vis.astype(F32)
        if not eval_dist:
            d = np.minimum(np.maximum(d, -muf), muf)                   # fusion.py:358
        cnt = np.zeros(p.shape[0], dtype=F32)
        acc = np.zeros(p.shape[0], dtype=F32)
        for v in range(pose.shape[0]):                                 # sum(0): views in order
            acc = (acc + (d[v] * visf[v]).astype(F32)).astype(F32)
            cnt = (cnt + visf[v]).astype(F32)
        dist = (acc / (cnt + _f(1e-6)).astype(F32)).astype(F32)        # fusion.py:364 / :427
        none = cnt == 0                                                # fusion.py:366 / :429
        if not eval_dist:
            dist[none] = _f(1e3)                                       # fusion.py:367
        out['dist'].append(dist)
        out['valid_mask'].append(~none)
        for k in names:
            inter = sample_bilinear(maps[k], pts_2d, H, W)             # fusion.py:373-379
            wv = (visf * weight).astype(F32)
            val = np.zeros(inter.shape[1:], dtype=F32)
            for v in range(pose.shape[0]):                             # fusion.py:385
                term = ((inter[v] * visf[v][:, None]).astype(F32) * weight[v][:, None]).astype(F32)
                val = (val + term).astype(F32)
            val = (val / (cnt + _f(1e-6)).astype(F32)[:, None]).astype(F32)
            val[none] = _f(0)                                          # fusion.py:386
            out[k].append(val)
            if return_inter:
                out[k + '_inter'].append(inter)
    res: Dict[str, np.ndarray] = {}
    for k, parts in out.items():
        if k.endswith('_inter'):
            res[k] = np.concatenate(parts, 1) if parts else np.zeros((pose.shape[0], 0, 0), F32)
        else:
            res[k] = np.concatenate(parts, 0)
    return res


def init_grid(boundaries: dict, step: float):
    """reference fusion.py:79-88 create_init_grid: (coords (N,3) f32, shape, (x,y,z) axis arrays).  The axis values
    are torch.arange's (float32, vectorised CPU kernel) plus step/2 — computed with torch itself, which is the
    third-party arithmetic the reference uses here."""
    import torch
    axes = [(torch.arange(boundaries[a + '_lower'], boundaries[a + '_upper'], step, dtype=torch.float32) + step / 2).numpy()
            for a in ('x', 'y', 'z')]
    xx, yy, zz = np.meshgrid(*axes, indexing='ij')
    return np.ascontiguousarray(np.stack([xx, yy, zz], -1).reshape(-1, 3)), xx.shape, axes


def select_candidates(pts: np.ndarray, pose, K, depth, H: int, W: int, mask: np.ndarray, mu: float = 0.02,
                      dist_threshold: float = 0.005, mask_threshold: float = 0.6, field_fn=None):
    """reference fusion.py:1428-1445 (select_features_rand) == :1484-1501 (select_features_from_pcd):

        out = batch_eval(pts, ['mask']);  dist_mask = |out.dist| < 0.005
        m = out.mask / (out.mask.sum(dim=1, keepdim=True) + 1e-7)
        for i in 1..num_inst-1:  selected_i = (m[:, i] > 0.6) & dist_mask & out.valid_mask

    Returns (index (K,) int64 ascending, inst (K,) int64, margin (N,) f32 = min_i |m_i - threshold| over i >= 1 for
    points inside the shell, +inf elsewhere — the distance of each point's decision from the threshold).
    torch's sum(dim=1) over the instance axis is sequential for num_inst <= 4 and == 8 (measured against torch 2.11
    CPU); other widths use a vector-width dependent cascade, i.e. the reference's own sum is defined to 1 ulp there.
    field_fn: alternative implementation of field_eval (e.g. the C oracle) with the same signature."""
    fe = field_fn or field_eval
    out = fe(pts, pose, K, depth, H, W, {'mask': mask}, ['mask'], mu=mu)
    m = out['mask'].astype(F32)
    shell = (np.abs(out['dist']) < _f(dist_threshold)) & out['valid_mask']
    s = np.zeros(m.shape[0], dtype=F32)
    for j in range(m.shape[1]):
        s = (s + m[:, j]).astype(F32)
    mn = (m / (s + _f(1e-7)).astype(F32)[:, None]).astype(F32)
    inst = np.zeros(m.shape[0], dtype=np.int64)
    for i in range(m.shape[1] - 1, 0, -1):
        inst[(mn[:, i] > _f(mask_threshold)) & shell] = i
    margin = np.full(m.shape[0], np.inf, dtype=F32)
    if m.shape[1] > 1:
        margin[shell] = np.abs(mn[shell, 1:] - _f(mask_threshold)).min(1)
    idx = np.nonzero(inst > 0)[0]
    return idx, inst[idx], margin
