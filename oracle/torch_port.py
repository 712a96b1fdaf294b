"""TEST INFRASTRUCTURE — the reference's CPU algorithm restated with the same torch operators.

The reference's implementation of this path IS a sequence of torch operators (reference
fusion.py:32-77, 305-394, 526-545); it is Python and cannot travel to the GPU box, so this
port stands in for it wherever bench.py reports a CPU baseline ("kind": "port").  It issues
the same operator sequence on the same shapes — the homogeneous broadcast matmul, the two
grid_sample calls per key, the (V,n,C) broadcast multiplies and the view sum — so its run
time on the host cores is the reference's run time; tests/test_oracle_golden.py checks its
outputs against the golden vectors of the unmodified reference (bit-exact dist/valid_mask).

Only bench.py (cpu_baseline / --impl reference) and tests/ may import this module.
"""
from __future__ import annotations

from typing import Dict, Iterable

import torch
import torch.nn.functional as F

CHUNK = 60000      # reference fusion.py:527


def _project(pts, Rt, K):
    """reference fusion.py:32-55"""
    n = pts.shape[0]
    V = Rt.shape[0]
    homog = torch.cat([pts, pts.new_ones((n, 1))], dim=1)
    bottom = pts.new_zeros((V, 1, 4))
    bottom[:, :, 3] = 1.0
    Hm = torch.cat([K @ Rt, bottom], dim=1)
    cam = (Hm[:, None] @ homog[None, :, :, None])[:, :, :3, 0]
    z = cam[:, :, 2:]
    near = z.abs() < 1e-4
    z[near] = 1e-3
    return cam[:, :, :2] / z, ~near[..., 0], z


def _sample(maps_nchw, pix, H, W, mode):
    """reference fusion.py:57-77 with align_corners=True, zero padding"""
    gx = pix[:, :, 0] / (W - 1) * 2 - 1
    gy = pix[:, :, 1] / (H - 1) * 2 - 1
    grid = torch.stack([gx, gy], dim=-1).unsqueeze(1)
    s = F.grid_sample(maps_nchw, grid, mode=mode, padding_mode='zeros', align_corners=True)
    return s.squeeze(2).permute(0, 2, 1)


def eval_chunk(obs: Dict[str, torch.Tensor], H: int, W: int, pts: torch.Tensor,
               return_names: Iterable[str] = (), mu: float = 0.02) -> Dict[str, torch.Tensor]:
    """reference fusion.py:305-394 on one chunk.  Differentiable with respect to pts exactly like the reference
    (rigid_tracking back-propagates through it, fusion.py:1650-1665); wrap in torch.no_grad() for timing."""
    pix, in_front, z = _project(pts, obs['pose'], obs['K'])
    z = z[..., 0]
    seen_depth = _sample(obs['depth'].unsqueeze(1), pix, H, W, 'nearest')[..., 0]
    d = seen_depth - z
    vis = (seen_depth > 0.0) & in_front & (d > -mu)
    wgt = torch.exp(torch.clamp(mu - torch.abs(d), max=0) / mu)
    d = torch.clamp(d, min=-mu, max=mu)
    visf = vis.float()
    cnt = visf.sum(0)
    dist = (d * visf).sum(0) / (cnt + 1e-6)
    none = cnt == 0
    dist[none] = 1e3
    out = {'dist': dist, 'valid_mask': ~none}
    for k in return_names:
        s = _sample(obs[k].permute(0, 3, 1, 2), pix, H, W, 'bilinear')
        val = (s * visf.unsqueeze(-1) * wgt.unsqueeze(-1)).sum(0) / (cnt.unsqueeze(-1) + 1e-6)
        val[none] = 0.0
        out[k] = val
    return out


@torch.no_grad()
def batch_eval(obs, H, W, pts, return_names=(), mu: float = 0.02):
    """reference fusion.py:526-545: 60 000-point chunks, concatenated"""
    parts: Dict[str, list] = {}
    for s in range(0, pts.shape[0], CHUNK):
        o = eval_chunk(obs, H, W, pts[s:s + CHUNK], return_names, mu)
        for k, v in o.items():
            parts.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in parts.items()}


def obs_from_scene(scene) -> Dict[str, torch.Tensor]:
    obs = {'pose': torch.from_numpy(scene.pose), 'K': torch.from_numpy(scene.K),
           'depth': torch.from_numpy(scene.depth)}
    for k, v in scene.maps.items():
        obs[k] = torch.from_numpy(v).float()
    return obs
