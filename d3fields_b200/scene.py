"""Seeded synthetic multi-view scenes for parity tests and the benchmark.

The reference ships no test data for the field query, so every
parity case is generated here from a seed: ring cameras looking at the origin,
an analytic depth buffer (sphere over a ground plane) with noise and holes,
normal-distributed feature volumes, block-random instance masks, and query
points laid out the way the reference's callers lay them out (a voxel-centre
grid with z fastest, reference fusion.py:79-88, or scattered keypoints).

numpy only: the module must import on a box that has neither the reference nor
a GPU.  All randomness comes from ``np.random.RandomState`` (a frozen stream),
so a given (config, seed) names the same bytes on every machine; the golden
fixtures under tests/golden/ additionally pin a sha256 of every input array.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

# Workspace of the reference's demo (reference vis_repr.py:39-44).
WORKSPACE = dict(x_lower=-0.4, x_upper=0.4, y_lower=-0.4, y_upper=0.3,
                 z_lower=-0.2, z_upper=0.02)


@dataclass
class Scene:
    """One observation in the layout of ``Fusion.curr_obs_torch`` (numpy side).

    pose   (V,3,4) f32  world->camera [R|t]      (reference fusion.py:711)
    K      (V,3,3) f32  intrinsics               (reference fusion.py:712)
    depth  (V,H,W) f32  metres, 0 = hole         (reference fusion.py:710)
    maps   name -> (V,h,w,C) f32 or u8, channels-last (reference fusion.py:618, :1171)
    """
    H: int
    W: int
    pose: np.ndarray
    K: np.ndarray
    depth: np.ndarray
    maps: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def V(self) -> int:
        return int(self.pose.shape[0])


def ring_cameras(V: int, H: int, W: int, radius: float = 0.8, height: float = 0.6
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """V cameras on a ring, looking at the origin, OpenCV axes (x right, y down, z forward)."""
    pose = np.zeros((V, 3, 4), dtype=np.float64)
    Kmat = np.zeros((V, 3, 3), dtype=np.float64)
    for v in range(V):
        th = 2.0 * np.pi * v / V
        eye = np.array([radius * np.cos(th), radius * np.sin(th), height])
        fwd = -eye / np.linalg.norm(eye)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], 0)          # rows: camera axes in world coords
        pose[v, :, :3] = R
        pose[v, :, 3] = -R @ eye
        Kmat[v] = [[0.8 * W, 0.0, W / 2.0], [0.0, 0.8 * W, H / 2.0], [0.0, 0.0, 1.0]]
    return pose.astype(np.float32), Kmat.astype(np.float32)


def raycast_depth(pose: np.ndarray, Kmat: np.ndarray, H: int, W: int,
                  sphere_r: float = 0.25, plane_z: float = -0.1) -> np.ndarray:
    """z-depth of the first hit of each pixel ray with a sphere at the origin or the plane z=plane_z."""
    V = pose.shape[0]
    out = np.zeros((V, H, W), dtype=np.float64)
    u, v_ = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    for i in range(V):
        R = pose[i, :, :3].astype(np.float64)
        t = pose[i, :, 3].astype(np.float64)
        Kd = Kmat[i].astype(np.float64)
        eye = -R.T @ t
        dc = np.stack([(u - Kd[0, 2]) / Kd[0, 0], (v_ - Kd[1, 2]) / Kd[1, 1], np.ones_like(u)], -1)
        dw = dc @ R                                   # camera dir -> world dir (R^T applied on the right)
        # plane
        with np.errstate(divide='ignore', invalid='ignore'):
            tp = (plane_z - eye[2]) / dw[..., 2]
        tp = np.where(tp > 0, tp, np.inf)
        # sphere: |eye + s*dw|^2 = r^2
        a = (dw * dw).sum(-1)
        b = 2.0 * (dw @ eye)
        c = eye @ eye - sphere_r ** 2
        disc = b * b - 4 * a * c
        ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
        ts = np.where(ts > 0, ts, np.inf)
        d = np.minimum(tp, ts)
        out[i] = np.where(np.isfinite(d), d, 0.0)
    return out


def grid_points(nx: int, ny: int, nz: int, bounds: Optional[dict] = None) -> np.ndarray:
    """Voxel-centre grid, z fastest then y then x — the order reference create_init_grid
    (fusion.py:79-88) produces — with a per-axis step so any (nx,ny,nz) fits the workspace."""
    b = bounds or WORKSPACE
    def centres(lo, hi, n):
        step = (hi - lo) / n
        return (lo + step * (np.arange(n, dtype=np.float64) + 0.5)).astype(np.float32)
    x = centres(b['x_lower'], b['x_upper'], nx)
    y = centres(b['y_lower'], b['y_upper'], ny)
    z = centres(b['z_lower'], b['z_upper'], nz)
    xx, yy, zz = np.meshgrid(x, y, z, indexing='ij')
    return np.ascontiguousarray(np.stack([xx, yy, zz], -1).reshape(-1, 3))


def scattered_points(n: int, seed: int, sigma: float = 0.25) -> np.ndarray:
    """Keypoint-like queries with no spatial locality."""
    rs = np.random.RandomState(seed + 7919)
    return (rs.standard_normal((n, 3)) * sigma).astype(np.float32)


def adversarial_points(scene: Scene, seed: int, n_each: int = 64) -> np.ndarray:
    """Points that hit the reference's corner cases (SURVEY.md §8a quirks 1,2,4):
    behind each camera, on each camera's z=0 plane (|z|<1e-4 -> replaced by 1e-3),
    far outside every image, and exactly at camera centres."""
    rs = np.random.RandomState(seed + 104729)
    out = []
    for v in range(scene.V):
        R = scene.pose[v, :, :3].astype(np.float64)
        t = scene.pose[v, :, 3].astype(np.float64)
        eye = -R.T @ t
        right, down, fwd = R[0], R[1], R[2]
        ab = rs.uniform(-0.3, 0.3, size=(n_each, 2))
        out.append(eye + ab[:, :1] * right + ab[:, 1:] * down)                    # on the z=0 plane
        out.append(eye + ab[:, :1] * right + ab[:, 1:] * down
                   + rs.uniform(-5e-5, 5e-5, size=(n_each, 1)) * fwd)             # |z| < 1e-4
        out.append(eye - rs.uniform(0.05, 1.0, size=(n_each, 1)) * fwd
                   + 0.2 * ab[:, :1] * right + 0.2 * ab[:, 1:] * down)            # behind the camera
        out.append(eye[None, :])                                                   # the centre itself
    out.append(rs.uniform(-30, 30, size=(n_each, 3)))                              # far away
    return np.concatenate(out, 0).astype(np.float32)


def make_scene(V: int, H: int, W: int, seed: int = 0,
               feat: Optional[Tuple[int, int, int]] = None,
               num_inst: int = 0, mask_dtype: str = 'f32',
               color: bool = False, hole_frac: float = 0.05,
               feat_name: str = 'dino_feats') -> Scene:
    """Build one synthetic observation.

    feat      (h, w, C) of the descriptor volume, e.g. (H//10, W//10, 1024) — the
              reference samples DINOv2 patch tokens at (H//10, W//10) (fusion.py:695-696)
    num_inst  >0 adds 'mask' (V,H,W,num_inst) one-hot of a 16x16-block random label image
              (layout of reference fusion.py:1171), stored as f32 or u8
    color     adds 'color_tensor' (V,H,W,3) in [0,1] (reference fusion.py:709)
    """
    rs = np.random.RandomState(seed)
    pose, Kmat = ring_cameras(V, H, W)
    depth = raycast_depth(pose, Kmat, H, W)
    depth = depth + rs.uniform(-0.01, 0.01, size=depth.shape)
    depth = np.where(rs.uniform(size=depth.shape) < hole_frac, 0.0, depth)
    depth = np.maximum(depth, 0.0).astype(np.float32)
    maps: Dict[str, np.ndarray] = {}
    if feat is not None:
        h, w, C = feat
        maps[feat_name] = rs.standard_normal((V, h, w, C)).astype(np.float32)
    if num_inst > 0:
        bh, bw = (H + 15) // 16, (W + 15) // 16
        lab = rs.randint(0, num_inst, size=(V, bh, bw))
        lab = np.repeat(np.repeat(lab, 16, 1), 16, 2)[:, :H, :W]
        onehot = (lab[..., None] == np.arange(num_inst)[None, None, None, :])
        maps['mask'] = onehot.astype(np.uint8 if mask_dtype == 'u8' else np.float32)
    if color:
        maps['color_tensor'] = (rs.randint(0, 256, size=(V, H, W, 3)).astype(np.float32)
                                / np.float32(255.0)).astype(np.float32)
    return Scene(H=H, W=W, pose=pose, K=Kmat, depth=depth, maps=maps)


def sha256_of(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------------------
# Named configurations of BASELINE.json.  'pts' is built lazily by
# config_points() so importing this module stays cheap.
# ---------------------------------------------------------------------------------------
CONFIGS = {
    # cfg1: the reference's CPU-runnable case: 10k grid points, 2 views 240x320, C=64 @ (24,32)
    'cfg1': dict(V=2, H=240, W=320, feat=(24, 32, 64), num_inst=8, grid=(28, 28, 14), n=10000),
    # cfg2a: 1M grid points, 4 views 480x640, C=1024 @ (48,64) (reference-faithful map size)
    'cfg2a': dict(V=4, H=480, W=640, feat=(48, 64, 1024), num_inst=0, grid=(100, 100, 100), n=1000000),
    # cfg3: 1M points, 4 views, SAM mask field num_inst=8 @ (480,640)
    'cfg3': dict(V=4, H=480, W=640, feat=None, num_inst=8, grid=(100, 100, 100), n=1000000),
}


def config_points(name: str) -> np.ndarray:
    c = CONFIGS[name]
    pts = grid_points(*c['grid'])
    return np.ascontiguousarray(pts[:c['n']])
