"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python -m oracle.gen_golden            # writes tests/golden/<case>.npz

Each fixture pins one seeded case (d3fields_b200.scene builds the inputs from the stored
parameters; a sha256 of every input array is stored so a drifting generator is caught)
and holds what reference fusion.py:305 Fusion.eval / :396 eval_dist / :526 batch_eval
returned on CPU for it:

  dist, valid_mask           in full when small, plus a sha256 of the reference's bytes
  <key>, <key>_inter         a strided subset of rows (ROWS) plus float64 sum / abs-sum
                             checksums over the full array

tests/test_oracle_golden.py checks oracle/field_oracle.py and oracle/d3f_oracle.c against
these; tests/test_parity_gpu.py checks the CUDA path against them on the GPU box.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from d3fields_b200 import scene as S            # noqa: E402
from oracle import ref_loader as RL             # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
ROWS = 192          # rows of each (N,C) output kept verbatim


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tie_scene() -> S.Scene:
    """Identity camera, W-1 and H-1 powers of two: the normalise/un-normalise round trip is
    exact, so half-integer pixel coordinates reach grid_sample as exact ties."""
    H, W = 129, 257
    rs = np.random.RandomState(11)
    pose = np.zeros((1, 3, 4), np.float32)
    pose[0, :, :3] = np.eye(3)
    K = np.eye(3, dtype=np.float32)[None].copy()
    depth = rs.uniform(0.9, 1.1, size=(1, H, W)).astype(np.float32)
    depth[0, ::7, ::5] = 0.0
    maps = {'dino_feats': rs.standard_normal((1, 17, 33, 8)).astype(np.float32),     # (h-1),(w-1) powers of two
            'mask': (rs.randint(0, 3, size=(1, H, W, 1)) == np.arange(3)).astype(np.float32)}
    return S.Scene(H=H, W=W, pose=pose, K=K, depth=depth, maps=maps)


def tie_points() -> np.ndarray:
    xs = np.array([-1.5, -0.5, 0.0, 0.5, 1.5, 2.5, 3.5, 10.5, 11.5, 127.5, 128.0, 254.5, 255.5, 256.0, 256.5, 257.5],
                  np.float32)
    ys = np.array([-0.5, 0.5, 1.5, 2.5, 63.5, 64.5, 127.5, 128.0, 128.5, 129.5], np.float32)
    xx, yy = np.meshgrid(xs, ys, indexing='ij')
    base = np.stack([xx.ravel(), yy.ravel(), np.ones(xx.size, np.float32)], -1)
    up = np.nextafter(base[:, :2], np.float32(1e9)).astype(np.float32)
    dn = np.nextafter(base[:, :2], np.float32(-1e9)).astype(np.float32)
    z1 = np.ones((len(base), 1), np.float32)
    pts = np.concatenate([base, np.concatenate([up, z1], 1), np.concatenate([dn, z1], 1)], 0)
    # same pixels seen from z=0.95 / 1.05 / behind (z=-1): exercises dist sign, clamp and mirrored hits
    more = []
    for z in (0.95, 1.05, 1.5, -1.0):
        q = base.copy()
        q[:, :2] *= z
        q[:, 2] = z
        more.append(q)
    return np.ascontiguousarray(np.concatenate([pts] + more, 0).astype(np.float32))


def cases():
    """name -> (scene kwargs | scene, points, return_names, mu, how)"""
    out = {}
    # BASELINE.json configs[0]: 10k grid points, 2 views 240x320, 64-dim features
    c = S.CONFIGS['cfg1']
    sc = S.make_scene(c['V'], c['H'], c['W'], seed=0, feat=c['feat'], num_inst=c['num_inst'], color=True)
    out['cfg1'] = dict(scene=sc, make=dict(V=c['V'], H=c['H'], W=c['W'], seed=0, feat=list(c['feat']),
                                           num_inst=c['num_inst'], color=True),
                       pts=S.config_points('cfg1'), pts_how='config_points:cfg1',
                       names=['dino_feats', 'mask', 'color_tensor'], mu=0.02)
    # cfg2-shaped views (4 x 480x640, map 48x64) with fewer channels; grid + scattered + adversarial points
    mk = dict(V=4, H=480, W=640, seed=1, feat=[48, 64, 128], num_inst=8, color=False)
    sc = S.make_scene(4, 480, 640, seed=1, feat=(48, 64, 128), num_inst=8)
    pts = np.concatenate([S.grid_points(30, 30, 30), S.scattered_points(12000, 1), S.adversarial_points(sc, 3)])
    out['mixed4v'] = dict(scene=sc, make=mk, pts=pts, pts_how='grid30+scattered12000(seed1)+adversarial(seed3)',
                          names=['dino_feats', 'mask'], mu=0.02)
    # odd sizes: C not a multiple of 4, non-multiple-of-16 image, 3 views, different mu
    mk = dict(V=3, H=97, W=131, seed=2, feat=[9, 13, 5], num_inst=3, color=True)
    sc = S.make_scene(3, 97, 131, seed=2, feat=(9, 13, 5), num_inst=3, color=True)
    pts = np.concatenate([S.grid_points(17, 13, 11), S.scattered_points(3001, 2), S.adversarial_points(sc, 5, 16)])
    out['odd3v'] = dict(scene=sc, make=mk, pts=pts, pts_how='grid17x13x11+scattered3001(seed2)+adversarial(seed5,16)',
                        names=['dino_feats', 'mask', 'color_tensor'], mu=0.05)
    # exact ties of round-half-to-even and exact cell borders of floor()
    out['ties'] = dict(scene=tie_scene(), make='tie_scene', pts=tie_points(), pts_how='tie_points',
                       names=['dino_feats', 'mask'], mu=0.02)
    # batch_eval across three 60 000-point chunks (reference fusion.py:526-545)
    mk = dict(V=4, H=240, W=320, seed=4, feat=[24, 32, 16], num_inst=4, color=False)
    sc = S.make_scene(4, 240, 320, seed=4, feat=(24, 32, 16), num_inst=4)
    out['batch3chunks'] = dict(scene=sc, make=mk, pts=S.grid_points(52, 50, 50), pts_how='grid52x50x50',
                               names=['dino_feats', 'mask'], mu=0.02, batch=True)
    # ---- CPU-only extras (tests/test_oracle_golden.py::EXTRA_CASES): more pinning of the restatements -------------
    mk = dict(V=1, H=64, W=80, seed=6, feat=[6, 8, 1024], num_inst=0, color=False)
    sc = S.make_scene(1, 64, 80, seed=6, feat=(6, 8, 1024))
    out['x_v1_c1024'] = dict(scene=sc, make=mk, pts=np.concatenate([S.grid_points(12, 12, 8), S.scattered_points(800, 6)]),
                             pts_how='grid12x12x8+scattered800(seed6)', names=['dino_feats'], mu=0.02)
    mk = dict(V=2, H=48, W=64, seed=7, feat=[48, 64, 4], num_inst=2, color=True)
    sc = S.make_scene(2, 48, 64, seed=7, feat=(48, 64, 4), num_inst=2, color=True)
    out['x_fullres_tinymu'] = dict(scene=sc, make=mk, pts=np.concatenate([S.grid_points(20, 20, 20), S.adversarial_points(sc, 7, 32)]),
                                   pts_how='grid20^3+adversarial(seed7,32)', names=['dino_feats', 'mask', 'color_tensor'], mu=0.002)
    mk = dict(V=5, H=90, W=120, seed=8, feat=[9, 12, 6], num_inst=0, color=False)
    sc = S.make_scene(5, 90, 120, seed=8, feat=(9, 12, 6))
    out['x_v5_far'] = dict(scene=sc, make=mk, pts=(S.scattered_points(6000, 8, sigma=1.5)), pts_how='scattered6000(seed8,sigma1.5)',
                           names=['dino_feats'], mu=0.1)
    return out


def run_reference(case):
    import torch
    sc = case['scene']
    F = RL.reference_fusion(sc, 'cpu')
    F.mu = case['mu']
    pts = torch.from_numpy(case['pts'])
    with torch.no_grad():
        if case.get('batch'):
            o = F.batch_eval(pts, return_names=list(case['names']))
            o = {k: v for k, v in o.items()}
        else:
            o = F.eval(pts, return_names=list(case['names']), return_inter=True)
        od = F.eval_dist(pts)
    o = {k: v.numpy() for k, v in o.items()}
    o['evaldist.dist'] = od['dist'].numpy()
    o['evaldist.valid_mask'] = od['valid_mask'].numpy()
    return o


SELECT_CASES = {
    # name -> scene kwargs, boundaries, grid resolution (None: explicit points), thresholds
    'select_grid': dict(make=dict(V=4, H=240, W=320, seed=11, feat=None, num_inst=5, color=False),
                        boundaries=dict(x_lower=-0.4, x_upper=0.4, y_lower=-0.4, y_upper=0.3, z_lower=-0.2, z_upper=0.3),
                        res=0.005, mu=0.02),
    'select_pcd': dict(make=dict(V=3, H=120, W=160, seed=12, feat=None, num_inst=8, color=False),
                       boundaries=None, res=None, mu=0.02),
}


def select_points(name, scene):
    """Explicit point cloud of the select_features_from_pcd case: jittered samples of the analytic surfaces."""
    rs = np.random.RandomState(4242)
    n = 40000
    th, ph = rs.uniform(0, 2 * np.pi, n), np.arccos(rs.uniform(0.0, 1.0, n))
    sph = 0.25 * np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], -1)
    pl = np.stack([rs.uniform(-0.4, 0.4, n), rs.uniform(-0.4, 0.4, n), np.full(n, -0.1)], -1)
    pts = np.concatenate([sph, pl], 0) + rs.normal(0, 0.004, (2 * n, 3))
    return np.ascontiguousarray(pts.astype(np.float32))


def run_reference_select(name):
    """The reference's own lines fusion.py:1420-1445 (1484-1501 for an explicit cloud), executed with its
    create_init_grid / Fusion.batch_eval, up to the per-instance boolean selection (FPS excluded)."""
    import torch
    c = SELECT_CASES[name]
    mk = c['make']
    sc = S.make_scene(mk['V'], mk['H'], mk['W'], seed=mk['seed'], feat=None, num_inst=mk['num_inst'])
    F = RL.reference_fusion(sc, 'cpu')
    F.mu = c['mu']
    ref = RL.load_reference()
    dist_threshold = 0.005
    if c['res'] is not None:
        grid, grid_shape = ref.create_init_grid(c['boundaries'], c['res'])
        grid = grid.to('cpu', dtype=torch.float32)
    else:
        grid, grid_shape = torch.from_numpy(select_points(name, sc)), None
    with torch.no_grad():
        out = F.batch_eval(grid, return_names=['mask'])
    dist_mask = torch.abs(out['dist']) < dist_threshold
    mask = out['mask']
    mask = mask / (mask.sum(dim=1, keepdim=True) + 1e-7)
    sel = {}
    for i in range(1, mk['num_inst']):
        instance_mask = mask[:, i] > 0.6
        sel[i] = torch.nonzero(instance_mask & dist_mask & out['valid_mask'])[:, 0].numpy().astype(np.int32)
    shell = dist_mask & out['valid_mask']
    margin = (mask[:, 1:] - 0.6).abs().min(dim=1).values
    near = torch.nonzero(shell & (margin < 2e-5))[:, 0].numpy().astype(np.int32)
    meta = dict(name=name, make=mk, boundaries=c['boundaries'], res=c['res'], mu=c['mu'], N=int(grid.shape[0]),
                grid_shape=list(grid_shape) if grid_shape is not None else None, torch=torch.__version__,
                reference='WangYixuan12/d3fields fusion.py:1420-1445 ops on create_init_grid + Fusion.batch_eval (unmodified, CPU)',
                input_sha256={'pts': _sha(grid.numpy()), 'depth': _sha(sc.depth), 'mask': _sha(sc.maps['mask'])},
                shell=int(shell.sum()), dist_sha256=_sha(out['dist'].numpy()), valid_sha256=_sha(out['valid_mask'].numpy()))
    blob = {'near': near, 'meta': np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for i, v in sel.items():
        blob[f'sel.{i}'] = v
    path = os.path.join(GOLDEN_DIR, name + '.npz')
    np.savez_compressed(path, **blob)
    print(f'{name}: N={grid.shape[0]} shell={int(shell.sum())} selected={[len(v) for v in sel.values()]} near={len(near)} '
          f'-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


def main():
    import torch
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if '--select-only' in sys.argv:
        for name in SELECT_CASES:
            run_reference_select(name)
        return
    for name, case in cases().items():
        sc = case['scene']
        ref = run_reference(case)
        N = case['pts'].shape[0]
        rows = np.unique(np.linspace(0, N - 1, min(ROWS, N)).astype(np.int64))
        blob = {}
        meta = dict(name=name, make=case['make'], pts_how=case['pts_how'], names=case['names'], mu=case['mu'],
                    N=int(N), H=sc.H, W=sc.W, V=sc.V, batch=bool(case.get('batch', False)),
                    torch=torch.__version__, reference='WangYixuan12/d3fields fusion.py (unmodified, CPU)',
                    input_sha256={'pts': _sha(case['pts']), 'pose': _sha(sc.pose), 'K': _sha(sc.K),
                                  'depth': _sha(sc.depth), **{k: _sha(v) for k, v in sc.maps.items()}},
                    output_sha256={}, checksum={})
        if case['make'] == 'tie_scene' or N <= 4096:
            blob['in.pts'] = case['pts']          # tiny hand-made cases carry their inputs verbatim
        blob['rows'] = rows
        for k, v in ref.items():
            if v.ndim == 1:                                  # dist / valid_mask (+ eval_dist variants)
                meta['output_sha256'][k] = _sha(v)
                blob['out.' + k] = v if N <= 20000 else v[rows]
                blob['full.' + k] = np.array(N <= 20000)
            elif k.endswith('_inter'):                       # (V,N,C)
                blob['out.' + k] = np.ascontiguousarray(v[:, rows])
                meta['checksum'][k] = [float(v.astype(np.float64).sum()), float(np.abs(v.astype(np.float64)).sum())]
            else:                                            # (N,C)
                blob['out.' + k] = np.ascontiguousarray(v[rows])
                meta['checksum'][k] = [float(v.astype(np.float64).sum()), float(np.abs(v.astype(np.float64)).sum())]
        blob['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(GOLDEN_DIR, name + '.npz')
        np.savez_compressed(path, **blob)
        print(f'{name}: N={N} valid={ref["valid_mask"].mean():.3f} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)')
    for name in SELECT_CASES:
        run_reference_select(name)


if __name__ == '__main__':
    main()
