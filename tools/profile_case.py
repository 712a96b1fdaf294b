#!/usr/bin/env python
"""One workload, a few launches — the command ncu wraps (profiles/README.md).

    ncu --set full --clock-control none --import-source on -k regex:field_tile --launch-skip 3 --launch-count 1 \
        -o gpurun_out/r02_visible python tools/profile_case.py visible

cases: grid (cfg2a), visible (1M grid points some view sees), none (1M points no view sees), scattered, binned
(scattered walked in bin order, order precomputed), cfg5 (256k keypoints), sweep (101.9M-point select), backward,
pca (eval_pca on 256k keypoints), dist16m (dist/valid only, 16M grid points), mask (cfg3: u8 instance-mask field).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3fields_b200 import Fusion, scene as S  # noqa: E402

DEV = 'cuda:0'


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else 'grid'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    V, H, W = 4, 480, 640
    sc = S.make_scene(V, H, W, seed=0, feat=(48, 64, 1024), num_inst=8)
    f = Fusion(num_cam=V, device=DEV)
    f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
    f.set_instance_masks(torch.from_numpy(sc.maps['mask']), as_uint8=True)
    names = ['dino_feats']
    kw = {}
    if case == 'grid':
        pts = torch.from_numpy(S.config_points('cfg2a')).to(DEV)
    elif case in ('visible', 'none'):
        g = torch.from_numpy(S.grid_points(400, 200, 200)).to(DEV)
        v = f.eval_dist(g)['valid_mask'] if False else f.eval(g, [])['valid_mask']
        sel = torch.nonzero(v if case == 'visible' else ~v)[:, 0][:1_000_000]
        pts = g[sel].contiguous()
        del g
    elif case in ('scattered', 'binned'):
        pts = torch.from_numpy(S.scattered_points(1_000_000, 0)).to(DEV)
        if case == 'binned':
            kw['binned'] = f.bin_order(pts)
    elif case == 'cfg5':
        pts = torch.from_numpy(S.scattered_points(262144, 0)).to(DEV)
    elif case == 'sweep':
        b = dict(x_lower=-0.4, x_upper=0.4, y_lower=-0.4, y_upper=0.3, z_lower=-0.2, z_upper=-0.019)
        for _ in range(reps):
            r = f.sweep_select(b, 0.001)
        torch.cuda.synchronize()
        print(case, r['count'])
        return
    elif case == 'pca':
        pts = torch.from_numpy(S.scattered_points(262144, 0)).to(DEV)
        comp = torch.randn(3, 1024, device=DEV); mean = torch.randn(1024, device=DEV)
        for _ in range(reps):
            r = f.eval_pca(pts, 'dino_feats', mean, comp)
        torch.cuda.synchronize()
        return
    elif case == 'dist16m':
        pts = torch.from_numpy(S.grid_points(400, 200, 200)).to(DEV)
        for _ in range(reps):
            r = f.eval(pts, [])
        torch.cuda.synchronize()
        return
    elif case == 'mask':
        pts = torch.from_numpy(S.config_points('cfg2a')).to(DEV)
        for _ in range(reps):
            r = f.eval(pts, ['mask'])
        torch.cuda.synchronize()
        return
    elif case == 'backward':
        pts = torch.from_numpy(S.scattered_points(262144, 0)).to(DEV)
        G = torch.randn(262144, 1024, device=DEV)
        for _ in range(reps):
            p = pts.clone().requires_grad_(True)
            (f.eval(p, names)['dino_feats'] * G).sum().backward()
        torch.cuda.synchronize()
        return
    else:
        raise SystemExit(f'unknown case {case}')
    for _ in range(reps):
        out = f.eval(pts, names, **kw)
    torch.cuda.synchronize()
    print(case, float(out['valid_mask'].float().mean()))


if __name__ == '__main__':
    main()
