#!/usr/bin/env python
"""A short fused tracking run for `ncu --metrics gpu__time_duration.sum` (per-kernel times of one Adam iteration)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3fields_b200 import Fusion, scene as S
from d3fields_b200.tracking import FusedRigidTracker
sc = S.make_scene(4, 480, 640, seed=0, feat=(48, 64, 1024))
f = Fusion(num_cam=4); f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
I, P = 4, 100
pts = torch.from_numpy(S.scattered_points(I * P, 23, sigma=0.12)).cuda().reshape(I, P, 3)
src = f.eval(pts.reshape(-1, 3), return_names=['dino_feats'])['dino_feats']
tr = FusedRigidTracker(f, I, P, 1024, iters=10, graph=False)
tr.track(src, pts + 0.004)
torch.cuda.synchronize()
