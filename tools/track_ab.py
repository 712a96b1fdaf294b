#!/usr/bin/env python
"""us per Adam iteration of FusedRigidTracker as one CUDA-graph replay: the single-launch iteration (d3f_track_step)
against the four-launch one, for a few (instances, points) sizes.  D3F_STEP_MINB=2|3 pins the occupancy variant.
One JSON line per size."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3fields_b200 import Fusion, scene as S
from d3fields_b200.tracking import FusedRigidTracker

sc = S.make_scene(4, 480, 640, seed=0, feat=(48, 64, 1024))
f = Fusion(num_cam=4)
f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
for I, P in ((2, 100), (4, 100), (8, 100)):
    pts = torch.from_numpy(S.scattered_points(I * P, 23, sigma=0.12)).cuda().reshape(I, P, 3)
    src = f.eval(pts.reshape(-1, 3), return_names=['dino_feats'])['dino_feats']
    row = {'instances': I, 'points': I * P, 'minb_env': os.environ.get('D3F_STEP_MINB', '')}
    res = {}
    for name, single in (('single_launch', True), ('four_launches', False)):
        tr = FusedRigidTracker(f, I, P, 1024, iters=100, single_launch=single)
        res[name] = tr.track(src, pts + 0.004)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            tr.track(src, pts + 0.004)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        row[name + '_us_per_iteration'] = round(best / 100 * 1e6, 2)
    row['max_abs_dt_between_forms'] = float((res['single_launch']['t'] - res['four_launches']['t']).abs().max())
    print(json.dumps(row), flush=True)
