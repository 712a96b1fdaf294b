#!/usr/bin/env python
"""Per-rank work of the weak-scaling benchmark, measured on ONE GPU: time Fusion.eval on each rank's shard of the
world-times finer grid, for contiguous x-slabs and for plane-interleaved sharding."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3fields_b200 import Fusion, scene as S

def timed(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c = S.CONFIGS['cfg2a']
sc = S.make_scene(c['V'], c['H'], c['W'], seed=0, feat=c['feat'])
f = Fusion(num_cam=4, device='cuda:0')
f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
gx, gy, gz = c['grid']
full = S.grid_points(gx * world, gy, gz).reshape(gx * world, gy * gz, 3)
res = {'contiguous': [], 'interleaved': []}
for r in range(world):
    for mode in res:
        pts = full[r * gx:(r + 1) * gx] if mode == 'contiguous' else full[r::world]
        p = torch.from_numpy(np.ascontiguousarray(pts.reshape(-1, 3))).cuda()
        ms = timed(lambda: f.eval(p, ['dino_feats']))
        valid = f.eval(p, [])['valid_mask'].float().mean().item()
        res[mode].append((round(ms, 4), round(valid, 3)))
print(json.dumps(res))
for mode, v in res.items():
    t = [x[0] for x in v]
    print(mode, 'ms min/mean/max', min(t), round(sum(t) / len(t), 4), max(t))
