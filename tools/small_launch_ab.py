"""Small launches: Fusion.eval timed (CUDA events, Python launch path included) for a few sizes, under the library's
tile-geometry thresholds (D3F_SMALL_TILE_N / D3F_TINY_TILE_N in the environment override them for A/B)."""
import os, sys, json, torch, numpy as np
sys.path.insert(0, ".")
from d3fields_b200 import Fusion, scene as S
sc = S.make_scene(4, 480, 640, seed=0, feat=(48, 64, 1024), num_inst=8)
f = Fusion(num_cam=4); f.update({"depth": sc.depth, "pose": sc.pose, "K": sc.K, "dino_feats": sc.maps["dino_feats"]})
f.set_instance_masks(torch.from_numpy(sc.maps['mask']), as_uint8=True)
def timed(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))*1e3
for n in [int(x) for x in os.environ.get("SIZES", "400,2000,8000,20000,32768,50000,100000").split(",")]:
    grid = torch.from_numpy(S.grid_points(100,100,100)[:n].copy()).cuda()
    scat = torch.from_numpy(S.scattered_points(n, 3, sigma=0.15)).cuda()
    row = {'n': n, 'small_tile_threshold': os.environ.get('D3F_SMALL_TILE_N', '100000 (library default)')}
    row['grid_desc_us'] = timed(lambda: f.eval(grid, ['dino_feats']))
    row['scattered_desc_us'] = timed(lambda: f.eval(scat, ['dino_feats']))
    row['scattered_3keys_us'] = timed(lambda: f.eval(scat, ['dino_feats', 'mask']))
    row['dist_only_us'] = timed(lambda: f.eval(scat, []))
    print(json.dumps(row), flush=True)
