import sys, json, torch, numpy as np
sys.path.insert(0, ".")
from d3fields_b200 import Fusion, _native, scene as S
sc = S.make_scene(4, 480, 640, seed=0, feat=(48, 64, 1024))
f = Fusion(num_cam=4); f.update({"depth": sc.depth, "pose": sc.pose, "K": sc.K, "dino_feats": sc.maps["dino_feats"]})
pts = torch.from_numpy(S.config_points('cfg2a')).cuda()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
ms = timed(lambda: f.eval(pts, ['dino_feats'], return_inter=True))
print(json.dumps({'name': 'cfg2a_return_inter', 'ms': ms, 'variant': _native.last_variant(0), 'bytes_out_gb': 1e6*(5*4096+5)/1e9, 'floor_ms_at_measured_peak': 1e6*(5*4096)/6558.1e9*1e3}))
