#!/usr/bin/env python
"""Secondary measurements for DESIGN.md (not the driver's bench contract): every BASELINE.json config that
fits one GPU, timed with CUDA events through the public Fusion API, plus the reference's operator sequence
run on the same GPU by torch (oracle/torch_port.py on device='cuda') as the "torch-GPU" column.

    python tools/bench_matrix.py [--out gpurun_out/matrix.jsonl] [--only name,...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3fields_b200 import Fusion, _native, scene as S  # noqa: E402

DEV = 'cuda:0'


def timed(fn, warm=3, reps=10, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def alg_bytes(n, V, H, W, keys):
    b = 12 * n + min(V * H * W * 4, 4 * n * V) + n * 5
    for (h, w, C, s) in keys:
        b += min(V * h * w * C * s, 4 * n * V * C * s) + n * 4 * C
    return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'matrix.jsonl'))
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    only = set(x for x in args.only.split(',') if x)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    V, H, W = 4, 480, 640
    sc = S.make_scene(V, H, W, seed=0, feat=(48, 64, 1024), num_inst=8, color=True)
    f = Fusion(num_cam=V, device=DEV)
    f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
    f.curr_obs_torch['mask'] = torch.from_numpy(sc.maps['mask']).to(DEV)
    f.curr_obs_torch['mask_u8'] = torch.from_numpy(sc.maps['mask']).to(DEV).to(torch.uint8)
    f.curr_obs_torch['color_tensor'] = torch.from_numpy(sc.maps['color_tensor']).to(DEV)
    grid1m = torch.from_numpy(S.config_points('cfg2a')).to(DEV)
    scat1m = torch.from_numpy(S.scattered_points(1_000_000, 0)).to(DEV)
    scat256k = scat1m[:262144].contiguous()
    grid16m = torch.from_numpy(S.grid_points(400, 200, 200)).to(DEV)
    results = []

    def rec(name, n, ms, ms_min, keys, note=''):
        B = alg_bytes(n, V, H, W, keys)
        r = dict(name=name, n=n, ms=ms, ms_min=ms_min, mpts_s=n / ms / 1e3, alg_bytes=B, gbs=B / ms / 1e6,
                 frac_of_measured_hbm=B / ms / 1e6 / peak, variant=_native.last_variant(0), note=note)
        results.append(r)
        print(json.dumps(r), flush=True)

    def want(name):
        return not only or name in only

    if want('cfg2a_grid'):
        ms, mn = timed(lambda: f.eval(grid1m, ['dino_feats']), flush=flush)
        rec('cfg2a_grid', 1_000_000, ms, mn, [(48, 64, 1024, 4)])
    if want('cfg2a_scattered'):
        ms, mn = timed(lambda: f.eval(scat1m, ['dino_feats']), flush=flush)
        rec('cfg2a_scattered', 1_000_000, ms, mn, [(48, 64, 1024, 4)], 'N(0,0.25^2) keypoints, no locality')
    if want('binned'):
        # keypoints without spatial order, walked in lattice-cell order (d3f_bin_order + d3f_eval_ordered)
        ms, mn = timed(lambda: f.eval(scat1m, ['dino_feats'], binned=True), flush=flush)
        rec('cfg2a_scattered_binned', 1_000_000, ms, mn, [(48, 64, 1024, 4)], 'bin_order (6 launches) + ordered walk, both timed')
        for cell in (0.005, 0.01, 0.02):
            ms, mn = timed(lambda: f.eval(scat1m, ['dino_feats'], binned=cell), flush=flush)
            rec(f'cfg2a_scattered_binned_cell{cell}', 1_000_000, ms, mn, [(48, 64, 1024, 4)])
        order = f.bin_order(scat1m)
        ms, mn = timed(lambda: f.eval(scat1m, ['dino_feats'], binned=order), flush=flush)
        rec('cfg2a_scattered_preordered', 1_000_000, ms, mn, [(48, 64, 1024, 4)], 'order computed once outside the timed call')
        ms, mn = timed(lambda: f.bin_order(scat1m), flush=flush)
        rec('bin_order_1m', 1_000_000, ms, mn, [], 'd3f_bin_order alone')
        ms, mn = timed(lambda: f.eval(scat256k, ['dino_feats'], binned=True), flush=flush)
        rec('cfg5_frame_256k_scattered_binned', 262144, ms, mn, [(48, 64, 1024, 4)])
        ms, mn = timed(lambda: f.eval(grid1m, ['dino_feats'], binned=True), flush=flush)
        rec('cfg2a_grid_binned', 1_000_000, ms, mn, [(48, 64, 1024, 4)], 'a grid gains nothing from binning (already ordered)')
    if want('visible'):
        # a scene where most points are seen by some view: points in a shell around the surfaces, z-fastest order kept
        d = f.eval(grid16m, [])
        keep = torch.nonzero(d['valid_mask'])[:, 0][:1_000_000]
        vis1m = grid16m[keep].contiguous()
        frac = float(f.eval(vis1m, [])['valid_mask'].float().mean())
        ms, mn = timed(lambda: f.eval(vis1m, ['dino_feats']), flush=flush)
        rec('cfg2a_grid_all_visible', 1_000_000, ms, mn, [(48, 64, 1024, 4)], f'first 1M grid points (of 16M, z fastest) that some view sees: valid fraction {frac:.3f}')
        none1m = grid16m[torch.nonzero(~d['valid_mask'])[:, 0][:1_000_000]].contiguous()
        ms, mn = timed(lambda: f.eval(none1m, ['dino_feats']), flush=flush)
        rec('cfg2a_grid_none_visible', 1_000_000, ms, mn, [(48, 64, 1024, 4)], 'points no view sees: pure zero-row stream')
    if want('backward'):
        kp = torch.from_numpy(S.scattered_points(800, 5, sigma=0.15)).to(DEV)
        G = torch.randn(800, 1024, device=DEV)
        def fb():
            p = kp.clone().requires_grad_(True)
            o = f.eval(p, ['dino_feats'])
            ((o['dino_feats'] * G).sum() + o['dist'].sum()).backward()
        ms, mn = timed(fb, warm=5, reps=20)
        r = dict(name='tracking_fwd_bwd_800pts_eager', n=800, ms=ms, ms_min=mn, note='autograd Function forward + d3f_eval_backward + torch loss, eager launches')
        results.append(r); print(json.dumps(r), flush=True)
        big = scat256k.clone()
        Gb = torch.randn(262144, 1024, device=DEV)
        def fb2():
            p = big.clone().requires_grad_(True)
            o = f.eval(p, ['dino_feats'])
            (o['dino_feats'] * Gb).sum().backward()
        ms, mn = timed(fb2, warm=2, reps=5)
        r = dict(name='fwd_bwd_256k', n=262144, ms=ms, ms_min=mn, note='forward + backward kernels + torch mul/sum on 256k keypoints')
        results.append(r); print(json.dumps(r), flush=True)
    if want('sweep_select'):
        f.curr_obs_torch['mask'] = f.curr_obs_torch['mask_u8']
        b = dict(x_lower=-0.4, x_upper=0.4, y_lower=-0.4, y_upper=0.3, z_lower=-0.2, z_upper=-0.019)
        f.sweep_select(b, 0.001)
        torch.cuda.synchronize()
        # kernel only: zero the counter + launch, no host read-back
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r_ = f.sweep_select(b, 0.001)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        nbig = int(r_['grid_shape'][0] * r_['grid_shape'][1] * r_['grid_shape'][2])
        r = dict(name='sweep_select_101m_u8', n=nbig, ms=float(np.median(ts)), ms_min=float(min(ts)), mpts_s=nbig / float(np.median(ts)) / 1e3,
                 survivors=int(r_['count']), note='d3f_sweep_select end to end (wall clock incl. count read-back, sort of survivors): no grid, no dense mask field in HBM')
        results.append(r); print(json.dumps(r), flush=True)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r_ = f.sweep_select(b, 0.001, mask_name=None, dense=True)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        r = dict(name='sweep_dense_dist_101m', n=nbig, ms=float(np.median(ts)), ms_min=float(min(ts)), mpts_s=nbig / float(np.median(ts)) / 1e3,
                 note='dense dist/valid from the linear index (extract_mesh input), no pts in HBM')
        results.append(r); print(json.dumps(r), flush=True)
        f.curr_obs_torch['mask'] = torch.from_numpy(sc.maps['mask']).to(DEV)
    if want('cfg5_frame'):
        ms, mn = timed(lambda: f.eval(scat256k, ['dino_feats']), flush=flush)
        rec('cfg5_frame_256k_scattered', 262144, ms, mn, [(48, 64, 1024, 4)])
        comp = torch.randn(3, 1024, device=DEV); mean = torch.randn(1024, device=DEV)
        ms, mn = timed(lambda: f.pca_project(f.eval(scat256k, ['dino_feats'])['dino_feats'], mean, comp), flush=flush)
        rec('cfg5_frame_256k_eval+pca3', 262144, ms, mn, [(48, 64, 1024, 4)], 'eval then separate PCA(3) kernel')
        ms, mn = timed(lambda: f.eval_pca(scat256k, 'dino_feats', mean, comp), flush=flush)
        r = dict(name='cfg5_frame_256k_pca3_projected_volume', n=262144, ms=ms, ms_min=mn, mpts_s=262144 / ms / 1e3,
                 variant=_native.last_variant(0), note='PCA(3) of the field via the projected volume (narrow path): 12 B/pt out instead of 4 KB/pt')
        results.append(r); print(json.dumps(r), flush=True)
        ms, mn = timed(lambda: f.eval_pca(grid1m, 'dino_feats', mean, comp), flush=flush)
        r = dict(name='cfg2a_grid_pca3_projected_volume', n=1000000, ms=ms, ms_min=mn, mpts_s=1000000 / ms / 1e3,
                 variant=_native.last_variant(0), note='1M grid points, PCA(3) via projected volume')
        results.append(r); print(json.dumps(r), flush=True)
    if want('cfg3'):
        ms, mn = timed(lambda: f.eval(grid1m, ['mask']), flush=flush)
        rec('cfg3_mask_f32', 1_000_000, ms, mn, [(480, 640, 8, 4)])
        ms, mn = timed(lambda: f.eval(grid1m, ['mask_u8']), flush=flush)
        rec('cfg3_mask_u8', 1_000_000, ms, mn, [(480, 640, 8, 1)])
        ms, mn = timed(lambda: f.eval(scat1m, ['mask_u8']), flush=flush)
        rec('cfg3_mask_u8_scattered', 1_000_000, ms, mn, [(480, 640, 8, 1)])
    if want('dist_only'):
        ms, mn = timed(lambda: f.eval(grid1m, []), flush=flush)
        rec('dist_only_1m', 1_000_000, ms, mn, [])
        ms, mn = timed(lambda: f.eval(grid16m, []), flush=flush)
        rec('dist_only_16m', 16_000_000, ms, mn, [])
    if want('sweep_101m'):
        # select_features_rand's sweep (reference fusion.py:1420-1428): 800x700x181 grid at 1 mm, batch_eval(grid, ['mask'])
        from d3fields_b200 import create_init_grid_device
        b = dict(x_lower=-0.4, x_upper=0.4, y_lower=-0.4, y_upper=0.3, z_lower=-0.2, z_upper=-0.019)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        big, shape = create_init_grid_device(b, 0.001, DEV)
        torch.cuda.synchronize(); t_grid = time.perf_counter() - t0
        nbig = big.shape[0]
        ms, mn = timed(lambda: f.eval(big, ['mask_u8']), warm=1, reps=3)
        r = dict(name='sweep_101m_grid_mask_u8', n=nbig, ms=ms, ms_min=mn, mpts_s=nbig / ms / 1e3, variant=_native.last_variant(0),
                 note=f'grid {tuple(shape)} generated on device in {t_grid * 1e3:.1f} ms; reference: 1690 chunks of 60 000')
        results.append(r); print(json.dumps(r), flush=True)
        ms, mn = timed(lambda: f.eval(big, []), warm=1, reps=3)
        r = dict(name='sweep_101m_grid_dist_only', n=nbig, ms=ms, ms_min=mn, mpts_s=nbig / ms / 1e3, variant=_native.last_variant(0), note='')
        results.append(r); print(json.dumps(r), flush=True)
        del big
    if want('multi_key'):
        ms, mn = timed(lambda: f.eval(grid1m, ['dino_feats', 'mask', 'color_tensor']), flush=flush)
        rec('vis_repr_3keys', 1_000_000, ms, mn, [(48, 64, 1024, 4), (480, 640, 8, 4), (480, 640, 3, 4)],
            'dino_feats+mask+color_tensor in one launch (reference vis_repr.py:103)')
    if want('v8'):
        # eight views: the NV = 8 tile instantiation (64-channel slices); D3F_FORCE_GENERIC=1 in the environment gives the
        # generic kernel's time for the same launch
        sc8 = S.make_scene(8, H, W, seed=0, feat=(48, 64, 1024))
        f8 = Fusion(num_cam=8, device=DEV)
        f8.update({'depth': sc8.depth, 'pose': sc8.pose, 'K': sc8.K, 'dino_feats': sc8.maps['dino_feats']})
        ms, mn = timed(lambda: f8.eval(grid1m, ['dino_feats']), flush=flush)
        B8 = 12 * 1_000_000 + 8 * H * W * 4 + 8 * 48 * 64 * 1024 * 4 + 1_000_000 * (5 + 4096)
        r = dict(name='v8_grid_1m', n=1_000_000, ms=ms, ms_min=mn, mpts_s=1_000_000 / ms / 1e3, alg_bytes=B8, gbs=B8 / ms / 1e6,
                 frac_of_measured_hbm=B8 / ms / 1e6 / peak, variant=_native.last_variant(0), valid=float(f8.eval(grid1m, [])['valid_mask'].float().mean()),
                 note='V=8 ring cameras, 1M grid points, C=1024 @ (48,64)')
        results.append(r); print(json.dumps(r), flush=True)
        del f8
    if want('cfg2b'):
        vol = f.curr_obs_torch['dino_feats']
        try:
            f.curr_obs_torch['dino_feats'] = torch.randn((V, 480, 640, 1024), device=DEV)
            ms, mn = timed(lambda: f.eval(grid1m, ['dino_feats']), flush=flush, reps=5)
            rec('cfg2b_fullres_volume', 1_000_000, ms, mn, [(480, 640, 1024, 4)], '5.03 GB volume, map (480,640)')
        finally:
            f.curr_obs_torch['dino_feats'] = vol
    if want('torch_gpu'):
        from oracle import torch_port as TP
        obs = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in TP.obs_from_scene(sc).items()}
        ms, mn = timed(lambda: TP.batch_eval(obs, H, W, grid1m, ['dino_feats']), warm=1, reps=3)
        rec('torch_gpu_reference_ops_cfg2a', 1_000_000, ms, mn, [(48, 64, 1024, 4)],
            'oracle/torch_port.py (reference operator sequence, 60k chunks) on the same B200')
    with open(args.out, 'w') as fh:
        for r in results:
            fh.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
