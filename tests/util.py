"""Helpers shared by the CPU and GPU tests."""
import numpy as np
import torch

from d3fields_b200 import Fusion


def make_fusion(scene, device='cuda:0', mask_u8=False, mu=0.02):
    """A d3fields_b200.Fusion holding `scene` (d3fields_b200.scene.Scene) through the public update() path."""
    f = Fusion(num_cam=scene.V, device=device)
    obs = {'depth': scene.depth, 'pose': scene.pose, 'K': scene.K}
    if 'dino_feats' in scene.maps:
        obs['dino_feats'] = scene.maps['dino_feats']
    f.update(obs)
    for k, v in scene.maps.items():
        if k == 'dino_feats':
            continue
        t = torch.from_numpy(v).to(device)
        if k == 'mask' and mask_u8:
            t = t.to(torch.uint8)
        f.curr_obs_torch[k] = t.contiguous()
    f.mu = mu
    return f


def assert_close_field(got, ref, rtol=1e-4, atol_scale=2e-6, what=''):
    """|a-b| <= rtol*|b| + atol_scale*max|b|  — the descriptor tolerance of BASELINE.json's north_star
    (1e-4 relative) with a floor for sums of signed terms that cancel (SURVEY.md §7 'tolerance definition')."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    tol = rtol * np.abs(ref) + atol_scale * scale
    bad = err > tol
    assert not bad.any(), f'{what}: {bad.sum()} of {bad.size} outside tolerance; max err {err.max():.3e} (scale {scale:.3g})'


def assert_bits_equal(got, ref, what=''):
    got = np.ascontiguousarray(got)
    ref = np.ascontiguousarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.dtype == np.float32:
        nbad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
    else:
        nbad = int((got != ref).sum())
    assert nbad == 0, f'{what}: {nbad} of {got.size} elements differ bitwise'
