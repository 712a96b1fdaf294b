"""CPU: the Python mirror of the reference interface — argument contracts and error behaviour
(reference fusion.py:313-320), create_init_grid, and that nothing silently falls back to CPU math."""
import numpy as np
import pytest
import torch

import d3fields_b200
from d3fields_b200 import Fusion, create_init_grid, scene as S
from oracle import ref_loader as RL


def test_create_init_grid_matches_reference_construction():
    b = S.WORKSPACE
    pts, shape = create_init_grid(b, 0.02)
    assert tuple(shape) == (40, 35, 11) and pts.shape == (40 * 35 * 11, 3) and pts.dtype == torch.float32
    # z fastest, then y, then x (reference fusion.py:86-87)
    assert pts[1, 2] > pts[0, 2] and pts[1, 0] == pts[0, 0] and pts[11, 1] > pts[0, 1] and pts[11, 2] == pts[0, 2]
    if RL.reference_available():
        rp, rs = RL.load_reference().create_init_grid(b, 0.02)
        assert tuple(rs) == tuple(shape) and torch.equal(rp, pts)


def test_eval_before_update_prints_and_exits(capsys):
    f = Fusion(num_cam=4)
    with pytest.raises(SystemExit):
        f.eval(torch.zeros(5, 3))
    assert 'Please call update() first!' in capsys.readouterr().out


def test_eval_rejects_bad_points_like_the_reference():
    f = Fusion(num_cam=1)
    f.curr_obs_torch['depth'] = torch.zeros(1, 4, 4)          # non-empty -> passes the update() check
    with pytest.raises(AssertionError):
        f.eval(np.zeros((5, 3), np.float32))
    with pytest.raises(AssertionError):
        f.eval(torch.zeros(5))
    with pytest.raises(AssertionError):
        f.eval(torch.zeros(5, 4))
    with pytest.raises(ValueError):
        f.eval(torch.zeros(5, 3, dtype=torch.float64))


def test_cpu_observation_is_refused_not_emulated():
    """No CPU fallback: an observation that is not on a CUDA device is an error, never a slow path."""
    sc = S.make_scene(2, 24, 32, seed=0, feat=(4, 6, 8))
    f = Fusion(num_cam=2, device='cpu')
    f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
    with pytest.raises(ValueError, match='CUDA'):
        f.eval(torch.zeros(5, 3), return_names=['dino_feats'])


def test_unknown_key_raises_keyerror_like_the_reference():
    sc = S.make_scene(2, 24, 32, seed=0)
    f = Fusion(num_cam=2, device='cpu')
    f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K})
    f.curr_obs_torch['depth'] = f.curr_obs_torch['depth']
    with pytest.raises((KeyError, ValueError)):
        f.eval(torch.zeros(5, 3), return_names=['nope'])


def test_update_layout_and_pose_4x4_accepted():
    sc = S.make_scene(3, 20, 30, seed=1, feat=(2, 3, 4))
    f = Fusion(num_cam=3, device='cpu')
    pose44 = np.concatenate([sc.pose, np.tile(np.array([[[0, 0, 0, 1]]], np.float32), (3, 1, 1))], 1)
    color = np.zeros((3, 20, 30, 3), np.uint8)
    f.update({'color': color, 'depth': sc.depth, 'pose': pose44, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
    o = f.curr_obs_torch
    assert o['pose'].shape == (3, 3, 4) and o['K'].shape == (3, 3, 3) and o['depth'].shape == (3, 20, 30)
    assert o['color_tensor'].shape == (3, 20, 30, 3) and o['dino_feats'].shape == (3, 2, 3, 4)
    assert (f.H, f.W, f.num_cam) == (20, 30, 3)


def test_mask_injection_and_perception_delegation():
    f = Fusion(num_cam=2, device='cpu')
    lab = torch.randint(0, 3, (2, 8, 8), dtype=torch.uint8)
    f.set_instance_masks(lab, labels=['bg', 'mug', 'fork'])
    assert f.curr_obs_torch['mask'].shape == (2, 8, 8, 3) and f.get_inst_num() == 3
    assert torch.equal(f.curr_obs_torch['mask'].argmax(-1).to(torch.uint8), lab)
    with pytest.raises(RuntimeError, match='front-end'):
        f.text_queries_for_inst_mask(['mug'], [0.3], S.WORKSPACE)

    class Fake:
        def text_queries_for_inst_mask_no_track(self, fusion, queries, thresholds, boundaries, **kw):
            return lab, ['bg'] + list(queries) + ['x']
    f2 = Fusion(num_cam=2, device='cpu', perception=Fake())
    f2.text_queries_for_inst_mask_no_track(['mug'], [0.3], S.WORKSPACE)
    assert f2.curr_obs_torch['consensus_mask_label'] == ['bg', 'mug', 'x']


def test_package_exports():
    for n in ('Fusion', 'create_init_grid', 'project_points_coords', 'interpolate_feats'):
        assert hasattr(d3fields_b200, n)


def test_module_level_helpers_match_the_oracle_and_the_reference():
    """project_points_coords / interpolate_feats are kept for callers that import them (reference fusion.py:32-77)."""
    from d3fields_b200 import interpolate_feats, project_points_coords
    from oracle import field_oracle as O
    sc = S.make_scene(3, 40, 56, seed=9, feat=(5, 7, 6))
    pts = np.concatenate([S.scattered_points(400, 9, sigma=0.3), S.adversarial_points(sc, 9, 4)])
    p2, ok, z = project_points_coords(torch.from_numpy(pts), torch.from_numpy(sc.pose), torch.from_numpy(sc.K))
    o2, ook, oz = O.project(pts, sc.pose, sc.K)
    assert np.array_equal(p2.numpy().view(np.uint32), o2.view(np.uint32)) and np.array_equal(ok.numpy(), ook)
    assert np.array_equal(z[..., 0].numpy().view(np.uint32), oz.view(np.uint32))
    vol = torch.from_numpy(sc.maps['dino_feats']).permute(0, 3, 1, 2)
    smp = interpolate_feats(vol, p2, h=40, w=56, padding_mode='zeros', align_corners=True, inter_mode='bilinear')
    ref = O.sample_bilinear(sc.maps['dino_feats'], o2, 40, 56)
    assert np.abs(smp.numpy() - ref).max() <= 1e-6
    if RL.reference_available():
        rf = RL.load_reference()
        r2, rok, rz = rf.project_points_coords(torch.from_numpy(pts), torch.from_numpy(sc.pose), torch.from_numpy(sc.K))
        assert torch.equal(r2, p2) and torch.equal(rok, ok) and torch.equal(rz, z)
        rs = rf.interpolate_feats(vol, r2, h=40, w=56, padding_mode='zeros', align_corners=True, inter_mode='bilinear')
        assert torch.equal(rs, smp)
