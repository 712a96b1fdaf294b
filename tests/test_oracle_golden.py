"""CPU: the oracle restatement (oracle/field_oracle.py) against the golden vectors produced by the
unmodified reference (oracle/gen_golden.py).  This is what pins the oracle; everything on the
GPU box is then compared with the oracle and with the same fixtures."""
import numpy as np
import pytest

from golden_util import CASES, EXTRA_CASES, SELECT_CASES, Golden, SelectGolden
from oracle import field_oracle as O


@pytest.mark.parametrize('name', CASES + EXTRA_CASES)
def test_numpy_oracle_matches_reference_golden(name):
    g = Golden(name)
    sc = g.scene
    inter = not g.meta['batch']
    out = O.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, g.names, mu=g.mu,
                       return_inter=inter)
    g.check_exact('dist', out['dist'])
    g.check_exact('valid_mask', out['valid_mask'])
    for k in g.names:
        g.check_close(k, out[k])
        if inter:
            g.check_close(k + '_inter', out[k + '_inter'])
    od = O.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, mu=g.mu, eval_dist=True)
    g.check_exact('evaldist.dist', od['dist'])
    g.check_exact('evaldist.valid_mask', od['valid_mask'])


def test_oracle_empty_and_single_point():
    g = Golden('ties')
    sc = g.scene
    out = O.field_eval(np.zeros((0, 3), np.float32), sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, ['dino_feats'])
    assert out['dist'].shape == (0,) and out['valid_mask'].shape == (0,) and out['dino_feats'].shape[0] == 0
    one = O.field_eval(g.pts[:1], sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, ['dino_feats'])
    full = O.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, ['dino_feats'])
    assert np.array_equal(one['dino_feats'][0], full['dino_feats'][0])


def test_uint8_mask_is_read_as_its_float_value():
    g = Golden('odd3v')
    sc = g.scene
    m8 = {'mask': sc.maps['mask'].astype(np.uint8)}
    a = O.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, m8, ['mask'], mu=g.mu)
    b = O.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, ['mask'], mu=g.mu)
    assert np.array_equal(a['mask'], b['mask'])


@pytest.mark.parametrize('name', CASES + EXTRA_CASES)
def test_torch_port_matches_reference_golden(name):
    """oracle/torch_port.py (the CPU-baseline stand-in for the reference) gives the reference's results."""
    import torch
    from oracle import torch_port as TP
    g = Golden(name)
    sc = g.scene
    obs = TP.obs_from_scene(sc)
    pts = torch.from_numpy(g.pts)
    out = TP.batch_eval(obs, sc.H, sc.W, pts, g.names, mu=g.mu)
    g.check_exact('dist', out['dist'].numpy())
    g.check_exact('valid_mask', out['valid_mask'].numpy())
    for k in g.names:
        g.check_close(k, out[k].numpy())


@pytest.mark.parametrize('name', CASES + EXTRA_CASES)
def test_c_oracle_matches_reference_golden_and_numpy_oracle(name):
    """oracle/d3f_oracle.c: pinned against the reference's golden vectors, and equal to the numpy
    restatement (bit-for-bit on dist/valid and on the per-view samples; the weighted mean only differs
    through libm's expf vs numpy's exp)."""
    from oracle import c_oracle as CO
    g = Golden(name)
    sc = g.scene
    inter = not g.meta['batch']
    out = CO.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, g.names, mu=g.mu, return_inter=inter)
    g.check_exact('dist', out['dist'])
    g.check_exact('valid_mask', out['valid_mask'])
    for k in g.names:
        g.check_close(k, out[k])
        if inter:
            g.check_close(k + '_inter', out[k + '_inter'])
    od = CO.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, mu=g.mu, eval_dist=True)
    g.check_exact('evaldist.dist', od['dist'])
    g.check_exact('evaldist.valid_mask', od['valid_mask'])
    if g.pts.shape[0] <= 40000:
        ref = O.field_eval(g.pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, g.names, mu=g.mu, return_inter=inter)
        for k in g.names:
            if inter:
                assert np.array_equal(out[k + '_inter'], ref[k + '_inter']), k
            assert np.abs(out[k] - ref[k]).max() <= 1e-6 * max(1.0, np.abs(ref[k]).max())


@pytest.mark.parametrize('name', SELECT_CASES)
def test_select_candidates_oracle_matches_reference_selection(name):
    """oracle/field_oracle.select_candidates (restating fusion.py:1428-1445) picks exactly the points the unmodified
    reference picked; the point coordinates of init_grid are the reference's create_init_grid bytes."""
    from oracle import c_oracle as CO
    g = SelectGolden(name)
    sc = g.scene
    pts = g.points()
    idx, inst, margin = O.select_candidates(pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps['mask'], mu=g.mu,
                                            field_fn=CO.field_eval)
    g.check(idx, inst)
    assert len(idx) == sum(len(v) for v in g.sel.values())
    assert np.isfinite(margin).sum() == g.meta['shell']
    if name == 'select_pcd':                       # the numpy restatement too, on the small case
        idx2, inst2, _ = O.select_candidates(pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps['mask'], mu=g.mu)
        assert np.array_equal(idx, idx2) and np.array_equal(inst, inst2)
