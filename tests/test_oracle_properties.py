"""CPU: property tests of the oracle restatements on random scenes (hypothesis): the C restatement equals the numpy
restatement bit-for-bit on everything that does not pass through exp(), and the field has the structural
properties the GPU tests rely on at full size (point independence, linearity in the map, zero rows)."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from d3fields_b200 import scene as S
from oracle import c_oracle as CO, field_oracle as O


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(V=st.integers(1, 6), H=st.integers(8, 40), W=st.integers(8, 50), h=st.integers(1, 9), w=st.integers(1, 9),
       C=st.integers(1, 9), seed=st.integers(0, 10_000), mu=st.sampled_from([0.002, 0.02, 0.1]))
def test_c_and_numpy_restatements_agree(V, H, W, h, w, C, seed, mu):
    sc = S.make_scene(V, H, W, seed=seed, feat=(h, w, C), num_inst=2)
    pts = np.concatenate([S.scattered_points(300, seed, sigma=0.3), S.adversarial_points(sc, seed, 3)])
    a = O.field_eval(pts, sc.pose, sc.K, sc.depth, H, W, sc.maps, ['dino_feats', 'mask'], mu=mu, return_inter=True)
    b = CO.field_eval(pts, sc.pose, sc.K, sc.depth, H, W, sc.maps, ['dino_feats', 'mask'], mu=mu, return_inter=True)
    assert np.array_equal(a['valid_mask'], b['valid_mask'])
    assert np.array_equal(a['dist'].view(np.uint32), b['dist'].view(np.uint32))
    for k in ('dino_feats', 'mask'):
        assert np.array_equal(a[k + '_inter'], b[k + '_inter'])
        scale = max(1.0, float(np.abs(a[k]).max()))
        assert np.abs(a[k] - b[k]).max() <= 1e-6 * scale           # only exp() differs (libm vs numpy), by an ulp
        assert (a[k][~a['valid_mask']] == 0).all()
    assert (a['dist'][~a['valid_mask']] == np.float32(1e3)).all()
    ad = O.field_eval(pts, sc.pose, sc.K, sc.depth, H, W, mu=mu, eval_dist=True)
    bd = CO.field_eval(pts, sc.pose, sc.K, sc.depth, H, W, mu=mu, eval_dist=True)
    assert np.array_equal(ad['dist'].view(np.uint32), bd['dist'].view(np.uint32)) and np.array_equal(ad['valid_mask'], bd['valid_mask'])


@settings(max_examples=6, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10_000))
def test_field_structure(seed):
    sc = S.make_scene(3, 30, 40, seed=seed, feat=(5, 7, 6))
    pts = S.scattered_points(500, seed, sigma=0.3)
    full = O.field_eval(pts, sc.pose, sc.K, sc.depth, 30, 40, sc.maps, ['dino_feats'])
    rs = np.random.RandomState(seed)
    perm = rs.permutation(len(pts))
    p = O.field_eval(pts[perm], sc.pose, sc.K, sc.depth, 30, 40, sc.maps, ['dino_feats'])
    assert np.array_equal(p['dino_feats'], full['dino_feats'][perm]) and np.array_equal(p['dist'], full['dist'][perm])
    twice = O.field_eval(pts, sc.pose, sc.K, sc.depth, 30, 40, {'dino_feats': sc.maps['dino_feats'] * 2}, ['dino_feats'])
    # exact except where a far point's distance weight (exp(-(|d|-mu)/mu)) drives a product into the denormal range
    assert np.abs(twice['dino_feats'] - full['dino_feats'] * 2).max() <= 1e-35
    big = np.abs(full['dino_feats']) > 1e-30
    assert np.array_equal(twice['dino_feats'][big], (full['dino_feats'] * 2)[big])
