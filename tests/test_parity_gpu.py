"""GPU: the CUDA path, called through the public Fusion API -> ctypes -> C ABI, against
  (1) the golden vectors the unmodified reference produced (tests/golden/, oracle/gen_golden.py),
  (2) the CPU oracle (oracle/field_oracle.py) on seeded inputs at sizes the oracle finishes in seconds,
  (3) size-independent properties at BASELINE.json's full size (1M points, V=4, C=1024).

Bars: dist and valid_mask bit-exact (they are decided by the integer pixel index and an ordered float32
sum); descriptors |a-b| <= 1e-4*|b| + 2e-6*max|b| (north_star: "fp32 descriptors within 1e-4 relative").
"""
import numpy as np
import pytest
import torch

from d3fields_b200 import _native, scene as S
from golden_util import CASES, Golden
from oracle import field_oracle as O
from util import assert_bits_equal, assert_close_field, make_fusion

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


@pytest.fixture(scope='module', autouse=True)
def _native_library_is_the_one_running():
    lib = _native.load()
    assert lib.d3f_abi_version() == _native.ABI_VERSION
    before = _native.launch_count()
    yield
    assert _native.launch_count() > before, 'no kernel of libd3f.so was launched by the GPU tests'


# ---------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize('name', CASES)
def test_matches_reference_golden(name):
    g = Golden(name)
    f = make_fusion(g.scene, DEV, mu=g.mu)
    pts = torch.from_numpy(g.pts).to(DEV)
    if g.meta['batch']:
        out = _np(f.batch_eval(pts, return_names=g.names))
    else:
        out = _np(f.eval(pts, return_names=g.names, return_inter=True))
    g.check_exact('dist', out['dist'])
    g.check_exact('valid_mask', out['valid_mask'])
    for k in g.names:
        g.check_close(k, out[k])
        if not g.meta['batch']:
            g.check_close(k + '_inter', out[k + '_inter'])
    od = _np(f.eval_dist(pts))
    g.check_exact('evaldist.dist', od['dist'])
    g.check_exact('evaldist.valid_mask', od['valid_mask'])


# ---------------------------------------------------------------------------- oracle, seeded cases
ORACLE_CASES = [
    # V, H, W, feat(h,w,C), num_inst, n_grid, n_scatter, seed
    (1, 60, 80, (6, 8, 4), 2, (9, 9, 9), 500, 10),
    (2, 240, 320, (24, 32, 64), 8, (20, 20, 10), 3000, 11),
    (4, 480, 640, (48, 64, 256), 8, (24, 24, 24), 4000, 12),
    (4, 480, 640, (48, 64, 1024), 0, (16, 16, 16), 2000, 13),        # cfg2a channels
    (6, 120, 160, (12, 16, 12), 5, (11, 13, 7), 1234, 14),
    (3, 97, 131, (97, 131, 7), 3, (10, 10, 10), 777, 15),            # full-resolution map, odd C
    (16, 48, 64, (5, 7, 3), 0, (8, 8, 8), 100, 16),                  # D3F_MAX_VIEWS
    (5, 120, 160, (12, 16, 128), 4, (14, 14, 9), 1500, 17),          # 5-8 views: the NV = 8 tile instantiation
    (8, 120, 160, (12, 16, 1024), 0, (12, 12, 12), 1500, 18),
    (7, 97, 131, (10, 13, 64), 3, (10, 10, 10), 900, 19),            # one 64-channel slice: tile split into runs
]


@pytest.mark.parametrize('case', ORACLE_CASES, ids=[f'V{c[0]}_C{c[3][2]}_{c[1]}x{c[2]}' for c in ORACLE_CASES])
@pytest.mark.parametrize('mask_u8', [False, True])
def test_matches_oracle(case, mask_u8):
    V, H, W, feat, num_inst, grid, n_sc, seed = case
    if mask_u8 and num_inst == 0:
        pytest.skip('no mask in this case')
    sc = S.make_scene(V, H, W, seed=seed, feat=feat, num_inst=num_inst, color=True)
    pts = np.concatenate([S.grid_points(*grid), S.scattered_points(n_sc, seed), S.adversarial_points(sc, seed, 8)])
    names = ['dino_feats', 'color_tensor'] + (['mask'] if num_inst else [])
    f = make_fusion(sc, DEV, mask_u8=mask_u8)
    got = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=names, return_inter=True))
    ref = O.field_eval(pts, sc.pose, sc.K, sc.depth, H, W, sc.maps, names, return_inter=True)
    assert got['valid_mask'].dtype == np.bool_ and got['dist'].dtype == np.float32
    assert_bits_equal(got['valid_mask'], ref['valid_mask'], 'valid_mask')
    assert_bits_equal(got['dist'], ref['dist'], 'dist')
    assert 0.05 < ref['valid_mask'].mean() < 0.95          # the case exercises both outcomes
    for k in names:
        assert_close_field(got[k], ref[k], what=k)
        assert_close_field(got[k + '_inter'], ref[k + '_inter'], what=k + '_inter')
        # points no view sees: descriptors exactly 0, dist exactly 1e3 (reference fusion.py:367, :386)
        none = ~ref['valid_mask']
        assert (got[k][none] == 0).all()
    assert (got['dist'][~ref['valid_mask']] == np.float32(1e3)).all()
    # the same query without per-view outputs takes the tile kernel for V <= 8 (the generic kernel above that)
    got2 = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=names))
    want = 'generic' if V > 8 else ('tile/wide' if feat[2] % (128 if V <= 4 else 64) == 0 else 'tile/narrow')
    assert _native.last_variant(0) == want, (_native.last_variant(0), want)
    assert_bits_equal(got2['valid_mask'], ref['valid_mask'], 'valid_mask (tile)')
    assert_bits_equal(got2['dist'], ref['dist'], 'dist (tile)')
    for k in names:
        assert_close_field(got2[k], ref[k], what=k + ' (tile)')
        assert (got2[k][~ref['valid_mask']] == 0).all()
    gd = _np(f.eval_dist(torch.from_numpy(pts).to(DEV)))
    rd = O.field_eval(pts, sc.pose, sc.K, sc.depth, H, W, eval_dist=True)
    assert_bits_equal(gd['dist'], rd['dist'], 'eval_dist.dist')
    assert_bits_equal(gd['valid_mask'], rd['valid_mask'], 'eval_dist.valid_mask')


@pytest.mark.parametrize('n', [0, 1, 2, 31, 63, 64, 65, 127, 129, 1000, 4097])
def test_ragged_sizes(n):
    sc = S.make_scene(4, 120, 160, seed=21, feat=(12, 16, 128), num_inst=4)
    pts = S.scattered_points(max(n, 1), 21, sigma=0.15)[:n]
    f = make_fusion(sc, DEV)
    got = _np(f.eval(torch.from_numpy(pts).to(DEV).reshape(n, 3), return_names=['dino_feats', 'mask']))
    ref = O.field_eval(pts.reshape(n, 3), sc.pose, sc.K, sc.depth, 120, 160, sc.maps, ['dino_feats', 'mask'])
    assert got['dist'].shape == (n,) and got['dino_feats'].shape == (n, 128) and got['mask'].shape == (n, 4)
    assert_bits_equal(got['dist'], ref['dist'], 'dist')
    assert_bits_equal(got['valid_mask'], ref['valid_mask'], 'valid_mask')
    assert_close_field(got['dino_feats'], ref['dino_feats'], what='dino_feats')
    assert_close_field(got['mask'], ref['mask'], what='mask')


def test_dist_only_and_default_names_keyerror():
    sc = S.make_scene(4, 120, 160, seed=22, feat=(12, 16, 8))
    pts = S.grid_points(20, 20, 20)
    f = make_fusion(sc, DEV)
    got = _np(f.batch_eval(torch.from_numpy(pts).to(DEV), return_names=[]))
    ref = O.field_eval(pts, sc.pose, sc.K, sc.depth, 120, 160)
    assert set(got) == {'dist', 'valid_mask'}
    assert_bits_equal(got['dist'], ref['dist'])
    assert_bits_equal(got['valid_mask'], ref['valid_mask'])
    with pytest.raises(KeyError):              # default return_names asks for 'mask' (reference fusion.py:305)
        f.eval(torch.from_numpy(pts).to(DEV))


def test_custom_mu():
    sc = S.make_scene(3, 120, 160, seed=23, feat=(12, 16, 16))
    pts = S.grid_points(16, 16, 16)
    for mu in (0.005, 0.02, 0.1):
        f = make_fusion(sc, DEV, mu=mu)
        got = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=['dino_feats']))
        ref = O.field_eval(pts, sc.pose, sc.K, sc.depth, 120, 160, sc.maps, ['dino_feats'], mu=mu)
        assert_bits_equal(got['dist'], ref['dist'], f'dist mu={mu}')
        assert_bits_equal(got['valid_mask'], ref['valid_mask'])
        assert_close_field(got['dino_feats'], ref['dino_feats'], what=f'feats mu={mu}')


def test_host_buffer_entry_point_equals_device_path():
    sc = S.make_scene(4, 240, 320, seed=24, feat=(24, 32, 256), num_inst=8)
    pts = np.concatenate([S.grid_points(40, 40, 40), S.scattered_points(7001, 24)])
    f = make_fusion(sc, DEV, mask_u8=True)
    dev = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=['dino_feats', 'mask']))
    host = f.eval(torch.from_numpy(pts), return_names=['dino_feats', 'mask'])
    assert all(not v.is_cuda for v in host.values())
    for k in dev:
        assert_bits_equal(host[k].numpy(), dev[k], k)
    out = {'dino_feats': torch.empty((len(pts), 256), dtype=torch.float32).pin_memory()}
    host2 = f.eval_host(torch.from_numpy(pts).pin_memory(), ['dino_feats'], out=out)
    assert host2['dino_feats'].data_ptr() == out['dino_feats'].data_ptr()
    assert_bits_equal(host2['dino_feats'].numpy(), dev['dino_feats'])
    assert_bits_equal(f.eval_host(torch.from_numpy(pts))['dist'].numpy(), dev['dist'])


def test_non_default_stream_and_noncontiguous_points():
    sc = S.make_scene(2, 120, 160, seed=25, feat=(12, 16, 32))
    pts = S.scattered_points(5000, 25)
    f = make_fusion(sc, DEV)
    base = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=['dino_feats']))
    wide = torch.zeros((5000, 6), device=DEV)
    wide[:, ::2] = torch.from_numpy(pts).to(DEV)
    s = torch.cuda.Stream(DEV)
    s.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(s):
        got = f.eval(wide[:, ::2], return_names=['dino_feats'])
    s.synchronize()
    for k in base:
        assert_bits_equal(got[k].cpu().numpy(), base[k], k)


def test_torch_cuda_rounding_flag_changes_only_boundary_pixels():
    """index_rounding='cuda' replays torch's CUDA kernels' rounding of the pixel normalisation; it may flip
    a handful of nearest/floor decisions that sit within float rounding of a pixel border, nothing else."""
    sc = S.make_scene(4, 480, 640, seed=26, feat=(48, 64, 16))
    pts = S.grid_points(40, 40, 40)
    f = make_fusion(sc, DEV)
    a = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=['dino_feats']))
    f.index_rounding = 'cuda'
    b = _np(f.eval(torch.from_numpy(pts).to(DEV), return_names=['dino_feats']))
    assert (a['valid_mask'] != b['valid_mask']).mean() < 1e-3
    same = a['dist'] == b['dist']
    assert same.mean() > 0.99


# ---------------------------------------------------------------------------- full size: properties
@pytest.fixture(scope='module')
def cfg2a():
    c = S.CONFIGS['cfg2a']
    sc = S.make_scene(c['V'], c['H'], c['W'], seed=0, feat=c['feat'])
    pts = S.config_points('cfg2a')
    f = make_fusion(sc, DEV)
    p = torch.from_numpy(pts).to(DEV)
    out = f.batch_eval(p, return_names=['dino_feats'])
    torch.cuda.synchronize()
    return sc, pts, f, p, out


def test_full_size_sampled_rows_match_oracle(cfg2a):
    sc, pts, f, p, out = cfg2a
    rs = np.random.RandomState(5)
    rows = np.sort(rs.choice(len(pts), 3000, replace=False))
    ref = O.field_eval(pts[rows], sc.pose, sc.K, sc.depth, sc.H, sc.W, sc.maps, ['dino_feats'])
    idx = torch.from_numpy(rows).to(DEV)
    assert_bits_equal(out['dist'][idx].cpu().numpy(), ref['dist'], 'dist')
    assert_bits_equal(out['valid_mask'][idx].cpu().numpy(), ref['valid_mask'], 'valid_mask')
    assert_close_field(out['dino_feats'][idx].cpu().numpy(), ref['dino_feats'], what='dino_feats')
    assert 0.2 < ref['valid_mask'].mean() < 0.9


def test_full_size_dist_valid_match_oracle_everywhere(cfg2a):
    sc, pts, f, p, out = cfg2a
    ref = O.field_eval(pts, sc.pose, sc.K, sc.depth, sc.H, sc.W, chunk=1 << 17)
    assert_bits_equal(out['dist'].cpu().numpy(), ref['dist'], 'dist (1M)')
    assert_bits_equal(out['valid_mask'].cpu().numpy(), ref['valid_mask'], 'valid_mask (1M)')


def test_full_size_permutation_and_slicing_invariance(cfg2a):
    """Every point is independent (reference fusion.py:305-394 has no cross-point term): evaluating a
    permutation or a slice must give bit-identical rows, whatever tile/cache state they land in."""
    sc, pts, f, p, out = cfg2a
    g = torch.Generator(device='cpu').manual_seed(3)
    perm = torch.randperm(len(pts), generator=g)[:200_000].to(DEV)
    o2 = f.eval(p[perm].contiguous(), return_names=['dino_feats'])
    assert torch.equal(o2['dist'], out['dist'][perm])
    assert torch.equal(o2['valid_mask'], out['valid_mask'][perm])
    assert torch.equal(o2['dino_feats'], out['dino_feats'][perm])
    a, b = 123_457, 323_456
    o3 = f.eval(p[a:b], return_names=['dino_feats'])
    assert torch.equal(o3['dino_feats'], out['dino_feats'][a:b])
    assert torch.equal(o3['dist'], out['dist'][a:b])


def test_full_size_linearity_in_the_feature_volume(cfg2a):
    """The field is linear in the sampled map: F(2*A) == 2*F(A) exactly (power-of-two scaling is exact in
    float32), and F(const) has every valid row equal to const * sum of its weights."""
    sc, pts, f, p, out = cfg2a
    vol = f.curr_obs_torch['dino_feats']
    f.curr_obs_torch['dino_feats'] = (vol * 2).contiguous()
    o2 = f.eval(p[:300_000], return_names=['dino_feats'])
    assert torch.equal(o2['dino_feats'], out['dino_feats'][:300_000] * 2)
    f.curr_obs_torch['dino_feats'] = torch.ones_like(vol)
    o1 = f.eval(p[:300_000], return_names=['dino_feats'])['dino_feats']
    spread = (o1.max(1).values - o1.min(1).values)
    assert float(spread.max()) == 0.0                      # all channels of a row see the same weights
    assert float(o1.max()) <= 1.0 + 1e-5 and float(o1.min()) >= 0.0
    f.curr_obs_torch['dino_feats'] = vol


# ---------------------------------------------------------------------------- neighbours of the path
def test_pca_projection_matches_numpy():
    rs = np.random.RandomState(0)
    for n, c, k in ((1000, 1024, 3), (257, 64, 8), (33, 7, 2)):
        x = rs.standard_normal((n, c)).astype(np.float32)
        mean = rs.standard_normal(c).astype(np.float32)
        comp = rs.standard_normal((k, c)).astype(np.float32)
        f = make_fusion(S.make_scene(1, 8, 8, seed=0), DEV)
        y = f.pca_project(torch.from_numpy(x).to(DEV), mean, comp).cpu().numpy()
        ref = (x.astype(np.float64) - mean) @ comp.T.astype(np.float64)
        assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max()


def test_device_grid_matches_create_init_grid():
    from d3fields_b200 import create_init_grid, create_init_grid_device
    b = S.WORKSPACE
    for step in (0.01, 0.004, 0.0173):
        ref, shape = create_init_grid(b, step)
        pts, shape_d = create_init_grid_device(b, step, DEV)
        assert tuple(shape_d) == tuple(shape) and pts.shape == ref.shape
        d = (pts.cpu() - ref).abs().max().item()
        assert d <= 6e-8, d        # torch.arange's vectorised path may differ from the scalar formula by 1 ulp


def test_sharded_helper_single_rank_writes_in_place():
    """world_size 1 path of d3fields_b200.sharded.eval_sharded over the CUDA kernels: same results as eval, and
    the gathered keys are the buffers the kernel wrote (no copy before the collective)."""
    from d3fields_b200.sharded import eval_sharded
    sc = S.make_scene(4, 120, 160, seed=31, feat=(12, 16, 128), num_inst=4)
    pts = torch.from_numpy(S.grid_points(25, 25, 25)).to(DEV)
    f = make_fusion(sc, DEV)
    seen = {}

    def eval_fn(local, return_names, return_inter=False, out=None):
        r = f.eval(local, return_names=return_names, return_inter=return_inter, out=out)
        seen.update({k: (r[k].data_ptr(), out[k].data_ptr()) for k in out})
        return r

    got = eval_sharded(eval_fn, pts, ['dino_feats', 'mask'], gather=('dist', 'valid_mask', 'mask'), channels={'mask': 4})
    ref = f.eval(pts, return_names=['dino_feats', 'mask'])
    assert got['shard'] == (0, len(pts))
    assert all(a == b for a, b in seen.values())
    for k in ('dist', 'valid_mask', 'mask', 'dino_feats'):
        assert torch.equal(got[k], ref[k]), k
    assert not any(k.endswith('_inter') for k in got)
    # the bound Fusion.eval passed directly (what the docs show): still the tile kernel, still in place
    got2 = eval_sharded(f.eval, pts, ['dino_feats'], gather=('dist', 'valid_mask'))
    assert _native.last_variant(0) == 'tile/wide'
    assert torch.equal(got2['dino_feats'], ref['dino_feats']) and torch.equal(got2['dist'], ref['dist'])
    assert not any(k.endswith('_inter') for k in got2)


def test_peer_comm_single_rank_gather_matches_eval():
    """d3f_eval_allgather at world size 1 (the in-kernel gather path with its epoch flags, no peers): contiguous
    and block-interleaved index maps land every point at its canonical position, twice in a row (double buffer)."""
    from d3fields_b200.sharded import PeerComm, eval_sharded, plan_share
    sc = S.make_scene(4, 120, 160, seed=33, feat=(12, 16, 128), num_inst=4)
    pts = torch.from_numpy(np.concatenate([S.grid_points(20, 20, 20), S.scattered_points(777, 5)])).to(DEV)
    f = make_fusion(sc, DEV)
    ref = f.eval(pts, return_names=['dino_feats'])
    comm = PeerComm(len(pts), device=DEV, staging_bytes=1 << 20)
    try:
        for rep in range(3):
            got = eval_sharded(f.eval, pts, ['dino_feats'], comm=comm)
            comm.check()
            assert got['shard'] == (0, len(pts))
            for k in ('dist', 'valid_mask', 'dino_feats'):
                assert torch.equal(got[k], ref[k]), (rep, k)
        share = plan_share(pts[:8000], block=400)
        got = eval_sharded(f.eval, None, [], comm=comm, share=share)
        comm.check()
        assert torch.equal(got['dist'], ref['dist'][:8000]) and torch.equal(got['valid_mask'], ref['valid_mask'][:8000])
        got = eval_sharded(f.eval, pts[:0], [], comm=comm)          # a rank without points still joins the exchange
        comm.check()
        assert got['dist'].numel() == 0
        t = torch.arange(1000, device=DEV, dtype=torch.float32)
        comm.broadcast(t, root=0)                                    # world 1: a no-op
        assert t[999].item() == 999
    finally:
        comm.close()


@pytest.mark.parametrize('C,k', [(1024, 3), (256, 4), (128, 1), (64, 8), (7, 2)])
def test_pca_field_via_projected_volume(C, k):
    """Fusion.eval_pca == eval(..)[name] followed by sklearn-style (x - mean) @ components.T (reference
    fusion.py:1386-1392).  The kernel samples the PROJECTED volume (linearity of the field in the map)."""
    sc = S.make_scene(4, 240, 320, seed=41, feat=(24, 32, C))
    pts_np = np.concatenate([S.grid_points(30, 30, 30), S.scattered_points(5000, 41), S.adversarial_points(sc, 41, 8)])
    pts = torch.from_numpy(pts_np).to(DEV)
    rs = np.random.RandomState(7)
    mean = rs.standard_normal(C).astype(np.float32)
    comp = (rs.standard_normal((k, C)) / np.sqrt(C)).astype(np.float32)
    f = make_fusion(sc, DEV)
    got = f.eval_pca(pts, 'dino_feats', mean, comp)
    assert _native.last_variant(0) == 'tile/narrow'
    ref = f.eval(pts, return_names=['dino_feats'])
    assert torch.equal(got['dist'], ref['dist']) and torch.equal(got['valid_mask'], ref['valid_mask'])
    x = ref['dino_feats'].cpu().numpy().astype(np.float64)
    y = (x - mean.astype(np.float64)) @ comp.T.astype(np.float64)
    g = got['dino_feats_pca'].cpu().numpy()
    assert g.shape == (len(pts_np), k)
    assert np.abs(g - y).max() <= 1e-4 * np.abs(y).max()
    none = ~ref['valid_mask'].cpu().numpy()
    base = -(mean.astype(np.float64) @ comp.T.astype(np.float64))
    assert np.abs(g[none] - base).max() <= 1e-5 * max(1.0, np.abs(base).max())
    # plain projection kernel without centring
    z = f.pca_project(ref['dino_feats'], None, comp).cpu().numpy()
    assert np.abs(z - x @ comp.T.astype(np.float64)).max() <= 1e-4 * max(1e-6, np.abs(x @ comp.T).max())


def test_backward_matches_torch_autograd_through_the_reference_operator_sequence():
    """d loss / d pts through Fusion.eval (d3f_eval_backward) against torch autograd through oracle/torch_port.py —
    the reference's own operator sequence, which is what its rigid_tracking differentiates (fusion.py:1650-1665).
    Points sit within a few mu of the surfaces so every differentiable term is active: bilinear coordinates,
    the distance weight (|d| > mu), the clamped dist (|d| < mu)."""
    from oracle import torch_port as TP
    sc = S.make_scene(4, 240, 320, seed=51, feat=(24, 32, 64), num_inst=4)
    rs = np.random.RandomState(51)
    base = S.grid_points(60, 60, 40)
    ref0 = O.field_eval(base, sc.pose, sc.K, sc.depth, sc.H, sc.W)
    near = base[np.abs(ref0['dist']) < 0.0199][::7][:1500]           # inside the truncation band
    pts_np = np.concatenate([near, near + rs.normal(0, 0.03, near.shape).astype(np.float32), S.scattered_points(500, 51)])
    n = len(pts_np)
    Gf = rs.standard_normal((n, 64)).astype(np.float32)
    Gm = rs.standard_normal((n, 4)).astype(np.float32)
    gd = rs.standard_normal(n).astype(np.float32)
    # reference operators, CPU float32 autograd
    obs = TP.obs_from_scene(sc)
    p_ref = torch.from_numpy(pts_np).clone().requires_grad_(True)
    o = TP.eval_chunk(obs, sc.H, sc.W, p_ref, ['dino_feats', 'mask'])
    loss = (o['dino_feats'] * torch.from_numpy(Gf)).sum() + (o['mask'] * torch.from_numpy(Gm)).sum() + (o['dist'] * torch.from_numpy(gd)).sum()
    loss.backward()
    g_ref = p_ref.grad.numpy()
    # native
    f = make_fusion(sc, DEV)
    p = torch.from_numpy(pts_np).to(DEV).requires_grad_(True)
    out = f.eval(p, return_names=['dino_feats', 'mask'])
    assert out['dino_feats'].requires_grad and out['dist'].requires_grad and not out['valid_mask'].requires_grad
    loss2 = (out['dino_feats'] * torch.from_numpy(Gf).to(DEV)).sum() + (out['mask'] * torch.from_numpy(Gm).to(DEV)).sum() \
        + (out['dist'] * torch.from_numpy(gd).to(DEV)).sum()
    loss2.backward()
    g = p.grad.cpu().numpy()
    assert np.isfinite(g).all()
    scale = np.abs(g_ref).max()
    assert scale > 1.0
    err = np.abs(g - g_ref)
    assert (err <= 2e-3 * np.abs(g_ref) + 2e-4 * scale).all(), f'max err {err.max():.3e} at scale {scale:.3e}'
    assert (np.abs(g_ref).max(1) > 0).mean() > 0.3           # the case has plenty of points with gradient
    # points no view sees get exactly zero
    none = ~out['valid_mask'].cpu().numpy()
    assert none.any() and (g[none] == 0).all() and (g_ref[none] == 0).all()
    # dist-only gradient (return_names=[])
    p2 = torch.from_numpy(pts_np).to(DEV).requires_grad_(True)
    (f.eval(p2, return_names=[])['dist'] * torch.from_numpy(gd).to(DEV)).sum().backward()
    p3 = torch.from_numpy(pts_np).clone().requires_grad_(True)
    (TP.eval_chunk(obs, sc.H, sc.W, p3, [])['dist'] * torch.from_numpy(gd)).sum().backward()
    e2 = np.abs(p2.grad.cpu().numpy() - p3.grad.numpy())
    assert (e2 <= 2e-3 * np.abs(p3.grad.numpy()) + 2e-4 * np.abs(p3.grad.numpy()).max()).all()


def test_cuda_rounding_mode_against_the_reference_operators_run_by_torch_on_the_gpu():
    """The parity oracle is the reference's CPU path.  A user who runs the reference on a GPU gets torch's CUDA
    kernels' rounding of the pixel normalisation instead (p * (1/(W-1)), ((c+1)/2)*(size-1)) and cuBLAS's order in
    the projection; index_rounding='cuda' replays the first two.  This measures how close that gets to the
    reference operator sequence actually executed by torch on this GPU (oracle/torch_port.py on device='cuda'):
    a handful of boundary pixels may still differ through the projection's summation order."""
    from oracle import torch_port as TP
    sc = S.make_scene(4, 480, 640, seed=61, feat=(48, 64, 32), num_inst=4)
    pts_np = np.concatenate([S.grid_points(60, 60, 60), S.scattered_points(50000, 61)])
    obs = {k: v.to(DEV) for k, v in TP.obs_from_scene(sc).items()}
    pts = torch.from_numpy(pts_np).to(DEV)
    ref = TP.batch_eval(obs, sc.H, sc.W, pts, ['dino_feats', 'mask'])
    f = make_fusion(sc, DEV)
    res = {}
    for mode in ('cpu', 'cuda'):
        f.index_rounding = mode
        out = f.eval(pts, return_names=['dino_feats', 'mask'])
        vm = (out['valid_mask'] != ref['valid_mask']).float().mean().item()
        same = ref['valid_mask'] & out['valid_mask']
        dd = (out['dist'][same] - ref['dist'][same]).abs()
        feat_err = (out['dino_feats'][same] - ref['dino_feats'][same]).abs().max().item()
        res[mode] = (vm, (dd > 1e-6).float().mean().item(), feat_err)
    print('mismatch vs torch-CUDA reference ops  (valid_mask frac, dist>1e-6 frac, max feat err):', res)
    assert res['cuda'][0] <= 2e-4 and res['cpu'][0] <= 2e-3
    assert res['cuda'][1] <= 2e-3


def test_pose_recovery_by_gradient_descent_through_eval():
    """The use the backward exists for (reference rigid_tracking, fusion.py:1608-1685): points on the sphere are moved
    by a small unknown translation; Adam on the translation, through Fusion.eval's descriptors and dist, brings them
    back.  A smooth low-frequency feature volume makes the descriptor loss informative."""
    V, H, W = 4, 240, 320
    sc = S.make_scene(V, H, W, seed=71, hole_frac=0.0)
    hh, ww, C = 24, 32, 128
    yy, xx = np.meshgrid(np.linspace(0, 1, hh), np.linspace(0, 1, ww), indexing='ij')
    rs = np.random.RandomState(71)
    fr = rs.uniform(0.5, 2.0, size=(C, 2)); ph = rs.uniform(0, 6.28, size=C)
    vol = np.sin(2 * np.pi * (fr[:, 0] * xx[..., None] + fr[:, 1] * yy[..., None]) + ph).astype(np.float32)   # (h,w,C)
    sc.maps['dino_feats'] = np.ascontiguousarray(np.broadcast_to(vol, (V, hh, ww, C))).copy()
    f = make_fusion(sc, DEV)
    # surface points: upper hemisphere of the r=0.25 sphere
    u = rs.uniform(0, 2 * np.pi, 400); cz = rs.uniform(0.3, 0.95, 400)
    src = (0.25 * np.stack([np.sqrt(1 - cz ** 2) * np.cos(u), np.sqrt(1 - cz ** 2) * np.sin(u), cz], -1)).astype(np.float32)
    src_t = torch.from_numpy(src).to(DEV)
    with torch.no_grad():
        target = f.eval(src_t, return_names=['dino_feats'])
    keep = target['valid_mask']
    assert keep.float().mean() > 0.8
    src_t, tfeat = src_t[keep], target['dino_feats'][keep]
    true_shift = torch.tensor([0.012, -0.009, 0.006], device=DEV)
    moved = src_t + true_shift
    t = torch.zeros(3, device=DEV, requires_grad=True)
    opt = torch.optim.Adam([t], lr=2e-3)
    first = None
    for it in range(150):
        opt.zero_grad()
        out = f.eval(moved - t, return_names=['dino_feats'])
        w = out['valid_mask'].float().unsqueeze(-1)
        loss = (((out['dino_feats'] - tfeat) ** 2) * w).sum() / w.sum() / C + 10.0 * (out['dist'].clamp(-0.02, 0.02) ** 2).mean()
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
    err0 = true_shift.norm().item()
    err = (t.detach() - true_shift).norm().item()
    assert loss.item() < 0.2 * first and err < 0.35 * err0, (first, loss.item(), err0, err)
