"""CPU, world_size 2 and 3 over gloo: the N>1 host logic — slab partition, in-place gather layout, ragged
slabs, replication of the observation.  The per-slab evaluation is the CPU oracle here (tests may call it);
on the GPU box tests/test_parity_gpu.py::test_sharded_* runs the same helper over the CUDA kernels at world size 1
and tests/test_multigpu.py spawns real NCCL ranks (2 GPUs) for both transports."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d3fields_b200 import scene as S
from d3fields_b200.sharded import (block_interleaved_index, broadcast_observation, deinterleave, eval_sharded,
                                   shard_range, slab_capacity)


def test_shard_ranges_tile_the_points():
    for n in (0, 1, 7, 128, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1 and slab_capacity(n, world) == max(sizes)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q, block=None):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import field_oracle as O
        sc = S.make_scene(2, 60, 80, seed=5, feat=(6, 8, 8), num_inst=3)
        obs = {'pose': torch.from_numpy(sc.pose), 'K': torch.from_numpy(sc.K), 'depth': torch.from_numpy(sc.depth),
               'mask': torch.from_numpy(sc.maps['mask']), 'dino_feats': torch.from_numpy(sc.maps['dino_feats'])}
        if rank != 0:                      # only rank 0 "ran update()": the others receive the observation
            for v in obs.values():
                v.zero_()
        broadcast_observation(obs, src=0)
        assert torch.equal(obs['depth'], torch.from_numpy(sc.depth))
        pts = torch.from_numpy(S.scattered_points(n, 9))
        maps = {'mask': obs['mask'].numpy(), 'dino_feats': obs['dino_feats'].numpy()}

        def eval_fn(local, return_names, return_inter=False, out=None):          # Fusion.eval's signature
            assert return_inter is False and isinstance(out, dict)
            r = O.field_eval(local.numpy(), obs['pose'].numpy(), obs['K'].numpy(), obs['depth'].numpy(), 60, 80, maps,
                             return_names)
            res = {k: torch.from_numpy(v) for k, v in r.items()}
            if rank == 0:                  # one rank honours `out` (in-place slot), the other returns fresh tensors
                for k in out:
                    out[k].copy_(res[k]); res[k] = out[k]
            return res

        got = eval_sharded(eval_fn, pts, ['dino_feats', 'mask'], gather=('dist', 'valid_mask', 'mask'), channels={'mask': 3},
                           block=block)
        full = O.field_eval(pts.numpy(), sc.pose, sc.K, sc.depth, 60, 80, sc.maps, ['dino_feats', 'mask'])
        if block is None:
            s, e = got['shard']
            assert (s, e) == shard_range(n, rank, world)
            mine = np.arange(s, e)
        else:
            mine = block_interleaved_index(n, rank, world, block).numpy()
            assert np.array_equal(got['index'].numpy(), mine)
        ok = (np.array_equal(got['dist'].numpy(), full['dist']) and np.array_equal(got['valid_mask'].numpy(), full['valid_mask'])
              and np.array_equal(got['mask'].numpy(), full['mask']) and np.array_equal(got['dino_feats'].numpy(), full['dino_feats'][mine])
              and got['dist'].shape == (n,) and got['mask'].shape == (n, 3))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,block', [(2, 1000, None), (2, 1001, None), (3, 500, None), (2, 1200, 50), (3, 900, 25)])
def test_sharded_eval_over_gloo(world, n, block):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q, block)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f'worker exit code {p.exitcode}'
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: True for r in range(world)}


def test_block_interleaved_shares_and_deinterleave():
    n, world, block = 2 * 3 * 5 * 4, 3, 5
    idx = [block_interleaved_index(n, r, world, block) for r in range(world)]
    allidx = torch.cat(idx)
    assert sorted(allidx.tolist()) == list(range(n))                 # a partition
    assert idx[1][:block].tolist() == list(range(block, 2 * block))  # rank 1 starts with block 1
    field = torch.arange(n, dtype=torch.float32) * 10
    gathered = torch.cat([field[i] for i in idx])                    # what a rank-major all_gather returns
    assert torch.equal(deinterleave(gathered, world, block), field)
    f2 = torch.stack([field, -field], 1)
    assert torch.equal(deinterleave(torch.cat([f2[i] for i in idx]), world, block), f2)
    with pytest.raises(ValueError):
        block_interleaved_index(n + 1, 0, world, block)


def test_share_plan_gather_index_formula_is_the_canonical_index():
    """The in-kernel gather places local point i at base + (i // block) * stride + i % block (include/d3f.h,
    d3f_eval_allgather).  For both layouts plan_share produces, that is the point's position in the full array."""
    from d3fields_b200.sharded import Share, plan_share
    import torch.distributed as dist
    assert not dist.is_initialized()
    pts = torch.arange(3 * 2 * 5 * 4 * 3, dtype=torch.float32).reshape(-1, 3)
    n = pts.shape[0]
    # world 1 through plan_share itself
    for block in (None, 5):
        sh = plan_share(pts, block=block)
        i = torch.arange(sh.local.shape[0])
        gi = sh.base + (i // sh.block) * sh.stride + i % sh.block
        assert torch.equal(pts[gi], sh.local) and sh.n == n
    # any world: the same formula with the ranks' parameters
    for world in (2, 3, 4):
        seen = torch.zeros(n, dtype=torch.int32)
        for rank in range(world):
            # contiguous slab
            s, e = shard_range(n, rank, world)
            i = torch.arange(e - s)
            base, blk, stride = s, max(e - s, 1), 0
            assert torch.equal(base + (i // blk) * stride + i % blk, torch.arange(s, e))
        block = 5
        if n % (block * world) == 0:
            for rank in range(world):
                idx = block_interleaved_index(n, rank, world, block)
                i = torch.arange(idx.numel())
                gi = rank * block + (i // block) * (world * block) + i % block
                assert torch.equal(gi, idx)
                seen[gi] += 1
            assert (seen == 1).all()
