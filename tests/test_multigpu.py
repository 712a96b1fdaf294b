"""GPU, world size >= 2 over NCCL on real devices: the sharded field query through BOTH transports of
d3fields_b200.sharded — the peer-memory in-kernel gather (d3f_comm_* / d3f_eval_allgather) and the NCCL in-place
all-gather — must give every rank the full dist / valid_mask bit-identical to a single-rank evaluation of the same
points; the observation broadcast over peer memory must replicate rank 0's tensors.  Skipped on a one-GPU box
(run it with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import datetime
    import torch.distributed as dist
    from d3fields_b200 import Fusion, scene as S
    from d3fields_b200 import sharded as SH
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=180))
    report = {}
    try:
        sc = S.make_scene(4, 240, 320, seed=21, feat=(24, 32, 256), num_inst=4)
        f = Fusion(num_cam=4, device=str(dev))
        obs = {'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']}
        if rank != 0:                                   # only rank 0 "ran update()"
            obs = {k: np.zeros_like(v) for k, v in obs.items()}
        f.update(obs)
        f.set_instance_masks(torch.from_numpy(sc.maps['mask'] if rank == 0 else np.zeros_like(sc.maps['mask'])))

        nx, ny, nz = 8 * world, 30, 25
        pts_np = S.grid_points(nx, ny, nz)
        n = len(pts_np)
        comm = SH.make_peer_comm(n + 1000, device=dev, staging_bytes=1 << 20)   # small staging: several chunks per tensor
        report['peer_comm'] = comm is not None

        # --- observation replication: peer memory when available, else NCCL
        if comm is not None:
            SH.broadcast_observation_peer(f.curr_obs_torch, comm, src=0)
        else:
            SH.broadcast_observation(f.curr_obs_torch, src=0)
        torch.cuda.synchronize(dev)
        for k, ref in (('depth', sc.depth), ('pose', sc.pose), ('K', sc.K), ('dino_feats', sc.maps['dino_feats']),
                       ('mask', sc.maps['mask'])):
            assert np.array_equal(f.curr_obs_torch[k].cpu().numpy(), ref), f'broadcast of {k}'
        report['broadcast'] = True

        pts = torch.from_numpy(pts_np).to(dev)
        full = f.eval(pts, return_names=['dino_feats', 'mask'])             # single-rank truth on every rank

        def check(got, mine, what):
            assert got['dist'].shape == (n,) and got['valid_mask'].shape == (n,), what
            assert torch.equal(got['dist'], full['dist']), what + ': dist'
            assert torch.equal(got['valid_mask'], full['valid_mask']), what + ': valid_mask'
            assert torch.equal(got['dino_feats'], full['dino_feats'][mine]), what + ': local descriptors'

        for block in (None, ny * nz):
            mine = (torch.arange(*SH.shard_range(n, rank, world), device=dev) if block is None
                    else SH.block_interleaved_index(n, rank, world, block).to(dev))
            # NCCL transport, with a narrow key gathered too
            got = SH.eval_sharded(f.eval, pts, ['dino_feats', 'mask'], gather=('dist', 'valid_mask', 'mask'),
                                  channels={'mask': 4}, block=block)
            check(got, mine, f'nccl block={block}')
            assert torch.equal(got['mask'], full['mask'])
            # peer-memory transport: several steps in a row (epochs, double buffering), planned share reused
            if comm is not None:
                share = SH.plan_share(pts, block=block)
                prev = None
                for step in range(4):
                    got = SH.eval_sharded(f.eval, None, ['dino_feats'], comm=comm, share=share)
                    comm.check()
                    check(got, mine, f'peer block={block} step={step}')
                    if prev is not None:                 # the previous step's arrays are still intact (double buffer)
                        assert torch.equal(prev['dist'], full['dist'])
                    prev = got
                # ragged: a different n, ranks with different numbers of points
                sub = pts[:n - 37] if block is None else pts
                got = SH.eval_sharded(f.eval, sub, [], comm=comm, block=block)
                comm.check()
                assert torch.equal(got['dist'], full['dist'][:len(sub)]) and torch.equal(got['valid_mask'], full['valid_mask'][:len(sub)])
        # stress: 600 back-to-back steps alternating between two point sets, every step's gathered arrays compared on
        # the device.  A stale or late remote store (an ordering hole in the per-CTA gpu-scope fence / last-CTA
        # system fence protocol) would show up as a mismatch against the set of THAT step.
        if comm is not None:
            big = torch.from_numpy(S.grid_points(32 * world, 50, 50)).to(dev)       # thousands of CTAs per rank and step
            comm.close()
            comm = SH.make_peer_comm(len(big), device=dev, staging_bytes=0)
            assert comm is not None
            sets = [big, (big + torch.tensor([0.003, -0.002, 0.001], device=dev)).contiguous()]
            truth = [f.eval(s_, return_names=[]) for s_ in sets]
            shares = [SH.plan_share(s_, block=50 * 50) for s_ in sets]
            bad = torch.zeros((), dtype=torch.int64, device=dev)
            for step in range(600):
                k = step & 1
                got = SH.eval_sharded(f.eval, None, [], comm=comm, share=shares[k])
                bad += (got['dist'] != truth[k]['dist']).sum() + (got['valid_mask'] != truth[k]['valid_mask']).sum()
            comm.check()
            assert int(bad.item()) == 0, f'{int(bad.item())} stale / wrong gathered entries over 600 steps'
            assert not torch.equal(truth[0]['dist'], truth[1]['dist'])
            report['stress_steps'] = 600
        if comm is not None:
            comm.close()
        report['ok'] = True
    except Exception as e:                                # noqa: BLE001 - reported to the parent
        import traceback
        report['ok'] = False
        report['error'] = f'{type(e).__name__}: {e}\n{traceback.format_exc()}'
    finally:
        q.put((rank, report))
        try:
            dist.destroy_process_group()
        except Exception:
            pass


@pytest.mark.parametrize('world', [2])
def test_sharded_eval_on_real_ranks(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs, this box has {torch.cuda.device_count()}')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=420) for _ in range(world))
    for p in procs:
        p.join(60)
    for r in range(world):
        assert res[r].get('ok'), f"rank {r}: {res[r].get('error')}"
    print({r: {k: v for k, v in res[r].items() if k != 'error'} for r in res})
    # on a box whose GPUs are NVLink peers the peer-memory transport must have come up (no silent NCCL-only run)
    assert all(res[r]['peer_comm'] for r in range(world)), 'peer-memory communicator did not come up'
