"""Sharding the field query across the GPUs of one box (one process per GPU, torch.distributed).

Every query point is independent (reference fusion.py:305-394 has no cross-point term), so the
path shards with no data-path collective: rank r evaluates the contiguous slab
[shard_range(N, r, world)) of the point array — for a create_init_grid grid that is a slab along x,
which keeps the z-fastest locality the kernel's texel cache relies on.  The observation
(pose, K, depth and the sampled maps, 55 MB in the reference layout) is replicated once per
update() with broadcast_observation().

What IS exchanged is only the compact per-point result a caller needs everywhere — dist (4 B/pt),
valid_mask (1 B/pt), optionally a narrow key such as the instance mask — with one in-place
all_gather per tensor over NCCL/NVLink: the kernel writes each rank's slab straight into its slot
of the gather buffer, so no copy precedes the collective.  The 1024-channel descriptor field is
left sharded: gathering it would move 4 KB/pt to every GPU (57 GB at 16 M points, ~64 ms on
NVLink 5 against ~1.6 ms of kernel time per GPU) and no caller of the reference needs it on every
device (dense grids are evaluated with return_names=[] or ['mask'], reference vis_repr.py:93,
fusion.py:1428; descriptors only on mesh vertices / keypoints).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab of rank `rank`: sizes differ by at most one point, slabs tile [0, n)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def slab_capacity(n: int, world: int) -> int:
    """Slot size of the in-place gather buffer: the largest slab."""
    return max(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)) if world > 0 else n


def broadcast_observation(obs: Dict[str, object], src: int = 0, group=None) -> None:
    """Replicate every tensor of Fusion.curr_obs_torch from rank `src` (same shapes on all ranks)."""
    for k in sorted(obs):
        v = obs[k]
        if isinstance(v, torch.Tensor):
            dist.broadcast(v, src=src, group=group)


def block_interleaved_index(n: int, rank: int, world: int, block: int) -> torch.Tensor:
    """Indices of rank `rank` when blocks of `block` consecutive points (e.g. one x-plane of a grid, ny*nz points) are
    dealt round-robin to the ranks.  Use it when the scene is spatially heterogeneous: contiguous slabs then differ in
    work (the cost of a point depends on how many views see it) and the step is the max over ranks; interleaved shares
    are statistically identical.  Requires n % (block*world) == 0."""
    if block <= 0 or n % (block * world) != 0:
        raise ValueError(f'n={n} must be a multiple of block*world={block * world}')
    blocks = torch.arange(rank, n // block, world)
    return (blocks[:, None] * block + torch.arange(block)[None, :]).reshape(-1)


def deinterleave(gathered: torch.Tensor, world: int, block: int) -> torch.Tensor:
    """Rank-major all-gather result of block-interleaved shares -> canonical point order."""
    n = gathered.shape[0]
    per = n // world
    g = gathered.reshape(world, per // block, block, *gathered.shape[1:])
    return g.transpose(0, 1).reshape(n, *gathered.shape[1:])


def eval_sharded(eval_fn: Callable[..., Dict[str, torch.Tensor]], pts: torch.Tensor,
                 return_names: Iterable[str] = (), gather: Sequence[str] = ('dist', 'valid_mask'),
                 channels: Optional[Dict[str, int]] = None, group=None, block: Optional[int] = None) -> Dict[str, object]:
    """Evaluate rank-local shares of `pts` (the full (N,3) array, identical on every rank) and all-gather
    the keys listed in `gather`.

    block=None   contiguous slabs (shard_range).
    block=B      blocks of B consecutive points (e.g. one x-plane of a grid: ny*nz) dealt round-robin: every rank gets
                 the same spatial mix, which matters because a point's cost depends on how many views see it (measured:
                 contiguous slabs of the benchmark workspace cost 0.66-0.92 ms, interleaved shares 0.78 ms each).
                 Needs N % (B*world) == 0.  Gathered keys come back in canonical point order; keys left sharded come
                 back as the rank's share together with 'index' (their positions in the full array).

    eval_fn(local_pts, return_names, out) -> dict — e.g. Fusion.eval: must write into the tensors of `out`
    when given (so the kernel fills the gather slot in place) and return tensors on pts' device.
    channels: C of each gathered key other than dist/valid_mask (needed to size its gather buffer).

    Returns {'shard': (start, end), gathered keys -> full (N, ...) tensors, other keys -> the local slab}.
    """
    names = list(return_names)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(pts.shape[0])
    if block is not None:
        return _eval_interleaved(eval_fn, pts, names, gather, channels, group, block, world, rank)
    start, end = shard_range(n, rank, world)
    cap = slab_capacity(n, world)
    dev = pts.device
    bufs: Dict[str, torch.Tensor] = {}
    out: Dict[str, torch.Tensor] = {}
    for k in gather:
        if k == 'dist':
            shape, dt = (world, cap), torch.float32
        elif k == 'valid_mask':
            shape, dt = (world, cap), torch.bool
        else:
            if k not in names:
                raise KeyError(f'gather key {k!r} is not in return_names')
            if not channels or k not in channels:
                raise ValueError(f'channels[{k!r}] is needed to size the gather buffer')
            shape, dt = (world, cap, int(channels[k])), torch.float32
        bufs[k] = torch.empty(shape, dtype=dt, device=dev)
        out[k] = bufs[k][rank, :end - start]                 # this rank's slot: the kernel writes here
    res = eval_fn(pts[start:end], names, out)
    result: Dict[str, object] = {'shard': (start, end)}
    for k, v in res.items():
        if k not in bufs:
            result[k] = v
    for k, b in bufs.items():
        if res[k].data_ptr() != out[k].data_ptr():           # eval_fn ignored `out`: copy into the slot
            out[k].copy_(res[k])
        if world > 1:
            flat = b.view(torch.uint8) if b.dtype == torch.bool else b
            # in place: the input is this rank's slot of the output buffer
            dist.all_gather_into_tensor(flat.view(world * cap, *flat.shape[2:]), flat[rank], group=group)
        if all(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] == cap for r in range(world)):
            result[k] = b.reshape(world * cap, *b.shape[2:])[:n]
        else:                                                # ragged slabs: drop each slot's padding
            result[k] = torch.cat([b[r, :shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)], 0)
    return result


def _eval_interleaved(eval_fn, pts, names, gather, channels, group, block, world, rank):
    n = int(pts.shape[0])
    dev = pts.device
    idx = block_interleaved_index(n, rank, world, block).to(dev)
    per = n // world
    local = pts.index_select(0, idx)
    bufs: Dict[str, torch.Tensor] = {}
    out: Dict[str, torch.Tensor] = {}
    for k in gather:
        if k == 'dist':
            shape, dt = (world, per), torch.float32
        elif k == 'valid_mask':
            shape, dt = (world, per), torch.bool
        else:
            if k not in names:
                raise KeyError(f'gather key {k!r} is not in return_names')
            if not channels or k not in channels:
                raise ValueError(f'channels[{k!r}] is needed to size the gather buffer')
            shape, dt = (world, per, int(channels[k])), torch.float32
        bufs[k] = torch.empty(shape, dtype=dt, device=dev)
        out[k] = bufs[k][rank]
    res = eval_fn(local, names, out)
    result: Dict[str, object] = {'index': idx}
    for k, v in res.items():
        if k not in bufs:
            result[k] = v
    for k, b in bufs.items():
        if res[k].data_ptr() != out[k].data_ptr():
            out[k].copy_(res[k])
        if world > 1:
            flat = b.view(torch.uint8) if b.dtype == torch.bool else b
            dist.all_gather_into_tensor(flat.view(world * per, *flat.shape[2:]), flat[rank], group=group)
        result[k] = deinterleave(b.reshape(world * per, *b.shape[2:]), world, block)
    return result
