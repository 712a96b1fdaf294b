"""Sharding the field query across the GPUs of one box (one process per GPU, torch.distributed).

Every query point is independent (reference fusion.py:305-394 has no cross-point term), so the
path shards with no data-path collective: rank r evaluates the contiguous slab
[shard_range(N, r, world)) of the point array — for a create_init_grid grid that is a slab along x,
which keeps the z-fastest locality the kernel's texel cache relies on.  The observation
(pose, K, depth and the sampled maps, 55 MB in the reference layout) is replicated once per
update() with broadcast_observation().

What IS exchanged is only the compact per-point result a caller needs everywhere — dist (4 B/pt),
valid_mask (1 B/pt), optionally a narrow key such as the instance mask.  Two transports:

  * PeerComm (d3f_comm_* / d3f_eval_allgather in include/d3f.h): every rank's gathered arrays live in a
    CUDA-IPC-mapped segment; the field kernel stores dist / valid_mask of each point straight into the arrays
    of ALL ranks (remote st.global over NVLink, 5 B/point/peer) and its last CTA exchanges epoch flags with the
    peers — one launch per step, no NCCL kernel, no side stream, results in canonical point order.
  * NCCL (any torch.distributed group; also gloo on CPU): the kernel writes the rank's slab into its slot of
    the gather buffer in place and one all_gather_into_tensor per key follows.  Used for narrow keys, and as
    the fallback when peer mapping is unavailable.

The 1024-channel descriptor field is
left sharded: gathering it would move 4 KB/pt to every GPU (57 GB at 16 M points, ~64 ms on
NVLink 5 against ~1.6 ms of kernel time per GPU) and no caller of the reference needs it on every
device (dense grids are evaluated with return_names=[] or ['mask'], reference vis_repr.py:93,
fusion.py:1428; descriptors only on mesh vertices / keypoints).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _native


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin the calling thread (and the threads it starts later) to the CPU cores next to GPU `device_index`, so the
    pinned host buffers it allocates afterwards are first-touched on that GPU's NUMA node.  With one rank per GPU and
    4 GB of results per step crossing PCIe, ranks that share one node's memory controller and root complex do not
    scale (round 1: 13 -> 57 Mpts/s from 1 to 8 GPUs).  Returns a description of what was done."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:                                                # CUDA_VISIBLE_DEVICES may renumber: go through the UUID
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
        except Exception:                                   # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 'gpu-local cpu set is empty under this cgroup: affinity unchanged'
        os.sched_setaffinity(0, cpus)
        return f'bound to {len(cpus)} cpus local to gpu {device_index} ({min(cpus)}-{max(cpus)})'
    except Exception as e:                                  # noqa: BLE001 - best effort
        return f'affinity unchanged ({type(e).__name__}: {e})'


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


class _DeviceArray:
    """A raw device address as something torch.as_tensor understands (__cuda_array_interface__ v2)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (ptr, False), 'version': 2,
                                         'strides': None}


class PeerComm:
    """Peer-memory communicator over the ranks of a torch.distributed group on ONE box (include/d3f.h d3f_comm_*).

    Collective constructor: every rank allocates its segment (gathered dist / valid_mask arrays for
    `capacity_points` points, double-buffered, plus `staging_bytes` for broadcasts), the 64-byte CUDA IPC handles
    are all-gathered through the group, and every rank maps its peers' segments.  torch.distributed is used for
    this handshake only; the data path afterwards is the field kernel's own remote stores and flags."""

    def __init__(self, capacity_points: int, group=None, device=None, staging_bytes: int = 64 << 20, _connect: bool = True):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.capacity = int(capacity_points)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self._views: Dict[Tuple[int, str], torch.Tensor] = {}
        self._comm = None
        self._handle = b''
        with torch.cuda.device(self.device):
            self._comm, self._handle = _native.comm_create(self.rank, self.world, self.capacity, int(staging_bytes))
        if _connect:
            self.connect(self.exchange_handles())
            if self.world > 1:
                dist.barrier(group)

    def exchange_handles(self, handle: Optional[bytes] = None) -> bytes:
        """All-gather of the 64-byte IPC handles, rank order (collective)."""
        if self.world == 1:
            return b''
        mine = torch.tensor(list(handle if handle is not None else self._handle), dtype=torch.uint8, device=self.device)
        allh = torch.empty(self.world * _native.D3F_IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        return bytes(allh.cpu().tolist())

    def connect(self, handles: bytes) -> None:
        if self.world > 1:
            with torch.cuda.device(self.device):
                _native.comm_connect(self._comm, handles)

    def _view(self, ptr: int, kind: str) -> torch.Tensor:
        t = self._views.get((ptr, kind))
        if t is None:
            if kind == 'dist':
                t = torch.as_tensor(_DeviceArray(ptr, self.capacity, '<f4'), device=self.device)
            else:
                t = torch.as_tensor(_DeviceArray(ptr, self.capacity, '|u1'), device=self.device).view(torch.bool)
            self._views[(ptr, kind)] = t
        return t

    def eval_allgather(self, V, H, W, pose_p, K_p, depth_p, pts_p, n, keys, outs, base, block, stride, flags, mu, stream):
        """d3f_eval_allgather; returns the gathered (capacity,) dist / valid_mask tensors of this call (views of the
        communicator's segment: valid until the call after next)."""
        d, v = _native.eval_allgather(self._comm, V, H, W, pose_p, K_p, depth_p, pts_p, n, keys, outs,
                                      int(base), int(block), int(stride), flags, mu, stream)
        if self.capacity == 0:
            return (torch.empty(0, dtype=torch.float32, device=self.device),
                    torch.empty(0, dtype=torch.bool, device=self.device))
        return self._view(d, 'dist'), self._view(v, 'valid')

    def broadcast(self, t: torch.Tensor, root: int = 0) -> None:
        """Replicate a contiguous device tensor from `root` (d3f_comm_broadcast), on torch's current stream."""
        if not (t.is_cuda and t.is_contiguous()):
            raise ValueError('broadcast needs a contiguous CUDA tensor')
        with torch.cuda.device(self.device):
            _native.comm_broadcast(self._comm, t.data_ptr(), t.numel() * t.element_size(), int(root),
                                   torch.cuda.current_stream(self.device).cuda_stream)

    def check(self) -> None:
        """Synchronise the current stream and raise if a wait on a peer timed out."""
        with torch.cuda.device(self.device):
            _native.comm_status(self._comm, torch.cuda.current_stream(self.device).cuda_stream)

    def close(self) -> None:
        if self._comm is not None:
            self._views.clear()
            with torch.cuda.device(self.device):
                _native.comm_destroy(self._comm)
            self._comm = None


def make_peer_comm(capacity_points: int, group=None, device=None, staging_bytes: int = 64 << 20) -> Optional[PeerComm]:
    """PeerComm if every rank of the group could allocate its segment and map its peers', else None on EVERY rank
    (callers then use the NCCL transport).  A failure on one rank never leaves the others waiting in a collective: the
    ranks agree after each stage (create, map) with an all-reduce."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())

    def agree(ok: bool) -> bool:
        if world == 1:
            return ok
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(flag.item())

    def warn(stage, e):
        import warnings
        warnings.warn(f'peer-memory communicator unavailable on this rank ({stage}): {e}', RuntimeWarning)

    comm = None
    try:
        comm = PeerComm(capacity_points, group=group, device=device, staging_bytes=staging_bytes, _connect=False)
    except (_native.D3FError, _native.NativeLibraryError, RuntimeError) as e:
        warn('segment allocation', e)
    if not agree(comm is not None):
        if comm is not None:
            comm.close()
        return None
    handles = comm.exchange_handles()
    ok = True
    try:
        comm.connect(handles)
    except (_native.D3FError, RuntimeError) as e:          # e.g. no peer access between two of the GPUs
        warn('mapping the peers', e)
        ok = False
    if not agree(ok):
        comm.close()
        return None
    if world > 1:
        dist.barrier(group)
    return comm


def broadcast_observation_peer(obs: Dict[str, object], comm: PeerComm, src: int = 0) -> None:
    """broadcast_observation over the peer-memory communicator instead of NCCL."""
    for k in sorted(obs):
        v = obs[k]
        if isinstance(v, torch.Tensor) and v.is_cuda:
            comm.broadcast(v, root=src)


@dataclass
class Share:
    """A rank's share of a point array, prepared once and reused by every eval_sharded call on those points."""
    n: int                       # points in the full array
    local: torch.Tensor          # (n_local,3) this rank's points, contiguous
    index: Optional[torch.Tensor]   # positions of the local points in the full array (block-interleaved), else None
    span: Tuple[int, int]        # contiguous slab [start, end) (block=None)
    base: int                    # gathered index of local point i: base + (i // block) * stride + i % block
    block: int
    stride: int
    interleave: Optional[int]    # the `block` argument this share was planned with


def plan_share(pts: torch.Tensor, group=None, block: Optional[int] = None) -> Share:
    """Slice this rank's share of `pts` (the full (N,3) array, identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(pts.shape[0])
    if block is None:
        start, end = shard_range(n, rank, world)
        return Share(n, pts[start:end].contiguous(), None, (start, end), start, max(end - start, 1), 0, None)
    idx = block_interleaved_index(n, rank, world, block).to(pts.device)
    return Share(n, pts.index_select(0, idx).contiguous(), idx, (0, 0), rank * block, block, world * block, block)


def eval_sharded(eval_fn: Callable[..., Dict[str, torch.Tensor]], pts: Optional[torch.Tensor],
                 return_names: Iterable[str] = (), gather: Sequence[str] = ('dist', 'valid_mask'),
                 channels: Optional[Dict[str, int]] = None, group=None, block: Optional[int] = None,
                 comm: Optional[PeerComm] = None, share: Optional[Share] = None) -> Dict[str, object]:
    """Evaluate rank-local shares of `pts` (the full (N,3) array, identical on every rank) and all-gather
    the keys listed in `gather`.

    block=None   contiguous slabs (shard_range).
    block=B      blocks of B consecutive points (e.g. one x-plane of a grid: ny*nz) dealt round-robin: every rank gets
                 the same spatial mix, which matters because a point's cost depends on how many views see it (measured:
                 contiguous slabs of the benchmark workspace cost 0.66-0.92 ms, interleaved shares 0.78 ms each).
                 Needs N % (B*world) == 0.  Gathered keys come back in canonical point order; keys left sharded come
                 back as the rank's share together with 'index' (their positions in the full array).

    eval_fn(local_pts, return_names=..., out=...) -> dict — e.g. the bound Fusion.eval: must write into the tensors
    of `out` when given (so the kernel fills the gather slot in place) and return tensors on pts' device.
    channels: C of each gathered key other than dist/valid_mask (needed to size its gather buffer).

    comm=PeerComm  dist / valid_mask are gathered INSIDE the field kernel through peer memory (one launch, no NCCL
                   call); eval_fn must then be the bound `Fusion.eval` and `gather` a subset of dist / valid_mask.
                   The gathered tensors are views of the communicator's double-buffered arrays: valid until the
                   call after next.
    share=Share    from plan_share(pts, group, block): skips re-slicing the rank's points on every call (`pts` may
                   then be None).

    Returns {'shard': (start, end) | 'index': positions, gathered keys -> full (N, ...) tensors in canonical order,
    other keys -> the rank's share}.
    """
    names = list(return_names)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if share is not None:
        block = share.interleave
    if comm is not None:
        return _eval_peer(eval_fn, pts, names, gather, group, block, comm, share)
    if share is not None and pts is None:
        raise ValueError('the NCCL transport needs the full point array (pts)')
    n = int(pts.shape[0])
    if block is not None:
        return _eval_interleaved(eval_fn, pts, names, gather, channels, group, block, world, rank, share)
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
    res = eval_fn(share.local if share is not None else pts[start:end], return_names=names, out=out)
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


def _eval_peer(eval_fn, pts, names, gather, group, block, comm: PeerComm, share: Optional[Share]):
    fusion = getattr(eval_fn, '__self__', None)
    if fusion is None or not hasattr(fusion, '_run'):
        raise ValueError('comm= needs eval_fn to be the bound Fusion.eval of a d3fields_b200.Fusion')
    extra = [k for k in gather if k not in ('dist', 'valid_mask')]
    if extra:
        raise ValueError(f'the peer-memory gather carries dist / valid_mask only; gather {extra} over NCCL (comm=None)')
    if share is None:
        share = plan_share(pts, group, block)
    if share.n > comm.capacity:
        raise ValueError(f'{share.n} points exceed the communicator capacity {comm.capacity}')
    res = fusion._run(share.local, names, False, False, gather=(comm, share.base, share.block, share.stride))
    result: Dict[str, object] = {'index': share.index} if share.index is not None else {'shard': share.span}
    for k, v in res.items():
        if k in ('dist', 'valid_mask'):
            if k in gather:
                result[k] = v[:share.n]
        else:
            result[k] = v
    return result


def _eval_interleaved(eval_fn, pts, names, gather, channels, group, block, world, rank, share=None):
    n = int(pts.shape[0])
    dev = pts.device
    per = n // world
    if share is not None:
        idx, local = share.index, share.local
    else:
        idx = block_interleaved_index(n, rank, world, block).to(dev)
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
    res = eval_fn(local, return_names=names, out=out)
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
