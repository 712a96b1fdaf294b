#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: M query-points/s through Fusion.eval (V=4, C=1024).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (torch port)

A "step" is one Fusion.eval(pts, ['dino_feats']) over the workload cfg2a of SURVEY.md §8d:
1 000 000 voxel-grid points (z fastest), 4 ring views 480x640, a (48,64,1024) float32 descriptor map
per view (the reference samples DINOv2 at (H//10, W//10), fusion.py:695-696), synthetic seed 0.
At N > 1 GPUs the step is the product's sharded path, d3fields_b200.sharded.eval_sharded, on a 2M*N-point grid (N=8:
cfg4's 16M points), x-planes dealt round-robin; the descriptor field stays sharded and dist/valid_mask
of every point are gathered to every rank inside the field kernel through peer memory (d3f_eval_allgather) — one
launch per step, no NCCL call.  After the timed region every rank checks the gathered field bit for bit against its own
single-rank evaluation of all points.

One JSON line is printed by rank 0.  `value` is device-timed with inputs resident in HBM; `e2e` is the
same metric through Fusion.eval with HOST points and HOST results (d3f_eval_host: H2D of the points and
D2H of every output inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from d3fields_b200 import scene as S  # noqa: E402

METRIC = 'M query-points/sec through Fusion.eval (V=4, C=1024)'
UNIT = 'Mpts/s'
CFG = S.CONFIGS['cfg2a']
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def algorithmic_bytes(n, V, H, W, keys):
    """points in + depth (each texel at most once, else one per point-view) + each sampled
    volume at most once (else 4 corners per point-view) + dist f32, valid u8 and float32 outputs out."""
    b = 12 * n + min(V * H * W * 4, 4 * n * V) + n * 5
    for (h, w, C, s) in keys:
        b += min(V * h * w * C * s, 4 * n * V * C * s) + n * 4 * C
    return b


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if there is one."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            return json.load(fh).get('dram_bytes_per_launch')
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def wait_started(self, timeout=15.0):
        """nvidia-smi can take more than a second to print its first row on an 8-GPU box: do not start the load before
        it is sampling, or the (short) timed region is over before the first sample."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)
        return bool(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (('hw_slowdown', 4), ('hw_thermal_slowdown', 5), ('sw_thermal_slowdown', 6), ('sw_power_cap', 7)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


MULTI_GPU_POINTS = 2_000_000       # per GPU at N > 1: cfg4 of BASELINE.json is 16M points over 8 GPUs


def global_grid(world, n_per_gpu):
    """The workspace grid all ranks share at N GPUs: cfg2a's 100x100 (y,z) planes, refined along x so that it holds
    world*n_per_gpu points (N=8, 2M per GPU: 1600x100x100 = cfg4's 16M-point grid).  z fastest, then y, then x."""
    gx, gy, gz = CFG['grid']
    assert n_per_gpu % (gy * gz) == 0
    return n_per_gpu // (gy * gz) * world, gy, gz


def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on (cgroup / affinity), not the box's total."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_rate(scene, pts, seconds_budget=20.0, max_reps=5):
    """Time oracle/torch_port.eval_chunk (the reference's operator sequence) on one 60 000-point chunk."""
    import torch
    from oracle import torch_port as TP
    torch.set_num_threads(host_threads())
    obs = TP.obs_from_scene(scene)
    chunk = torch.from_numpy(np.ascontiguousarray(pts[:TP.CHUNK]))
    torch.set_grad_enabled(False)
    TP.eval_chunk(obs, scene.H, scene.W, chunk[:2000], ['dino_feats'])          # touch code paths
    best, t_all, reps = None, 0.0, 0
    while reps < max_reps and (reps == 0 or t_all + (best or 0) < seconds_budget):
        t0 = time.perf_counter()
        TP.eval_chunk(obs, scene.H, scene.W, chunk, ['dino_feats'])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        t_all += dt
        reps += 1
    torch.set_grad_enabled(True)
    return chunk.shape[0] / best / 1e6, torch.get_num_threads(), reps, chunk.shape[0]


def torch_gpu_rate(scene, pts_dev, dev, reps=3):
    """The reference's real deployment (device='cuda:0' is its default, fusion.py:203): its operator sequence run by
    torch on this same GPU, batch_eval's 60 000-point chunks over the whole workload."""
    import torch
    from oracle import torch_port as TP
    obs = {k: v.to(dev) for k, v in TP.obs_from_scene(scene).items()}
    TP.batch_eval(obs, scene.H, scene.W, pts_dev[:120000], ['dino_feats'])
    torch.cuda.synchronize(dev)
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        TP.batch_eval(obs, scene.H, scene.W, pts_dev, ['dino_feats'])
        b.record()
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return pts_dev.shape[0] / best / 1e3, best


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle/torch_port.py; the Python reference cannot
    travel to the GPU box) on the host cores, each step one 60 000-point chunk of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import torch_port as TP
    torch.set_num_threads(host_threads())
    torch.set_grad_enabled(False)
    sc = S.make_scene(CFG['V'], CFG['H'], CFG['W'], seed=0, feat=CFG['feat'])
    pts = S.config_points('cfg2a')
    obs = TP.obs_from_scene(sc)
    n = TP.CHUNK
    chunks = [torch.from_numpy(np.ascontiguousarray(pts[i * n:(i + 1) * n])) for i in range(len(pts) // n)]
    for i in range(args.warmup):
        TP.eval_chunk(obs, sc.H, sc.W, chunks[i % len(chunks)][:6000], ['dino_feats'])
    t0 = time.perf_counter()
    for i in range(args.steps):
        TP.eval_chunk(obs, sc.H, sc.W, chunks[i % len(chunks)], ['dino_feats'])
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e6
    sample = f'{args.steps} steps x one batch_eval chunk of {n} points of cfg2a (warm-up on 6000-point chunks)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'cfg2a: 1M grid points, 4 views 480x640, dino_feats (48,64,1024) f32; CPU arm runs '
                               '60 000-point chunks (reference fusion.py:527)', 'points_per_step': n},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample,
                         'cores_on_box': os.cpu_count()},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def wall_rate(fn, n_pts, steps, sync):
    fn(); fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    sync()
    dt = time.perf_counter() - t0
    return n_pts * steps / dt / 1e6, dt / steps * 1e3


def e2e_variants(f, sc, pts_np, dev, steps):
    """End to end (host points in, host results out) for the return sets the reference's callers actually pull to the
    host: [] (vis_repr.py:93-97 -> extract_mesh), ['mask'] (fusion.py:1428) and the PCA(3)'d descriptors
    (fusion.py:1386-1392) — 5, 37 and 17 bytes per point instead of the 4 KB/point descriptor field."""
    import torch
    n = len(pts_np)
    pts_h = torch.from_numpy(pts_np).pin_memory()
    sync = lambda: torch.cuda.synchronize(dev)
    res = {}
    out0 = {'dist': torch.empty(n, dtype=torch.float32).pin_memory(), 'valid_mask': torch.empty(n, dtype=torch.bool).pin_memory()}
    v, ms = wall_rate(lambda: f.eval_host(pts_h, [], out=out0), n, steps, sync)
    res['dist_valid'] = {'value': v, 'unit': UNIT, 'ms_per_step': ms, 'h2d_bytes_per_step': n * 12, 'd2h_bytes_per_step': n * 5,
                         'api': "Fusion.eval_host(pts, []) (vis_repr.py:93)"}
    if 'mask' in f.curr_obs_torch:
        C = int(f.curr_obs_torch['mask'].shape[-1])
        outm = dict(out0, mask=torch.empty((n, C), dtype=torch.float32).pin_memory())
        v, ms = wall_rate(lambda: f.eval_host(pts_h, ['mask'], out=outm), n, steps, sync)
        res['mask'] = {'value': v, 'unit': UNIT, 'ms_per_step': ms, 'h2d_bytes_per_step': n * 12, 'd2h_bytes_per_step': n * (5 + 4 * C),
                       'api': "Fusion.eval_host(pts, ['mask']) (fusion.py:1428)"}
    gen = torch.Generator().manual_seed(5)
    comp = torch.randn(3, CFG['feat'][2], generator=gen).to(dev)
    mean = torch.randn(CFG['feat'][2], generator=gen).to(dev)
    pts_d = torch.empty((n, 3), dtype=torch.float32, device=dev)
    y_h = torch.empty((n, 3), dtype=torch.float32).pin_memory()

    def pca_step():
        pts_d.copy_(pts_h, non_blocking=True)
        r = f.eval_pca(pts_d, 'dino_feats', mean, comp)
        y_h.copy_(r['dino_feats_pca'], non_blocking=True)
        out0['dist'].copy_(r['dist'], non_blocking=True)
        out0['valid_mask'].copy_(r['valid_mask'], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    v, ms = wall_rate(pca_step, n, steps, sync)
    res['pca3'] = {'value': v, 'unit': UNIT, 'ms_per_step': ms, 'h2d_bytes_per_step': n * 12, 'd2h_bytes_per_step': n * 17,
                   'api': 'pinned H2D + Fusion.eval_pca + pinned D2H (fusion.py:1386-1392 applies the PCA on the host)'}
    return res


def cfg5_frames(f, sc, dev, world, rank, frames=64, keypoints=262144):
    """BASELINE.json configs[4]: 64 frames x 4 views x 256k keypoints, descriptor eval + PCA(3), frame-parallel
    (each rank owns its frames' observation, no communication).  Per frame a fresh descriptor volume and
    depth (seeded by the frame number, generated on the device before the timed region), 262 144 scattered keypoints."""
    import torch
    V, H, W = CFG['V'], CFG['H'], CFG['W']
    h, w, C = CFG['feat']
    mine = list(range(rank, frames, world))
    gen = torch.Generator(device=dev)
    vols, depths = [], []
    depth0 = torch.from_numpy(sc.depth).to(dev)
    for fr in mine:
        gen.manual_seed(1000 + fr)
        vols.append(torch.randn((V, h, w, C), device=dev, generator=gen))
        depths.append((depth0 + (torch.rand(depth0.shape, device=dev, generator=gen) - 0.5) * 0.004) * (depth0 > 0))
    kp = torch.from_numpy(S.scattered_points(keypoints, 17)).to(dev)
    g2 = torch.Generator().manual_seed(5)
    comp = torch.randn(3, C, generator=g2).to(dev)
    mean = torch.randn(C, generator=g2).to(dev)
    keep_v, keep_d = f.curr_obs_torch['dino_feats'], f.curr_obs_torch['depth']

    def run(mode):
        outs = None
        for i in range(len(mine)):
            f.curr_obs_torch['dino_feats'], f.curr_obs_torch['depth'] = vols[i], depths[i]
            if mode == 'pca':
                outs = f.eval_pca(kp, 'dino_feats', mean, comp)
            else:
                outs = f.eval(kp, return_names=['dino_feats'], binned=(mode == 'desc_binned'))
        return outs

    res = {}
    try:
        for mode in ('pca', 'desc', 'desc_binned'):
            run(mode)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(mode); b.record()
            torch.cuda.synchronize(dev)
            res[mode] = a.elapsed_time(b)
    finally:
        f.curr_obs_torch['dino_feats'], f.curr_obs_torch['depth'] = keep_v, keep_d
    return res, len(mine)


def tracking_latency(f, dev, iters=100):
    """Row f-3: the reference's rigid_tracking loop (fusion.py:1643-1665) — 100 Adam iterations of forward + backward on
    num_inst*100 points — as one CUDA-graph replay, against the same loop launched eagerly and against the reference
    operator sequence with torch autograd on the same GPU."""
    import torch
    from d3fields_b200.tracking import FusedRigidTracker, RigidTracker
    from oracle import torch_port as TP
    I, P, C = 4, 100, CFG['feat'][2]
    pts = torch.from_numpy(S.scattered_points(I * P, 23, sigma=0.12)).to(dev).reshape(I, P, 3)
    src = f.eval(pts.reshape(-1, 3), return_names=['dino_feats'])['dino_feats']
    moved = pts + 0.004
    obs = {k: v for k, v in f.curr_obs_torch.items() if isinstance(v, torch.Tensor)}
    out = {}
    for name, kw in (('single_launch_graph', dict(fused=1)), ('fused_graph', dict(fused=4)), ('graph', dict(graph=True)),
                     ('eager', dict(graph=False)),
                     ('torch_reference_ops', dict(graph=False, eval_fn=lambda p, names: TP.eval_chunk(obs, f.H, f.W, p, names)))):
        fused = kw.pop('fused', 0)
        tr = (FusedRigidTracker(f, I, P, C, iters=iters, single_launch=fused == 1) if fused
              else RigidTracker(f, I, P, C, iters=iters, **kw))
        tr.track(src, moved)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            tr.track(src, moved)
        torch.cuda.synchronize(dev)
        out[name] = (time.perf_counter() - t0) / reps / iters * 1e6
    return {'points': I * P, 'iterations': iters, 'us_per_iteration': out,
            'what': 'track() wall time / iterations: forward + backward + Adam step (reference fusion.py:1643-1665). '
                    'single_launch_graph: ONE launch per iteration (d3f_track_step: a CTA per point, texels in registers for '
                    'forward and backward, the last CTA of an instance takes the Adam step), 100 in one CUDA graph; fused_graph: '
                    '4 launches per iteration (d3f_eval, d3f_track_loss_grad, d3f_eval_backward, d3f_track_update) in one CUDA graph; '
                    'graph / eager: torch autograd around the two field kernels; torch_reference_ops: the reference operator sequence'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--points-per-gpu', type=int, default=0, help='default: 1M at N=1 (cfg2a), 2M at N>1 (cfg4 at N=8)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip cfg5 / tracking / torch-GPU / e2e variants')
    ap.add_argument('--transport', default='auto', choices=['auto', 'peer', 'nccl'])
    ap.add_argument('--scattered', action='store_true', help='keypoint-like points with no locality instead of the grid')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # Watchdog: a whole run takes a few minutes.  If a collective ever hangs (a rank died, a rendezvous
    # problem), fail loudly instead of holding N GPUs until somebody's time limit expires.
    limit = float(os.environ.get('D3F_BENCH_WATCHDOG_S', '900'))
    def _abort():
        sys.stderr.write(f'bench.py watchdog: no result after {limit:.0f} s, aborting\n')
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(limit, _abort)
    wd.daemon = True
    wd.start()

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')

    import torch
    import torch.distributed as dist
    from d3fields_b200 import Fusion, _native
    from d3fields_b200 import sharded as SH

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback for the field query)')
    local %= max(torch.cuda.device_count(), 1)      # a launcher that exposes one device per process
    # before any pinned allocation: this rank's threads and host buffers live next to its GPU
    numa = SH.bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import datetime
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL is used for the handshake (IPC handles, barriers, the final max-reduction) — and for the data path only
        # with --transport nccl.  NVLink-SHARP multicast buys nothing for that and its setup is the slowest and most
        # fragile part of communicator creation; a bounded timeout turns any rendezvous problem into an error.
        os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=240))

    n = args.points_per_gpu or (CFG['n'] if world == 1 else MULTI_GPU_POINTS)
    V, H, W = CFG['V'], CFG['H'], CFG['W']
    h, w, C = CFG['feat']
    sc = S.make_scene(V, H, W, seed=0, feat=CFG['feat'], num_inst=8)
    f = Fusion(num_cam=V, device=str(dev))
    f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
    f.set_instance_masks(torch.from_numpy(sc.maps['mask']), as_uint8=True)
    names = ['dino_feats']

    comm, share, transport = None, None, 'none'
    if world == 1:
        pts_np = S.scattered_points(n, rank) if args.scattered else S.config_points('cfg2a')[:n] if n <= CFG['n'] else S.grid_points(*global_grid(1, n))
        pts = torch.from_numpy(pts_np).to(dev)
        n_total = n
    else:
        # N > 1: the product's sharded path (d3fields_b200/sharded.py).  Every rank holds the whole grid (192 MB at 16M
        # points); x-planes are dealt round-robin so every rank sees the same spatial mix (a point's cost depends on how
        # many views see it: contiguous slabs of this workspace differ by 40 %, tools/slab_balance.py); dist /
        # valid_mask of every point are gathered to every rank INSIDE the field kernel through peer memory.
        gx, gy, gz = global_grid(world, n)
        n_total = gx * gy * gz
        pts_full = torch.from_numpy(S.grid_points(gx, gy, gz)).to(dev)
        block = gy * gz
        share = SH.plan_share(pts_full, block=block)
        pts_np = None
        if args.transport in ('auto', 'peer'):
            comm = SH.make_peer_comm(n_total, device=dev, staging_bytes=64 << 20)
            if comm is None and args.transport == 'peer':
                raise SystemExit('--transport peer: the peer-memory communicator did not come up')
        transport = 'peer' if comm is not None else 'nccl'
        if comm is not None:                 # observation replication over the same communicator (once per update())
            SH.broadcast_observation_peer({k: v for k, v in f.curr_obs_torch.items() if isinstance(v, torch.Tensor)}, comm, src=0)
            comm.check()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step():
        if world == 1:
            return f.eval(pts, return_names=names)
        if comm is not None:
            return SH.eval_sharded(f.eval, None, names, comm=comm, share=share)
        return SH.eval_sharded(f.eval, pts_full, names, gather=('dist', 'valid_mask'), share=share)

    clk = ClockSampler(local)
    if rank == 0:
        clk.__enter__()                # samples cover the warm-up, the load loop and the timed region (~20 ms)
        clk.wait_started()
        clk.rows.clear()               # keep only samples taken under load
    if world > 1:
        dist.barrier()
    for _ in range(args.warmup):
        out = step()
        flush.zero_()
    torch.cuda.synchronize(dev)
    # Keep the GPU under this load for ~0.5 s so nvidia-smi samples it.  The count is FIXED, never time-based: every
    # step contains a collective at N > 1, so all ranks must run exactly the same number of steps.
    for _ in range(int(os.environ.get('D3F_BENCH_LOAD_STEPS', '600'))):      # (shortened under ncu: a launch list of 600 warm-up kernels helps nobody)
        out = step()
        flush.zero_()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = _native.launch_count()
    torch.cuda.synchronize(dev)
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        starts[i].record()
        out = step()
        ends[i].record()
        flush.zero_()                         # L2 flush between timed steps, outside the event pairs
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    if rank == 0:
        clk.__exit__(None, None, None)
    launches = _native.launch_count() - launches0
    if comm is not None:
        comm.check()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    variant = _native.last_variant(0)
    checksum = float(out['dino_feats'][::997].double().abs().sum().item())

    # ---- N > 1, outside the timed region: the gathered field equals a single-rank evaluation, bit for bit ------------
    parity = None
    if world > 1:
        single = f.eval(pts_full, return_names=[])
        ok_d = bool(torch.equal(out['dist'], single['dist']))
        ok_v = bool(torch.equal(out['valid_mask'], single['valid_mask']))
        mine = f.eval(share.local[:100000], return_names=names)
        ok_f = bool(torch.equal(out['dino_feats'][:100000], mine['dino_feats']))
        flag = torch.tensor([int(ok_d and ok_v and ok_f)], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {'gathered_dist_bit_identical_to_single_rank': ok_d, 'gathered_valid_mask_identical': ok_v,
                  'local_descriptor_rows_identical': ok_f, 'all_ranks': bool(flag.item()), 'points_checked': n_total}
        if not flag.item():
            sys.stderr.write(f'rank {rank}: gathered field differs from the single-rank evaluation: {parity}\n')

    # ---- N > 1: what the remaining inefficiency is made of (outside the timed region) -----------------------------------
    breakdown = None
    if world > 1:
        def timed_ms(fn, reps=10):
            fn(); fn()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                flush.zero_()
                torch.cuda.synchronize(dev)
                ts.append(a.elapsed_time(b))
            return float(np.median(ts))
        local_ms = timed_ms(lambda: f.eval(share.local, return_names=names))        # same points, no exchange, no peers
        dist.barrier()
        shard_ms = timed_ms(step)
        t = torch.tensor([local_ms, shard_ms], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        loc = [float(x[0]) for x in allt]
        shd = [float(x[1]) for x in allt]
        breakdown = {'local_kernel_ms_per_rank': loc, 'sharded_step_ms_per_rank': shd,
                     'rank_skew_ms': max(loc) - min(loc), 'exchange_overhead_ms': max(shd) - max(loc),
                     'what': 'median of 10: the rank\'s share evaluated alone (no gather) vs the sharded step; the step costs the slowest rank plus the flag exchange'}

    # ---- end to end: host points in, host results out, through the public API --------------------------
    e2e = None
    if not args.no_e2e:
        # per rank the first 1M points of its share (4.1 GB of pinned results per rank, as at N=1)
        ne = min(n, CFG['n'])
        loc_np = share.local[:ne].cpu().numpy() if world > 1 else pts_np[:ne]
        pts_h = torch.from_numpy(loc_np).pin_memory()
        host_out = {'dist': torch.empty(ne, dtype=torch.float32).pin_memory(),
                    'valid_mask': torch.empty(ne, dtype=torch.bool).pin_memory(),
                    'dino_feats': torch.empty((ne, C), dtype=torch.float32).pin_memory()}
        f.eval_host(pts_h, names, out=host_out)                       # warm-up (allocates device slabs)
        f.eval_host(pts_h, names, out=host_out)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            f.eval_host(pts_h, names, out=host_out)                   # synchronous: returns after the last D2H
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        same = bool(torch.equal(host_out['dino_feats'][::997], out['dino_feats'][:ne][::997].cpu()))
        # the ceiling this number sits under: plain pinned D2H copies of the same buffer, all ranks at once
        d_probe = out['dino_feats'][:min(ne, 262144)]
        h_probe = host_out['dino_feats'][:d_probe.shape[0]]
        h_probe.copy_(d_probe, non_blocking=True)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        tp0 = time.perf_counter()
        for _ in range(4):
            h_probe.copy_(d_probe, non_blocking=True)
        torch.cuda.synchronize(dev)
        tp = time.perf_counter() - tp0
        if world > 1:
            t = torch.tensor([tp], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tp = float(t.item())
        link_gbs = 4 * d_probe.numel() * 4 / tp / 1e9
        e2e = {'value': world * ne * args.e2e_steps / dt / 1e6, 'unit': UNIT, 'points_per_gpu': ne,
               'h2d_bytes_per_step': int(world * ne * 12), 'd2h_bytes_per_step': int(world * ne * (5 + 4 * C)),
               'steps': args.e2e_steps, 'ms_per_step': dt / args.e2e_steps * 1e3, 'matches_device_path': same,
               'host_binding': numa,
               'd2h_link_gbs_per_gpu_all_ranks_copying': link_gbs,
               'link_ceiling_mpts_s': world * link_gbs * 1e9 / (5 + 4 * C) / 1e6,
               'api': 'Fusion.eval_host -> d3f_eval_host (pinned host buffers, slab-pipelined copies)'}
        del host_out
        if world == 1 and not args.no_extras:
            e2e['other_return_sets'] = e2e_variants(f, sc, loc_np, dev, args.e2e_steps)

    # ---- secondary workloads of BASELINE.json, reported in the same line ------------------------------------------
    extras = {}
    if not args.no_extras and not args.scattered:
        ms5, nfr = cfg5_frames(f, sc, dev, world, rank)
        t = torch.tensor([ms5['pca'], ms5['desc'], ms5['desc_binned']], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extras['cfg5'] = {'frames': 64, 'keypoints_per_frame': 262144, 'frames_per_rank': nfr, 'sharding': 'frame-parallel, no communication',
                          'pca3_frames_per_s': 64 / (float(t[0]) * 1e-3), 'pca3_ms_per_frame': float(t[0]) / nfr,
                          'descriptors_frames_per_s': 64 / (float(t[1]) * 1e-3), 'descriptors_ms_per_frame': float(t[1]) / nfr,
                          'descriptors_binned_frames_per_s': 64 / (float(t[2]) * 1e-3),
                          'what': 'eval_pca (projected volume, 12 B/pt out) / eval(dino_feats) (4 KB/pt out) per frame, device-timed, max over ranks'}
        if world == 1:
            extras['tracking'] = tracking_latency(f, dev)
            rate, ms = torch_gpu_rate(sc, pts, dev)
            extras['torch_gpu_baseline'] = {'value': rate, 'unit': UNIT, 'ms_per_step': ms,
                                            'what': 'the reference operator sequence (oracle/torch_port.py, batch_eval chunks of 60 000) run by torch on this GPU — the reference defaults to device=cuda:0 (fusion.py:203)',
                                            'speedup_of_this_kernel': value / rate}

    if rank != 0:
        if comm is not None:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    B = algorithmic_bytes(n, V, H, W, [(h, w, C, 4)])
    kernel_ms = float(np.mean(step_ms)) if world == 1 else ms_per_step
    achieved = B / (kernel_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(), 'traffic_source': 'static: profiles/traffic.json, one ncu --set full capture of this kernel on cfg2a',
                'peak_source': peak_src, 'algorithmic_bytes_per_launch': B,
                'kernel': variant, 'kernel_ms': kernel_ms}
    cpu = None
    if world == 1 and not args.no_cpu:
        rate, cores, reps, cn = cpu_port_rate(sc, pts_np)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'cores_on_box': os.cpu_count(), 'kind': 'port',
               'sample': f'best of {reps} runs of one batch_eval chunk ({cn} points of cfg2a, full C=1024) through '
                         f'oracle/torch_port.py (the reference operator sequence on torch CPU)'}
    if world == 1:
        workload = f'cfg2a: {n} grid points (z fastest), V={V} views {H}x{W}, dino_feats ({h},{w},{C}) f32 per view, return_names=[dino_feats]' + (' [scattered]' if args.scattered else '')
        sharding, collective = 'none', 'none'
    else:
        gx, gy, gz = global_grid(world, n)
        workload = (f'cfg4-style: {n_total} grid points ({gx}x{gy}x{gz}, z fastest) over {world} GPUs = {n} per GPU (cfg4 of BASELINE.json '
                    f'is the N=8 case: 16M points), V={V} views {H}x{W}, dino_feats ({h},{w},{C}) f32, return_names=[dino_feats]; '
                    f'N=1 runs cfg2a (1M points)')
        sharding = f'd3fields_b200.sharded.eval_sharded: x-planes ({gy * gz} points) dealt round-robin to {world} ranks (equal spatial mix per rank); descriptor field stays sharded'
        collective = ('dist f32 + valid_mask u8 of all points gathered to every rank inside the field kernel: remote st.global into CUDA-IPC-mapped peer segments + epoch flags (d3f_eval_allgather); no NCCL call in the step'
                      if transport == 'peer' else 'in-place NCCL all_gather_into_tensor of dist and valid_mask (peer-memory transport unavailable or disabled)')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'points_per_gpu': n, 'global_points': world * n, 'sharding': sharding,
                   'collective': collective, 'transport': transport,
                   'l2': 'outputs 4.1 GB per 1M points exceed L2; plus a 256 MiB flush between timed steps (not timed)',
                   'timing': 'CUDA events per step on the launching stream, summed; max over ranks'},
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches),
        'clocks': clk.summary(), 'wall_s_timed_region': t_wall, 'step_ms_min': float(min(step_ms)),
        'step_ms_max': float(max(step_ms)), 'checksum': checksum, 'multi_gpu_parity': parity,
        'multi_gpu_breakdown': breakdown,
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
