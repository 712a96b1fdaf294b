#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: M query-points/s through Fusion.eval (V=4, C=1024).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (torch port)

A "step" is one Fusion.eval(pts, ['dino_feats']) over the workload cfg2a of SURVEY.md §8d:
1 000 000 voxel-grid points (z fastest), 4 ring views 480x640, a (48,64,1024) float32 descriptor map
per view (the reference samples DINOv2 at (H//10, W//10), fusion.py:695-696), synthetic seed 0.
At N GPUs every rank evaluates 1M points of an N-times finer grid, its x-planes dealt round-robin (weak scaling,
no data-path collective for the descriptor field — SURVEY.md §8e) and the compact fields dist/valid_mask
are all-gathered in place over NCCL inside the timed step.

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


def shard_points(rank, world, n_per_gpu):
    """Rank's share of a grid refined world-times along x (z fastest, then y, then x): the x-planes rank, rank+world,
    rank+2*world, ...  Every rank then sees the same spatial mix of the scene.  Contiguous x-slabs do not: measured
    on one GPU (tools/slab_balance.py) the 8 slabs of this workspace cost 0.66-0.92 ms each (22-64 % of their points
    are seen by a camera) while the 8 interleaved shares cost 0.78 ms each, and the step is the max over ranks."""
    gx, gy, gz = CFG['grid']
    assert n_per_gpu % (gy * gz) == 0
    nx_local = n_per_gpu // (gy * gz)
    pts = S.grid_points(nx_local * world, gy, gz).reshape(nx_local * world, gy * gz, 3)
    return np.ascontiguousarray(pts[rank::world].reshape(-1, 3))


def cpu_port_rate(scene, pts, seconds_budget=20.0, max_reps=5):
    """Time oracle/torch_port.eval_chunk (the reference's operator sequence) on one 60 000-point chunk."""
    import torch
    from oracle import torch_port as TP
    torch.set_num_threads(os.cpu_count() or 1)
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
    return chunk.shape[0] / best / 1e6, torch.get_num_threads(), reps, chunk.shape[0]


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle/torch_port.py; the Python reference cannot
    travel to the GPU box) on the host cores, each step one 60 000-point chunk of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import torch_port as TP
    torch.set_num_threads(os.cpu_count() or 1)
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
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--points-per-gpu', type=int, default=CFG['n'])
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--scattered', action='store_true', help='keypoint-like points with no locality instead of the grid')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # Watchdog: a whole run takes well under two minutes.  If a collective ever hangs (a rank died, a rendezvous
    # problem), fail loudly instead of holding N GPUs until somebody's time limit expires.
    limit = float(os.environ.get('D3F_BENCH_WATCHDOG_S', '900'))
    def _abort():
        sys.stderr.write(f'bench.py watchdog: no result after {limit:.0f} s, aborting\n')
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(limit, _abort)
    wd.daemon = True
    wd.start()

    import torch
    import torch.distributed as dist
    from d3fields_b200 import Fusion, _native

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback for the field query)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import datetime
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # The only collective here is a 5 B/point all-gather: NVLink-SHARP multicast buys nothing for it, and its setup
        # is the slowest and most fragile part of communicator creation when jobs of different rank counts follow each
        # other on one box.  A bounded timeout turns any rendezvous problem into an error instead of a hang.
        os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=240))

    n = args.points_per_gpu
    V, H, W = CFG['V'], CFG['H'], CFG['W']
    h, w, C = CFG['feat']
    sc = S.make_scene(V, H, W, seed=0, feat=CFG['feat'])
    pts_np = S.scattered_points(n, rank) if args.scattered else shard_points(rank, world, n)
    f = Fusion(num_cam=V, device=str(dev))
    f.update({'depth': sc.depth, 'pose': sc.pose, 'K': sc.K, 'dino_feats': sc.maps['dino_feats']})
    pts = torch.from_numpy(pts_np).to(dev)
    names = ['dino_feats']

    # N > 1: in-place all-gather layout for the compact fields (d3fields_b200/sharded.py): the kernel writes the
    # rank's dist / valid_mask straight into its slot of the gather buffers, then one all_gather per buffer
    if world > 1:
        assert n % 4 == 0
        g_pack = torch.empty((world, 5 * n), dtype=torch.uint8, device=dev)      # per rank: n float32 dist | n bool valid
        slot = {'dist': g_pack[rank, :4 * n].view(torch.float32), 'valid_mask': g_pack[rank, 4 * n:].view(torch.bool)}

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    side = torch.cuda.Stream(dev) if world > 1 else None

    def step():
        if world == 1:
            return f.eval(pts, return_names=names)
        # N > 1: the compact fields and their collective run on a side stream — the light dist/valid kernel writes this
        # rank's slot, one in-place all_gather follows — while the descriptor kernel runs on the main stream; the
        # step ends when both have finished.
        main = torch.cuda.current_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            f.eval(pts, return_names=[], out=slot)
            dist.all_gather_into_tensor(g_pack.view(-1), g_pack[rank])
        out = f.eval(pts, return_names=names)
        main.wait_stream(side)
        return out

    clk = ClockSampler(local)
    clk.__enter__()                    # samples cover the warm-up and the timed region (the latter lasts ~20 ms)
    for _ in range(args.warmup):
        out = step()
        flush.zero_()
    torch.cuda.synchronize(dev)
    # Keep the GPU under this load for ~0.5 s so nvidia-smi samples it.  The count is FIXED, never time-based: every
    # step contains a collective at N > 1, so all ranks must run exactly the same number of steps.
    for _ in range(500):
        out = step()
        flush.zero_()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = _native.launch_count()
    torch.cuda.synchronize(dev)
    if True:
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
    clk.__exit__(None, None, None)
    launches = _native.launch_count() - launches0
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

    # ---- end to end: host points in, host results out, through the public API --------------------------
    e2e = None
    if not args.no_e2e:
        pts_h = torch.from_numpy(pts_np).pin_memory()
        host_out = {'dist': torch.empty(n, dtype=torch.float32).pin_memory(),
                    'valid_mask': torch.empty(n, dtype=torch.bool).pin_memory(),
                    'dino_feats': torch.empty((n, C), dtype=torch.float32).pin_memory()}
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
        same = bool(torch.equal(host_out['dino_feats'][::997], out['dino_feats'][::997].cpu()))
        e2e = {'value': world * n * args.e2e_steps / dt / 1e6, 'unit': UNIT,
               'h2d_bytes_per_step': int(world * n * 12), 'd2h_bytes_per_step': int(world * n * (5 + 4 * C)),
               'steps': args.e2e_steps, 'ms_per_step': dt / args.e2e_steps * 1e3, 'matches_device_path': same,
               'api': 'Fusion.eval_host -> d3f_eval_host (pinned host buffers, slab-pipelined copies)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    B = algorithmic_bytes(n, V, H, W, [(h, w, C, 4)])
    kernel_ms = float(np.mean(step_ms)) if world == 1 else ms_per_step
    achieved = B / (kernel_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(), 'peak_source': peak_src, 'algorithmic_bytes_per_launch': B,
                'kernel': variant, 'kernel_ms': kernel_ms}
    cpu = None
    if world == 1 and not args.no_cpu:
        rate, cores, reps, cn = cpu_port_rate(sc, pts_np)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'best of {reps} runs of one batch_eval chunk ({cn} points of cfg2a, full C=1024) through '
                         f'oracle/torch_port.py (the reference operator sequence on torch CPU)'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'cfg2a: {n} grid points per GPU (z fastest), V={V} views {H}x{W}, dino_feats '
                               f'({h},{w},{C}) f32 per view, return_names=[dino_feats]' + (' [scattered]' if args.scattered else ''),
                   'points_per_gpu': n, 'global_points': world * n, 'sharding': f'x-planes of an {world}x finer grid dealt round-robin to {world} ranks (equal spatial mix per rank)',
                   'collective': 'one in-place all_gather of the packed (dist f32 | valid_mask u8) slots, 5 B/point, on a side stream overlapping the descriptor kernel, joined inside the timed step' if world > 1 else 'none',
                   'l2': 'outputs 4.1 GB per step exceed L2; plus a 256 MiB flush between timed steps (not timed)',
                   'timing': 'CUDA events per step on the launching stream, summed; max over ranks'},
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches),
        'clocks': clk.summary(), 'wall_s_timed_region': t_wall, 'step_ms_min': float(min(step_ms)),
        'step_ms_max': float(max(step_ms)), 'checksum': checksum,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
