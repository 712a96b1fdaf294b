"""rigid_tracking (reference fusion.py:1608-1685) with its optimisation loop captured in a CUDA graph.

The reference tracks `num_instance` rigid objects per frame by running 100 Adam iterations on a translation and an
axis-angle rotation per instance; every iteration is

    R = so3_exp_map(log_r);  pts = last_pts @ R + t                       (pytorch3d Transform3d, row vectors)
    out = Fusion.eval(pts, ['dino_feats'])                                 (fusion.py:1650, with autograd)
    loss = mean(|feat - src_feat| * valid) + 100 * mean(clamp(dist * valid, 0)) + |t| + |log_r|     (:1651-1662)
    loss.backward();  optimizer.step()

on ~num_instance*100 points: ~50 small kernels forward, as many backward, launch-latency bound.  Here the field query
and its gradient are one kernel each (d3f_eval / d3f_eval_backward through the autograd Function of
d3fields_b200.fusion), the surrounding few torch ops stay torch, and the whole `iters`-iteration loop is captured once
into a CUDA graph and replayed per frame: one graph launch instead of ~10 000 kernel launches.

Only the loop is restated here; so3_exp_map follows pytorch3d.transforms.so3_exp_map (Rodrigues with the angle
clamped at eps=1e-4), which the reference imports at fusion.py:1627-1628.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


def hat(v: torch.Tensor) -> torch.Tensor:
    """(n,3) -> (n,3,3) skew-symmetric matrices (pytorch3d.transforms.so3.hat)."""
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], -1).reshape(-1, 3, 3)


def so3_exp_map(log_rot: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """pytorch3d.transforms.so3_exp_map: R = I + sin(a)/a K + (1-cos(a))/a^2 K^2 with a = sqrt(clamp(|w|^2, eps))."""
    nrms = (log_rot * log_rot).sum(1)
    ang = torch.clamp(nrms, eps).sqrt()
    inv = 1.0 / ang
    f1 = inv * ang.sin()
    f2 = inv * inv * (1.0 - ang.cos())
    K = hat(log_rot)
    eye = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]
    return f1[:, None, None] * K + f2[:, None, None] * torch.bmm(K, K) + eye


def tracking_loss(out: Dict[str, torch.Tensor], src_feats: torch.Tensor, t: torch.Tensor, log_r: torch.Tensor,
                  reg_w: float = 1.0, dist_w: float = 100.0) -> torch.Tensor:
    """reference fusion.py:1651-1662 (the out-of-bounds term has weight 0 there and is left out of `loss`)."""
    valid = out['valid_mask']
    feat_loss = (torch.norm(out['dino_feats'] - src_feats, dim=-1) * valid).mean()
    dist_loss = dist_w * torch.clamp(out['dist'] * valid, min=0).mean()
    reg_loss = reg_w * (torch.norm(t) + torch.norm(log_r))
    return feat_loss + dist_loss + reg_loss


class RigidTracker:
    """Per-frame rigid tracking of `num_instance` point sets against their source descriptors.

    eval_fn(pts, names) -> dict with 'dino_feats' (differentiable), 'dist', 'valid_mask'; default: fusion.eval.
    graph=True captures the `iters`-iteration Adam loop into one CUDA graph at the first track() call; later calls
    copy their inputs into the captured buffers and replay it.  The observation tensors of `fusion` are captured by
    address: update them IN PLACE between frames (tensor.copy_), or call invalidate() after replacing them."""

    def __init__(self, fusion, num_instance: int, rand_ptcl_num: int, feat_dim: int, iters: int = 100, lr: float = 0.01,
                 reg_w: float = 1.0, dist_w: float = 100.0, graph: bool = True, name: str = 'dino_feats',
                 eval_fn: Optional[Callable] = None, device=None):
        self.fusion, self.iters, self.lr, self.reg_w, self.dist_w, self.name = fusion, iters, lr, reg_w, dist_w, name
        self.use_graph = graph
        dev = torch.device(device if device is not None else fusion.device)
        self.dev = dev
        self.I, self.P = num_instance, rand_ptcl_num
        self.t = torch.zeros(num_instance, 3, device=dev, requires_grad=True)
        self.log_r = torch.zeros(num_instance, 3, device=dev, requires_grad=True)
        self.last_pts = torch.zeros(num_instance, rand_ptcl_num, 3, device=dev)
        self.src_feats = torch.zeros(num_instance * rand_ptcl_num, feat_dim, device=dev)
        self.curr_pts = torch.zeros(num_instance * rand_ptcl_num, 3, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.opt = torch.optim.Adam([self.t, self.log_r], lr=lr, betas=(0.9, 0.999), capturable=dev.type == 'cuda')
        self.eval_fn = eval_fn or (lambda pts, names: fusion.eval(pts, return_names=names))
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    # one Adam iteration of the reference loop (fusion.py:1643-1665)
    def _iteration(self):
        R = so3_exp_map(self.log_r)
        pts = (torch.bmm(self.last_pts, R) + self.t[:, None, :]).reshape(-1, 3)
        out = self.eval_fn(pts, [self.name])
        feats = {'dino_feats': out[self.name], 'dist': out['dist'], 'valid_mask': out['valid_mask']}
        loss = tracking_loss(feats, self.src_feats, self.t, self.log_r, self.reg_w, self.dist_w)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.curr_pts.copy_(pts.detach())
        self.loss.copy_(loss.detach())

    def _reset(self):
        with torch.no_grad():
            self.t.zero_()
            self.log_r.zero_()
            for st in self.opt.state.values():
                for v in st.values():
                    if isinstance(v, torch.Tensor):
                        v.zero_()

    def invalidate(self):
        self._graph = None

    def _capture(self):
        # warm-up on a side stream (allocator, Adam state, lazy kernels), as torch's graph recipe prescribes
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(3):
                self._iteration()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self._reset()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(self.iters):
                self._iteration()
        self._graph = g

    def track(self, src_feats: torch.Tensor, last_match_pts: torch.Tensor) -> Dict[str, torch.Tensor]:
        """src_feats (I*P, C), last_match_pts (I, P, 3) -> {'match_pts' (I,P,3), 't', 'log_r', 'loss'} after `iters`
        iterations from the identity pose (the reference restarts from zero parameters every frame, fusion.py:1633-1637)."""
        import contextlib
        with (torch.cuda.device(self.dev) if self.dev.type == 'cuda' else contextlib.nullcontext()):
            self.src_feats.copy_(src_feats.reshape(self.src_feats.shape))
            self.last_pts.copy_(last_match_pts.reshape(self.last_pts.shape))
            if self.use_graph:
                if self._graph is None:
                    self._capture()
                self._reset()
                self._graph.replay()
            else:
                if not self.opt.state:
                    self._iteration()                      # creates the optimiser state
                self._reset()
                for _ in range(self.iters):
                    self._iteration()
        return {'match_pts': self.curr_pts.reshape(self.I, self.P, 3), 't': self.t.detach(), 'log_r': self.log_r.detach(),
                'loss': self.loss}


class FusedRigidTracker:
    """RigidTracker with every torch op of the iteration done by the native library, the whole `iters`-iteration loop one
    CUDA graph.  Two forms of the iteration:
      * single_launch (default where d3f_track_step supports the observation: V <= 4, float32 map, C % 4 == 0, C <= 1024):
        ONE launch per Adam iteration — one CTA per point computes the point from the pose parameters, queries the field,
        forms the loss gradient and chains it back to the point with the texels held in registers; the last CTA of an
        instance takes the Adam step (csrc/d3f_track.cuh);
      * four launches: field query (d3f_eval) + loss gradient (d3f_track_loss_grad) + field backward (d3f_eval_backward) +
        pose update (d3f_track_update).
    Same loop, loss, pose parametrisation and Adam arithmetic as RigidTracker / the reference (fusion.py:1633-1665);
    agreement of both forms with the torch-autograd version is checked in tests/test_tracking.py.  The observation (pose, K, depth, the descriptor map) is captured by address: update the
    tensors of fusion.curr_obs_torch in place between frames, or call invalidate()."""

    def __init__(self, fusion, num_instance: int, rand_ptcl_num: int, feat_dim: int, iters: int = 100, lr: float = 0.01,
                 reg_w: float = 1.0, dist_w: float = 100.0, graph: bool = True, name: str = 'dino_feats',
                 single_launch: Optional[bool] = None):
        from . import _native
        self._n = _native
        self.fusion, self.iters, self.lr, self.reg_w, self.dist_w, self.name = fusion, iters, lr, reg_w, dist_w, name
        self.use_graph = graph
        self.single_launch = single_launch          # None: decided from the observation at the first track()
        self.launches_per_iteration = None
        dev = torch.device(fusion.device)
        self.dev = dev
        self.I, self.P, self.C = num_instance, rand_ptcl_num, feat_dim
        n = num_instance * rand_ptcl_num
        z = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
        self.t = [z(num_instance, 3), z(num_instance, 3)]          # ping-pong parameter buffers
        self.r = [z(num_instance, 3), z(num_instance, 3)]
        self.m_t, self.v_t, self.m_r, self.v_r = (z(num_instance, 3) for _ in range(4))
        self.last_pts = z(num_instance, rand_ptcl_num, 3)
        self.src_feats = z(n, feat_dim)
        self.pts, self.grad_pts = z(n, 3), z(n, 3)
        self.feat, self.g_feat = z(n, feat_dim), z(n, feat_dim)
        self.dist, self.g_dist, self.loss_terms = z(n), z(n), z(n)
        self.valid = torch.zeros(n, dtype=torch.bool, device=dev)
        self.arrivals = torch.zeros(num_instance, dtype=torch.int32, device=dev)   # d3f_track_step leaves them zero
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    def invalidate(self):
        self._graph = None
        self.arrivals.zero_()

    def _enqueue(self):
        """The whole loop on torch's current stream: `iters` launches (single_launch) or 1 + 4 * iters."""
        N, f = self._n, self.fusion
        V, H, W, pose_p, K_p, depth_p = f._obs_ptrs()
        kt, _vol = f._key_tuple(self.name, V)
        n = self.I * self.P
        st = torch.cuda.current_stream(self.dev).cuda_stream
        flags, mu = f._flags(False), float(f.mu)
        common = dict(m_t=self.m_t.data_ptr(), v_t=self.v_t.data_ptr(), m_r=self.m_r.data_ptr(), v_r=self.v_r.data_ptr(),
                      last_pts=self.last_pts.data_ptr(), n_inst=self.I, n_pts=self.P, lr=self.lr, beta1=0.9, beta2=0.999,
                      eps=1e-8, reg_w=self.reg_w)
        single = self.single_launch
        supported = N.track_step_supported(V, H, W, pose_p, K_p, depth_p, kt)
        if single is None:
            single = supported
        elif single and not supported:
            raise ValueError('single_launch=True: d3f_track_step does not support this observation / descriptor map '
                             '(needs V <= 4, float32, C % 4 == 0, C <= 1024, aligned)')
        self.launches_per_iteration = 1 if single else 4
        if single:
            for k in range(1, self.iters + 1):
                a, b = (k - 1) % 2, k % 2
                N.track_step(V, H, W, pose_p, K_p, depth_p, kt, self.src_feats.data_ptr(), self.dist_w,
                             self.grad_pts.data_ptr(), self.arrivals.data_ptr(), self.loss_terms.data_ptr(), flags, mu, st,
                             t_in=self.t[a].data_ptr(), r_in=self.r[a].data_ptr(), t_out=self.t[b].data_ptr(),
                             r_out=self.r[b].data_ptr(), grad_pts=None,
                             pts=self.pts.data_ptr() if k == self.iters else None,    # the last forward's points
                             step=float(k), **common)
            return
        N.track_update(st, t_in=self.t[0].data_ptr(), r_in=self.r[0].data_ptr(), t_out=None, r_out=None, grad_pts=None,
                       pts=self.pts.data_ptr(), step=0.0, **common)
        for k in range(1, self.iters + 1):
            a, b = (k - 1) % 2, k % 2
            N.eval_device(V, H, W, pose_p, K_p, depth_p, self.pts.data_ptr(), n, [kt], self.dist.data_ptr(),
                          self.valid.data_ptr(), [self.feat.data_ptr()], None, flags, mu, st)
            N.track_loss_grad(self.feat.data_ptr(), self.src_feats.data_ptr(), self.dist.data_ptr(), self.valid.data_ptr(),
                              n, self.C, self.dist_w, self.g_feat.data_ptr(), self.g_dist.data_ptr(),
                              self.loss_terms.data_ptr(), st)
            N.eval_backward(V, H, W, pose_p, K_p, depth_p, self.pts.data_ptr(), n, [kt], [self.g_feat.data_ptr()],
                            self.g_dist.data_ptr(), self.grad_pts.data_ptr(), flags, mu, st)
            N.track_update(st, t_in=self.t[a].data_ptr(), r_in=self.r[a].data_ptr(), t_out=self.t[b].data_ptr(),
                           r_out=self.r[b].data_ptr(), grad_pts=self.grad_pts.data_ptr(),
                           pts=self.pts.data_ptr() if k < self.iters else None,     # keep the last forward's points
                           step=float(k), **common)

    def _reset(self):
        for x in (self.t[0], self.r[0], self.m_t, self.v_t, self.m_r, self.v_r):
            x.zero_()

    def track(self, src_feats: torch.Tensor, last_match_pts: torch.Tensor) -> Dict[str, torch.Tensor]:
        with torch.cuda.device(self.dev):
            self.src_feats.copy_(src_feats.reshape(self.src_feats.shape))
            self.last_pts.copy_(last_match_pts.reshape(self.last_pts.shape))
            self._reset()
            if self.use_graph:
                if self._graph is None:
                    s = torch.cuda.Stream(self.dev)                     # warm-up outside the capture
                    s.wait_stream(torch.cuda.current_stream(self.dev))
                    with torch.cuda.stream(s):
                        self._enqueue()
                    torch.cuda.current_stream(self.dev).wait_stream(s)
                    self._reset()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._enqueue()
                    self._graph = g
                self._graph.replay()
            else:
                self._enqueue()
            fin = self.iters % 2
            prev = (self.iters - 1) % 2                                  # parameters the last forward was computed with
            data = self.loss_terms.sum()
            loss = data + self.reg_w * (torch.norm(self.t[prev]) + torch.norm(self.r[prev]))
        return {'match_pts': self.pts.reshape(self.I, self.P, 3), 't': self.t[fin], 'log_r': self.r[fin], 'loss': loss,
                'data_loss': data}
