"""Optimisation loop on the device: fit triangles to a target image.

The usage pattern of Teg's applications (README.md:3; ``reverse_deriv(..., out_deriv_vals=Tup(Var('dL')))``,
teg/derivs/reverse_deriv.py:211-272): render, form the upstream per-pixel gradient ``dL = image - target`` of the loss
``1/2 * sum_p (image(p) - target(p))^2``, pull it back through the reverse-derivative program to the shared parameters
and take an optimiser step.  In the reference every one of those evaluations is a Python loop over pixels with one
process per pixel (tests/plot.py:8-29, teg/eval/c_eval.py:142-157).

Here one iteration is

    forward   (grouped launches, teg_b200/raster.py)          image band           on the device
    dL        image - target (elementwise)                    upstream gradient    on the device
    loss      1/2 sum dL^2 -> losses[it]                      a device scalar, never read inside the loop
    backward  one grouped launch, per-triangle sums in-kernel; all-reduced over NVLink peer memory when N > 1
    Adam      moments, bias correction, parameter update      [K, 7] device tensors
    update    the new parameters overwrite the device arrays the kernels read

with **no host synchronisation and no host round trip**: parameters, moments, step counter, gradients and losses live in
device memory, and the whole iteration can be replayed as ONE captured CUDA graph (``capture()``) because nothing in it
depends on host values.  Read ``losses`` / ``params`` after the loop.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .raster import MultiTriangleRasterizer


class AdamFit:
    """Adam on the vertices (and optionally the colours) of the triangles of a ``MultiTriangleRasterizer``.

    rasterizer: configured with ``set_scene`` (windows fixed: the optimisation must stay inside them).
    verts0 [K, 6], col0 [K]: start values (host or device).  target: image of this rank's rows, device tensor
    [rows, res].  Multi-GPU: pass ``peers`` (a ``parallel.PeerGroup`` with max_values >= 7 K) so that ``backward``
    delivers the all-reduced gradient itself; without it the caller's process group is used (not capturable)."""

    def __init__(self, rasterizer: MultiTriangleRasterizer, verts0, col0, target, lr: float = 4e-3,
                 betas=(0.9, 0.999), eps: float = 1e-12, fit_color: bool = False, max_iters: int = 4096, peers=None):
        import torch
        self.torch = torch
        self.r = rasterizer
        dev, tdt = rasterizer.dev, rasterizer.tdt
        self.K = rasterizer.K
        p0 = np.concatenate([np.asarray(verts0, dtype=np.float64).reshape(self.K, 6),
                             np.asarray(col0, dtype=np.float64).reshape(self.K, 1)], axis=1) \
            if not torch.is_tensor(verts0) else None
        # parameters and Adam state in float64 on the device (7 K numbers); the kernels read working-precision copies
        self.p = (torch.as_tensor(p0, device=dev) if p0 is not None else
                  torch.cat([verts0.double().reshape(self.K, 6), col0.double().reshape(self.K, 1)], dim=1).to(dev))
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.t = torch.zeros((), dtype=torch.float64, device=dev)          # step counter (device: graph-replayable)
        self.lr, self.b1, self.b2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        self.mask = torch.ones(7, dtype=torch.float64, device=dev)
        if not fit_color:
            self.mask[6] = 0.0
        self.target = target
        self.img = torch.empty((1, rasterizer.n_rows, rasterizer.res), dtype=tdt, device=dev)
        self.dL = torch.empty((rasterizer.n_rows, rasterizer.res), dtype=tdt, device=dev)
        self.g = torch.zeros((self.K, 7), dtype=tdt, device=dev)
        self.losses = torch.zeros(max_iters, dtype=torch.float64, device=dev)
        self.it = torch.zeros((), dtype=torch.int64, device=dev)
        self._slot = torch.zeros(1, dtype=torch.int64, device=dev)
        # pixel values are integrals over a pixel of area 1 / (rows_total * res): rescale so that lr is resolution free
        self.scale = float(rasterizer.grid_res[0] * rasterizer.grid_res[1]) / float(rasterizer.bands)
        self.peers = peers
        if peers is not None:
            rasterizer.set_peers(peers, diff=True)
        self._w = torch.empty((self.K, 7), dtype=tdt, device=dev)
        self._push_params()
        self.graph = None

    def _push_params(self):
        self._w.copy_(self.p)                                   # float64 -> working precision, on the device
        self.r.update_params(self._w[:, :6], self._w[:, 6].contiguous())

    def iteration(self, stream: Optional[int] = None):
        """One iteration, enqueued on ``stream`` (default: torch's current stream).  No synchronisation."""
        torch = self.torch
        sp = torch.cuda.current_stream(self.r.dev).cuda_stream if stream is None else stream
        self.r.forward(out=self.img, stream=sp)
        torch.sub(self.img[0], self.target, out=self.dL)
        loss = 0.5 * self.scale * (self.dL.double() ** 2).sum()
        self._slot.copy_(self.it.reshape(1))
        self.losses.index_copy_(0, self._slot, loss.reshape(1))
        # (the kernel forms image - target itself from the two images; dL above only feeds the loss)
        self.r.backward((self.img[0], self.target), out=self.g, stream=sp)
        if self.peers is None and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(self.g)
        g = self.g.double() * self.scale * self.mask
        self.t.add_(1.0)
        self.it.add_(1)
        self.m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
        self.v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
        mh = self.m / (1.0 - torch.pow(torch.full_like(self.t, self.b1), self.t))
        vh = self.v / (1.0 - torch.pow(torch.full_like(self.t, self.b2), self.t))
        self.p.sub_(self.lr * mh / (vh.sqrt() + self.eps))
        self._push_params()

    def capture(self):
        """Capture one iteration as a CUDA graph; ``run`` then replays it (one submission per iteration)."""
        self.graph = self.r.capture(lambda sp: self.iteration(sp))
        return self.graph

    def run(self, iters: int):
        """``iters`` iterations back to back; returns nothing and waits for nothing (read results afterwards)."""
        for _ in range(iters):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.iteration()

    # ---- results (these synchronise: call them after the loop) -----------------------------------------------------
    def params(self):
        """(verts [K, 6], col [K]) as float64 numpy arrays."""
        p = self.p.cpu().numpy()
        return p[:, :6].copy(), p[:, 6].copy()

    def loss_history(self):
        """losses of the iterations run so far, summed over ranks when a process group is initialised"""
        torch = self.torch
        n = int(self.it.item())
        ls = self.losses[:n].clone()
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(ls)
        return ls.cpu().numpy()


def fit_triangle(res: int = 256, iters: int = 150, lr: float = 4e-3, device: int = 0, start=None, target_verts=None,
                 graph: bool = True, primal=None, grad=None, peers=None):
    """Fit one triangle's vertices to the image of a target triangle (the config-2 scene).  Returns
    (loss history, fitted vertices [6], target vertices [6], the AdamFit).  Under torchrun the rows are sharded."""
    import os
    import torch
    import torch.distributed as dist
    from .parallel import shard_rows
    from .tegp import Program
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    pdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'programs')
    primal = Program.load(os.path.join(pdir, 'tri_color_primal.tegp')) if primal is None else primal
    grad = Program.load(os.path.join(pdir, 'tri_color_grad.tegp')) if grad is None else grad
    target_verts = np.array([0.15, 0.2, 0.85, 0.3, 0.4, 0.9]) if target_verts is None else np.asarray(target_verts, dtype=np.float64)
    start = np.array([0.25, 0.25, 0.75, 0.35, 0.45, 0.8]) if start is None else np.asarray(start, dtype=np.float64)
    rows = shard_rows(res, world, rank)
    r = MultiTriangleRasterizer(primal, grad, res, device=device, rows=rows)
    window = np.array([[0, res, 0, res]], dtype=np.int32)           # one group whose window is the whole image
    r.set_scene(target_verts.reshape(1, 6), np.ones(1), window, [(0, 1)], res, res)
    target = r.forward()[0].clone()
    fit = AdamFit(r, start.reshape(1, 6), np.ones(1), target, lr=lr, max_iters=max(iters, 4096), peers=peers)
    if graph and (world == 1 or peers is not None):
        fit.capture()
        # capturing ran one iteration eagerly (allocations): start over from the same state
        fit.p.copy_(torch.as_tensor(np.concatenate([start, [1.0]]).reshape(1, 7), device=r.dev))
        fit.m.zero_(); fit.v.zero_(); fit.t.zero_(); fit.it.zero_(); fit.losses.zero_()
        fit._push_params()
    fit.run(iters)
    verts, _ = fit.params()
    return fit.loss_history(), verts[0], target_verts, fit
