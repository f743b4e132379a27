"""Multi-triangle rasterization driver.

The reference renders by looping over pixels in Python and evaluating ONE program per pixel
(tests/plot.py:8-29); a scene of K triangles would be K such loops.  Here the same single-triangle
programs (value, and reverse derivative with an upstream per-pixel ``dL``) are evaluated for every
(pixel, triangle) pair inside each triangle's bounding window, as *grouped* launches:

  forward    image[p] = sum_k value_k(p)        one launch per window-disjoint class of triangles, adding
                                                into the image in a fixed class order (no atomics)
  backward   grad[k, :] = sum_p d value_k(p) . dL(p)     ONE launch over all triangles; per-triangle sums are
                                                reduced in-kernel (fixed order) -> [K, 7]
  multi-GPU  every rank owns a band of pixel rows; windows are clipped to the band; the [K, 7] gradient is
             the one quantity exchanged (a single all-reduce).

Programs are ordinary reduced Teg programs (``tri_color_primal`` / ``tri_color_grad`` in
tests/golden/programs, produced by the reference frontend): nothing about triangles is hard-wired into
the kernels; the driver only knows which arguments are the pixel box, which are per-group parameters and
which is the upstream gradient image.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from .eval import CUDA_EvalMode
from .tegp import Program
from .workloads import base_name

BOX = ('x_lb', 'x_ub', 'y_lb', 'y_ub')
PARAMS = ('px0', 'py0', 'px1', 'py1', 'px2', 'py2', 'col')


class MultiTriangleRasterizer:
    def __init__(self, primal: Program, grad: Program, res: int, bands: int = 1, rank: int = 0,
                 num_samples: int = 16, device: int = 0, float_width: str = 'float', cull: bool = True):
        import torch
        self.torch = torch
        self.res = int(res)
        self.bands = int(bands)
        self.rank = int(rank)
        self.device = device
        self.dev = torch.device('cuda', device)
        self.tdt = torch.float32 if float_width == 'float' else torch.float64
        self.m_primal = CUDA_EvalMode(primal, num_samples=num_samples, device=device, float_width=float_width,
                                      cull=cull)
        self.m_grad = CUDA_EvalMode(grad, num_samples=num_samples, device=device, float_width=float_width, cull=cull)
        self.p_lab = {base_name(lab): lab for lab in primal.args}
        self.g_lab = {base_name(lab): lab for lab in grad.args}
        self.bounds = ((0.0, float(bands)), (0.0, 1.0))
        self.grid_res = (self.res * self.bands, self.res)
        self.rows = (self.res * self.rank, self.res * (self.rank + 1))
        self.K = 0
        self.peers = None

    def set_peers(self, peers) -> None:
        """peers: a ``parallel.PeerGroup`` with max_values >= 7 K.  ``backward`` then returns the gradient summed over
        all ranks: the last per-triangle column-sum kernel exchanges the [K, 7] sums over NVLink peer memory itself
        (no separate collective).  Call after ``set_scene`` on every rank; every rank must call ``backward`` the same
        number of times."""
        assert self.K > 0, 'set_scene first'
        _, prog = self.m_grad.grouped_program([self.g_lab[b] for b in BOX], [self.g_lab['dL']], reduce=True)
        peers.attach_groups(prog, self.K)
        self.peers = peers

    def set_scene(self, verts, col, windows, classes: Sequence[Tuple[int, int]], max_rows: int, max_cols: int,
                  active: Optional[Tuple[int, int]] = None):
        """verts [K, 6], col [K] (host or device), windows int32 [K, 4], classes: disjoint-window group ranges.
        active: the contiguous group range whose windows can touch this rank's band (default: all groups);
        groups outside it are never launched here and get zero gradient rows."""
        torch = self.torch
        verts = torch.as_tensor(np.asarray(verts) if not torch.is_tensor(verts) else verts)
        col = torch.as_tensor(np.asarray(col) if not torch.is_tensor(col) else col)
        verts = verts.to(self.dev, self.tdt)
        self.K = int(verts.shape[0])
        # structure of arrays: one contiguous [K] tensor per parameter
        self.params: Dict[str, object] = {nm: verts[:, i].contiguous() for i, nm in enumerate(PARAMS[:6])}
        self.params['col'] = col.to(self.dev, self.tdt).contiguous()
        self.windows = torch.as_tensor(np.asarray(windows) if not torch.is_tensor(windows) else windows)
        self.windows = self.windows.to(self.dev, torch.int32).contiguous()
        self.active = (0, self.K) if active is None else (int(active[0]), int(active[1]))
        self.classes = [(max(a, self.active[0]), min(b, self.active[1])) for a, b in classes
                        if max(a, self.active[0]) < min(b, self.active[1])]
        self.max_rows, self.max_cols = int(max_rows), int(max_cols)

    def update_params(self, verts, col):
        """New parameter values for the same windows (an optimisation step that stays inside them)."""
        for i, nm in enumerate(PARAMS[:6]):
            self.params[nm].copy_(verts[:, i])
        self.params['col'].copy_(col)

    def forward(self, out=None, stream: int = 0):
        """image band [rows, res]: sum over triangles of the per-pixel value integral."""
        torch = self.torch
        if out is None:
            out = torch.empty((1, self.res, self.res), dtype=self.tdt, device=self.dev)
        out.zero_()
        box = [self.p_lab[b] for b in BOX]
        gb = {self.p_lab[nm]: self.params[nm] for nm in PARAMS}
        for g0, g1 in self.classes:
            self.m_primal.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows,
                                        self.max_cols, rows=self.rows, out=out, accumulate=True,
                                        group_range=(g0, g1), stream=stream)
        return out

    def backward(self, dL, out=None, stream: int = 0):
        """grad [K, 7]: d(sum_p dL(p) * image(p)) / d(px0, py0, px1, py1, px2, py2, col) of every triangle,
        summed over this rank's band (all-reduce it across ranks)."""
        torch = self.torch
        if out is None:
            out = torch.empty((self.K, 7), dtype=self.tdt, device=self.dev)
        box = [self.g_lab[b] for b in BOX]
        gb = {self.g_lab[nm]: self.params[nm] for nm in PARAMS}
        a0, a1 = self.active
        if self.peers is not None:
            # all K rows are written by the fused exchange (groups outside the active range contribute zero)
            self.m_grad.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows, self.max_cols,
                                      pixel_bindings={self.g_lab['dL']: dL}, rows=self.rows, out=out, reduce=True,
                                      group_range=(a0, a1), stream=stream)
            return out
        if (a0, a1) != (0, self.K):
            out.zero_()
        self.m_grad.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows, self.max_cols,
                                  pixel_bindings={self.g_lab['dL']: dL}, rows=self.rows, out=out[a0:a1], reduce=True,
                                  group_range=(a0, a1), stream=stream)
        return out

    def capture(self, fn):
        """CUDA graph of ``fn(stream_ptr)`` (any mix of ``forward`` / ``backward`` / torch ops on fixed tensors).

        Grouped launches read the triangles' parameters from device arrays (``update_params`` overwrites them in
        place), never from launch arguments, so a captured step stays valid while the parameters move: replaying it
        is one submission instead of ~30 launches.  ``fn`` runs once outside the capture first (allocations)."""
        torch = self.torch
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            fn(side.cuda_stream)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn(torch.cuda.current_stream(self.dev).cuda_stream)
        return graph
