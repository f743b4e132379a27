"""Multi-triangle rasterization driver.

The reference renders by looping over pixels in Python and evaluating ONE program per pixel
(tests/plot.py:8-29); a scene of K triangles would be K such loops.  Here the same single-triangle
programs (value, and reverse derivative with an upstream per-pixel ``dL``) are evaluated for every
(pixel, triangle) pair inside each triangle's bounding window, as *grouped* launches:

  forward    image[p] = sum_k value_k(p)        ONE gather launch over the image: every block walks the triangles
                                                of its bin in ascending order (no atomics, no fill; `gather=False`:
                                                one accumulating launch per window-disjoint class, the same bits)
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


def select_groups(windows, classes: Sequence[Tuple[int, int]], rows: Tuple[int, int]):
    """The groups a rank holds: those whose pixel windows reach into its rows [rows[0], rows[1]).

    windows: int array [K, 4] = (row_begin, row_end, col_begin, col_end); classes: ranges of global group indices with
    pairwise disjoint windows.  Returns (ids, local_classes, local_of): the held groups' global indices class by class
    (so every class is a contiguous LOCAL range, possibly empty and then dropped), and the inverse map [K] (-1 = not
    held) the fused all-reduce uses to find a global group's local column."""
    w = np.asarray(windows).astype(np.int64)
    r0, r1 = int(rows[0]), int(rows[1])
    hit = (w[:, 0] < r1) & (w[:, 1] > r0) & (w[:, 2] < w[:, 3]) & (w[:, 0] < w[:, 1])
    ids, local_classes = [], []
    for a, b in classes:
        mine = [k for k in range(a, b) if hit[k]]
        if mine:
            local_classes.append((len(ids), len(ids) + len(mine)))
            ids.extend(mine)
    local_of = np.full(len(w), -1, dtype=np.int32)
    if ids:
        local_of[np.asarray(ids, dtype=np.int64)] = np.arange(len(ids), dtype=np.int32)
    return ids, local_classes, local_of


def group_bins(windows, rows: Tuple[int, int], res_y: int, block_cols: int, tile_rows: int):
    """Bins of a gather launch (``tegb_render_groups_gather``): the band ``rows`` in blocks of ``block_cols`` columns x
    ``tile_rows`` rows (row-major: bin = block_row * ceil(res_y / block_cols) + block_col); every block lists, ASCENDING,
    the groups whose windows reach into it.  Returns (bin_begin int32 [n_bins + 1], bin_groups int32 [n_entries])
    (``bin_groups`` has one padding element when empty, so that it always has an address)."""
    w = np.asarray(windows).astype(np.int64).reshape(-1, 4)
    r0, r1 = int(rows[0]), int(rows[1])
    nbx = -(-int(res_y) // block_cols)
    nby = -(-(r1 - r0) // tile_rows)
    bins, grps = [], []
    for k in range(len(w)):
        a, b = max(int(w[k, 0]), r0), min(int(w[k, 1]), r1)
        c, d = max(int(w[k, 2]), 0), min(int(w[k, 3]), int(res_y))
        if a >= b or c >= d:
            continue
        by = np.arange((a - r0) // tile_rows, (b - 1 - r0) // tile_rows + 1, dtype=np.int64)
        bx = np.arange(c // block_cols, (d - 1) // block_cols + 1, dtype=np.int64)
        ids = (by[:, None] * nbx + bx[None, :]).ravel()
        bins.append(ids)
        grps.append(np.full(ids.shape, k, dtype=np.int64))
    begin = np.zeros(nbx * nby + 1, dtype=np.int32)
    if not bins:
        return begin, np.zeros(1, dtype=np.int32)
    bins, grps = np.concatenate(bins), np.concatenate(grps)
    order = np.lexsort((grps, bins))                    # by bin, then ascending group
    begin[1:] = np.cumsum(np.bincount(bins, minlength=nbx * nby))
    return begin, grps[order].astype(np.int32)


class MultiTriangleRasterizer:
    def __init__(self, primal: Program, grad: Program, res: int, bands: int = 1, rank: int = 0,
                 num_samples: int = 16, device: int = 0, float_width: str = 'float', cull: bool = True,
                 min_blocks=(4, 4), rows: Optional[Tuple[int, int]] = None, tile_rows=(32, 32), gather: bool = True,
                 **codegen_opts):
        import torch
        self.torch = torch
        self.res = int(res)
        self.bands = int(bands)
        self.rank = int(rank)
        self.device = device
        self.dev = torch.device('cuda', device)
        self.tdt = torch.float32 if float_width == 'float' else torch.float64
        # tile_rows = (value kernel, gradient kernel): pixel rows per block / classification tile (measured on config 5:
        # 32 / 32 rows with the gather forward, 16 / 32 with one launch per class; the gradient kernel pays its
        # per-block sum once per twice as many pixels)
        tr = (tile_rows, tile_rows) if isinstance(tile_rows, int) else tuple(tile_rows)
        # min_blocks = (value kernel, gradient kernel): resident blocks per SM the grouped kernels are compiled for
        # (__launch_bounds__; None = the compiler's choice)
        self.m_primal = CUDA_EvalMode(primal, num_samples=num_samples, device=device, float_width=float_width,
                                      cull=cull, min_blocks=min_blocks[0], tile_rows=tr[0], **codegen_opts)
        self.m_grad = CUDA_EvalMode(grad, num_samples=num_samples, device=device, float_width=float_width, cull=cull,
                                    min_blocks=min_blocks[1], tile_rows=tr[1], **codegen_opts)
        self.p_lab = {base_name(lab): lab for lab in primal.args}
        self.g_lab = {base_name(lab): lab for lab in grad.args}
        self.bounds = ((0.0, float(bands)), (0.0, 1.0))
        self.grid_res = (self.res * self.bands, self.res)
        # rows: the pixel rows this rank owns.  Default: band `rank` of `bands` (weak scaling: one res x res band per
        # rank).  Any other split of the (res * bands) rows works too, e.g. a slice of ONE band's rows (strong scaling):
        # windows are clipped to the rows, groups whose windows miss them cost an empty block each.
        self.rows = (self.res * self.rank, self.res * (self.rank + 1)) if rows is None else (int(rows[0]), int(rows[1]))
        assert 0 <= self.rows[0] <= self.rows[1] <= self.res * self.bands
        self.n_rows = self.rows[1] - self.rows[0]
        self.K = 0
        self.peers = None
        # forward pass: ONE launch over the image, every block walking the triangles of its bin (False: one accumulating
        # launch per class of window-disjoint triangles; the same image bit for bit)
        self.gather = bool(gather)
        self.tile_rows = tr

    def set_peers(self, peers, diff: bool = False) -> None:
        """peers: a ``parallel.PeerGroup`` with max_values >= 7 K.  ``backward`` then returns the gradient summed over
        all ranks: the last per-triangle column-sum kernel exchanges the [K, 7] sums over NVLink peer memory itself
        (no separate collective).  Call after ``set_scene`` on every rank; every rank must call ``backward`` the same
        number of times."""
        assert self.K > 0, 'set_scene first'
        # (one compiled variant per way dL is bound: attach both the image form and the (image, target) form on demand)
        _, prog = self.m_grad.grouped_program([self.g_lab[b] for b in BOX], [self.g_lab['dL']], reduce=True,
                                              pixel_diff=[self.g_lab['dL']] if diff else [])
        peers.attach_groups(prog, self.K, self.local_of)
        self.peers = peers

    def set_scene(self, verts, col, windows, classes: Sequence[Tuple[int, int]], max_rows: int, max_cols: int,
                  active: Optional[Tuple[int, int]] = None):
        """verts [K, 6], col [K] (host or device), windows int32 [K, 4] (host), classes: group ranges with pairwise
        disjoint windows (the forward pass adds one class per launch, in this order).
        Only the groups whose windows reach into this rank's rows are kept on the device (``self.ids``: their global
        indices, class by class, so every class is a contiguous LOCAL range): a rank never launches blocks for triangles
        it cannot see.  ``active`` (a global group range) is accepted for compatibility; the windows decide."""
        torch = self.torch
        verts = torch.as_tensor(np.asarray(verts) if not torch.is_tensor(verts) else verts).to(self.dev, self.tdt)
        col = torch.as_tensor(np.asarray(col) if not torch.is_tensor(col) else col).to(self.dev, self.tdt)
        w = np.asarray(windows.cpu() if torch.is_tensor(windows) else windows).astype(np.int64)
        self.K = int(verts.shape[0])
        r0, r1 = self.rows
        ids, local_classes, local_of = select_groups(w, classes, self.rows)
        self.K_local = len(ids)
        self.identity = ids == list(range(self.K))        # every group is held, in the global order
        self.ids = torch.as_tensor(np.asarray(ids, dtype=np.int64), device=self.dev)
        self.local_of = torch.as_tensor(local_of, device=self.dev)
        self.classes = local_classes
        # structure of arrays: one contiguous [K_local] tensor per parameter
        self.params: Dict[str, object] = {nm: verts[self.ids, i].contiguous() for i, nm in enumerate(PARAMS[:6])}
        self.params['col'] = col[self.ids].contiguous()
        wl = w[np.asarray(ids, dtype=np.int64)] if ids else np.zeros((0, 4), dtype=np.int64)
        if not ids:
            # nothing reaches into this rank's rows: placeholders of one element keep every pointer valid; the launches
            # cover zero groups (the rank still takes part in the fused all-reduce)
            self.params = {nm: torch.zeros(1, dtype=self.tdt, device=self.dev) for nm in PARAMS}
        self.windows = torch.as_tensor((wl if ids else np.zeros((1, 4), dtype=np.int64)).astype(np.int32),
                                       device=self.dev).contiguous()
        if self.K_local:
            self.max_rows = int(min(max_rows, (np.minimum(wl[:, 1], r1) - np.maximum(wl[:, 0], r0)).max()))
            self.max_cols = int(min(max_cols, (wl[:, 3] - wl[:, 2]).max()))
        else:
            self.max_rows, self.max_cols = 1, 1
        self.g_local = torch.zeros((max(self.K_local, 1), 7), dtype=self.tdt, device=self.dev)
        self.bins = None
        if self.gather and self.K_local:
            bb, bg = group_bins(wl, self.rows, self.res, self.m_primal.block_size, self.tile_rows[0])
            self.bins = (torch.as_tensor(bb, device=self.dev), torch.as_tensor(bg, device=self.dev))

    def update_params(self, verts, col):
        """New parameter values ([K, 6] / [K] device tensors, global order) for the same windows (an optimisation step
        that stays inside them)."""
        if self.K_local == 0:
            return
        v = verts.index_select(0, self.ids)
        for i, nm in enumerate(PARAMS[:6]):
            self.params[nm].copy_(v[:, i])
        self.params['col'].copy_(col.index_select(0, self.ids))

    def forward(self, out=None, stream: int = 0):
        """image band [rows, res]: sum over triangles of the per-pixel value integral."""
        torch = self.torch
        if out is None:
            out = torch.empty((1, self.n_rows, self.res), dtype=self.tdt, device=self.dev)
        box = [self.p_lab[b] for b in BOX]
        gb = {self.p_lab[nm]: self.params[nm] for nm in PARAMS}
        if self.bins is not None:
            # (no fill pass: every block of the gather launch clears its pixels before it adds to them)
            self.m_primal.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows,
                                        self.max_cols, rows=self.rows, out=out, accumulate=True, stream=stream,
                                        gather=self.bins, clear=True)
            return out
        out.zero_()
        if len(self.classes) > 1:
            # uniform rows + tile classification of ALL held triangles once, then one accumulating launch per class
            self.m_primal.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows,
                                        self.max_cols, rows=self.rows, accumulate=True,
                                        group_range=(0, self.K_local), stream=stream, prepare_only=True)
        for g0, g1 in self.classes:
            self.m_primal.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows,
                                        self.max_cols, rows=self.rows, out=out, accumulate=True,
                                        group_range=(g0, g1), stream=stream, prepared=len(self.classes) > 1)
        return out

    def backward(self, dL, out=None, stream: int = 0):
        """grad [K, 7]: d(sum_p dL(p) * image(p)) / d(px0, py0, px1, py1, px2, py2, col) of every triangle,
        summed over this rank's band (all-reduce it across ranks).  dL: the upstream gradient image of this rank's
        rows, or a pair ``(image, target)`` for dL = image - target formed inside the kernel."""
        torch = self.torch
        if out is None:
            out = torch.empty((self.K, 7), dtype=self.tdt, device=self.dev)
        box = [self.g_lab[b] for b in BOX]
        gb = {self.g_lab[nm]: self.params[nm] for nm in PARAMS}
        if self.peers is not None:
            # all K rows are written by the fused exchange (groups this rank does not hold contribute zero)
            self.m_grad.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows, self.max_cols,
                                      pixel_bindings={self.g_lab['dL']: dL}, rows=self.rows, out=out, reduce=True,
                                      group_range=(0, self.K_local), stream=stream)
            return out
        if self.identity:
            self.m_grad.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows, self.max_cols,
                                      pixel_bindings={self.g_lab['dL']: dL}, rows=self.rows, out=out, reduce=True,
                                      group_range=(0, self.K), stream=stream)
            return out
        out.zero_()
        if self.K_local:
            self.m_grad.render_groups(box, self.grid_res, self.bounds, gb, self.windows, self.max_rows, self.max_cols,
                                      pixel_bindings={self.g_lab['dL']: dL}, rows=self.rows, out=self.g_local,
                                      reduce=True, group_range=(0, self.K_local), stream=stream)
            out.index_copy_(0, self.ids, self.g_local[:self.K_local])
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
