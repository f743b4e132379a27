"""``CUDA`` evaluation backend behind Teg's plugin interface.

Mirrors, for this path only:
  EvalMode                      /root/reference/teg/eval/eval_mode.py:1-7
  C_EvalMode(expr, num_samples) /root/reference/teg/eval/c_eval.py:22-157   (ctor options, eval(bindings))
  BACKENDS / evaluate           /root/reference/teg/eval/backend.py:6-41    (registry + per-expression mode cache)
  render_image                  /root/reference/tests/plot.py:8-29          (per-pixel driver)

Same argument convention: ``bindings`` is keyed by ``Var`` objects and resolved through the labels
``f'{name}_{uid}'`` (c_eval.py:135); unbound arguments fall back to ``Var.value`` (c_eval.py:136); any
argument still ``None`` is an ``AssertionError`` (c_eval.py:139-140).  A scalar-only call returns a
Python ``float`` (or ``list`` of floats for ``Tup`` programs), like c_eval.py:149-157.

Extension (the point of the backend): a binding may be a length-n array / torch tensor.  All array
bindings must have the same length; scalars broadcast; the result is then ``ndarray[n]`` (or
``[k, n]`` for Tup programs) -- torch CUDA tensors in, torch CUDA tensors out, without host copies.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import runtime
from .codegen import KernelSource, generate
from .dag import Dag, build_dag
from .schedule import make_schedule
from .tegp import Program, from_teg, label_of


class EvalMode:
    """reference teg/eval/eval_mode.py:1-7"""

    def __init__(self, name=None):
        self.name = name

    def eval(self, bindings, **kwargs):
        raise NotImplementedError('This backend does not support evaluation')


def _is_torch(x) -> bool:
    return type(x).__module__.split('.')[0] == 'torch'


def _np_dtype(float_width: str):
    return np.float32 if float_width == 'float' else np.float64


class CUDA_EvalMode(EvalMode):
    """Batched sm_100a evaluation of one reduced Teg program.

    expr: a base-language expression (reference ``teg`` objects or ``teg_b200.lang``), a
          FwdDeriv/RevDeriv wrapper (unwrapped like compile.py:182-183), or a ``tegp.Program``.
    num_samples: midpoint samples per integral (reference default 50, c_eval.py:26).
    float_width: 'float' (what C_EvalMode hard-wires, c_eval.py:32) or 'double'
                 (the CLI's ``-f double``, teg/__main__.py:49-52).
    integration_mode: 'rectangular_quadrature' (default, the C backend), 'trapezoidal_quadrature' (the numpy
                 backend's rule) or 'monte_carlo_uniform' (uniform samples from Philox4x32-10 keyed by ``seed``, the
                 instance index, the integral's nesting depth and the sample index: reproducible on any launch
                 geometry; instance index = position in the batch, or the pixel's index in the whole image for
                 ``render``); names from to_c.py:389-393.
    cull: domain culling (cull.py, default on): bit-identical results, far fewer evaluated comparisons.
    """

    def __init__(self, expr, num_samples: int = 50, float_width: str = 'float', device: int = 0,
                 arglist: Sequence = (), name: str = 'teg_method', use_cache: bool = True,
                 block_size: int = 128, team='auto', nvrtc_opts: str = '',
                 integration_mode: str = 'rectangular_quadrature', cull: bool = True, seed: int = 0, **codegen_opts):
        super().__init__(name='CUDA')
        if float_width in ('single', 'f32', np.float32):
            float_width = 'float'
        if float_width in ('f64', np.float64):
            float_width = 'double'
        assert float_width in ('float', 'double'), f'Invalid precision setting {float_width}'
        if isinstance(expr, (list, tuple)):
            # several programs over the same variables, evaluated together by one kernel (tegp.bundle)
            from .tegp import bundle
            expr = bundle([e if isinstance(e, Program) else from_teg(e, arglist, name=f'{name}{k}')
                           for k, e in enumerate(expr)], name=name if name != 'teg_method' else None)
        self.program: Program = expr if isinstance(expr, Program) else from_teg(expr, arglist, name=name)
        self.dag: Dag = build_dag(self.program)
        self.num_samples = int(num_samples)
        self.float_width = float_width
        self.device = device
        self.use_cache = use_cache
        self.block_size = block_size
        # team: threads cooperating on one instance.  1 = thread per instance.  'auto' (default) keeps 1 unless the
        # batch is too small to fill the GPU *and* each instance carries >= 4096 samples; then the outermost
        # sample loop is split over a team (slice of a warp / warp / block): members evaluate consecutive outer
        # samples (their inner integrals are ordinary sequential folds) and every member adds the outer terms in
        # sample order (tegb_team_fold).  Either way every sum is the reference's left fold: bit-exact w.r.t. the
        # C backend.
        assert team == 'auto' or (isinstance(team, int) and team >= 1 and team & (team - 1) == 0 and team <= 1024)
        self.team = team
        # integration_mode: 'rectangular_quadrature' = the C backend's midpoint rule (the only template the
        # reference ships, to_c.py:389-393); 'trapezoidal_quadrature' = the numpy backend's rule
        # (numpy_eval.py:57-68: linspace end points included).
        from .codegen import INTEGRATION_MODES
        assert integration_mode in INTEGRATION_MODES, f'unknown integration mode {integration_mode}'
        self.integration_mode = integration_mode
        if integration_mode != 'rectangular_quadrature':
            codegen_opts = dict(codegen_opts, integration_mode=integration_mode)
            if integration_mode == 'monte_carlo_uniform':
                codegen_opts['mc_seed'] = int(seed)
            assert team in (1, 'auto'), 'teams are defined for rectangular_quadrature only'
            self.team = 1
        self.codegen_opts = codegen_opts
        # cull: per-instance interval evaluation of the select conditions over the sampled box; instances whose
        # discontinuities miss the box run comparison-free specialised nests (bit-identical results, cull.py)
        self.cull = bool(cull) and integration_mode == 'rectangular_quadrature'
        self.tiles = self.cull and bool(codegen_opts.pop('tiles', True))     # tile-level classification for renders
        self.codegen_opts = codegen_opts
        self.sched_dag: Dag = self.dag
        if self.cull:
            from .cull import cull as _cull
            self.sched_dag = _cull(self.dag, self.num_samples)
        self.nvrtc_opts = nvrtc_opts
        self.arglist: List[Tuple[str, Optional[float]]] = [
            (lab, self.program.defaults.get(lab)) for lab in self.program.args]
        self.out_size = len(self.dag.outputs)
        self.out_is_tuple = self.dag.out_is_tuple
        self._variants: Dict[tuple, Tuple[KernelSource, runtime.Program]] = {}
        self._schedules: Dict[tuple, object] = {}

    # ------------------------------------------------------------------ compile
    def schedule(self, array_args: Sequence[int], grid_args: Sequence[int] = (), tiled: bool = False):
        """tiled (plain render launches with culling): the program additionally takes one tri-state per tile
        condition, classified per tile of pixels by ``<name>_tiles`` (cull.tile_cull)."""
        key = tuple(sorted(set(array_args) | set(grid_args)))
        tiled = bool(tiled) and self.cull and len(grid_args) == 4
        if (key, tiled) not in self._schedules:
            dag, arr = self.sched_dag, list(key)
            if tiled:
                from .cull import tile_cull
                dag = tile_cull(self.sched_dag, list(grid_args), self.num_samples,
                                other_array_args=sorted(set(array_args) - set(grid_args)))
                arr += [len(self.program.args) + k for k in range(len(dag.tile_conds))]
            self._schedules[(key, tiled)] = make_schedule(dag, arr)
        return self._schedules[(key, tiled)]

    def reference_schedule(self, array_args: Sequence[int]):
        """Schedule of the program as the reference states it (no culling): what the algorithmic operation
        count (opcount.py) is defined on."""
        return make_schedule(self.dag, sorted(set(array_args)))

    def _n_reduce(self, reduce) -> int:
        """reduce: False / True (all outputs) / int (that many TRAILING outputs, e.g. the derivative part of a
        value + derivative bundle)."""
        k = len(self.dag.outputs)
        nr = k if reduce is True else int(reduce)
        assert 0 <= nr <= k
        return nr

    def pick_team(self, n: int, array_args: Sequence[int]) -> int:
        if self.team != 'auto':
            return int(self.team)
        depth = self.schedule(array_args).loop_depth()
        if depth == 0 or n >= 32768 or float(self.num_samples) ** depth < 4096:
            return 1
        cap = 1
        while cap < self.num_samples and cap < 1024:
            cap *= 2
        team = 1
        while team * 2 <= cap and n * team < 148 * 1024:
            team *= 2
        return team

    def kernel_source(self, array_args: Sequence[int], grid_args: Sequence[int] = (),
                      reduce: bool = False, team: int = 1, vec: int = 1) -> KernelSource:
        sch = self.schedule(array_args, grid_args, tiled=bool(grid_args) and team == 1 and self.tiles)
        block = team if team > 32 else self.block_size
        return generate(sch, self.float_width, self.num_samples, grid_args=list(grid_args),
                        reduce_outputs=reduce, block_size=block, team=team, vec=vec, **self.codegen_opts)

    def pick_vec(self, array_args: Sequence[int], reduce: bool, team: int, ptrs: Sequence[int]) -> int:
        """4 instances per thread (16-byte vector loads/stores) for streaming programs: no sample loop left
        after delta elimination (e.g. the README derivative), plain per-instance arrays, aligned buffers."""
        if reduce or team != 1 or not array_args or any(p & 15 for p in ptrs):
            return 1
        return 4 if self.schedule(array_args).loop_depth() == 0 else 1

    def compiled(self, array_args: Sequence[int], grid_args: Sequence[int] = (), reduce: bool = False,
                 team: int = 1, grouped: bool = False, pixel_args: Sequence[int] = (), accumulate: bool = False,
                 vec: int = 1, private: bool = False, pixel_diff: Sequence[int] = ()):
        """(KernelSource, runtime.Program) of one variant, compiled once per mode.  private: a NEW handle on the same
        module (own uniform table and scratch buffers) for callers that keep it to themselves."""
        reduce = self._n_reduce(reduce)
        key = (tuple(sorted(array_args)), tuple(grid_args), reduce, int(team), bool(grouped),
               tuple(pixel_args), bool(accumulate), int(vec), tuple(sorted(pixel_diff)))
        hit = self._variants.get(key)
        if hit is None:
            if grouped:
                sch = self.schedule(list(grid_args) + list(pixel_args), grid_args, tiled=self.tiles)
                ks = generate(sch, self.float_width, self.num_samples, grid_args=list(grid_args),
                              reduce_outputs=reduce, grouped=True, pixel_args=list(pixel_args),
                              accumulate=accumulate, block_size=self.block_size, pixel_diff=list(pixel_diff),
                              **self.codegen_opts)
            else:
                ks = self.kernel_source(array_args, grid_args, reduce, team, vec)
            mod = runtime.Module(ks.source, opts=self.nvrtc_opts, use_cache=self.use_cache)
            prog = runtime.Program(mod, ks, device=self.device)
            hit = (ks, prog)
            self._variants[key] = hit
        if private:
            return hit[0], runtime.Program(hit[1].module, hit[0], device=self.device)
        return hit

    # --------------------------------------------------------------------- eval
    def _resolve(self, bindings) -> List:
        by_label = {}
        for k, v in (bindings or {}).items():
            by_label[k if isinstance(k, str) else label_of(k)] = v
        live = self.program.live
        args = []
        for lab, default in self.arglist:
            if lab in by_label:
                args.append(by_label[lab])
            elif lab in live:                     # c_eval.py:136: the Var's *current* value
                args.append(live[lab].value)
            else:
                args.append(default)
        assert all(a is not None for a in args), \
            f'No bindings found for {[self.arglist[i][0] for i, a in enumerate(args) if a is None]}'
        return args

    def _resolve_scalars(self, bindings, arg_indices: Sequence[int]) -> List[float]:
        """Launch-wide arguments of a render call, resolved like ``_resolve`` (c_eval.py:135-140): the binding, else
        the Var's *current* value, else the default captured at conversion time; ``AssertionError`` when unbound."""
        by_label = {(k if isinstance(k, str) else label_of(k)): v for k, v in (bindings or {}).items()}
        live = self.program.live
        out = []
        for a in arg_indices:
            lab, default = self.arglist[a]
            v = by_label[lab] if lab in by_label else (live[lab].value if lab in live else default)
            assert v is not None, f'No bindings found for {[lab]}'
            out.append(float(v))
        return out

    def eval(self, bindings={}, reduce: bool = False, **kwargs):
        """reduce=True: return the per-component SUM over the instances (reverse-mode gradients w.r.t.
        shared parameters) instead of per-instance values; the summation happens inside the kernel."""
        args = self._resolve(bindings)
        dt = _np_dtype(self.float_width)
        is_arr = []
        n = None
        any_torch = False
        for a in args:
            if _is_torch(a) and a.dim() > 0:
                is_arr.append(True)
                any_torch = True
                ln = a.numel()
            elif not _is_torch(a) and np.ndim(a) > 0:
                is_arr.append(True)
                ln = int(np.size(a))
            else:
                is_arr.append(False)
                continue
            assert n is None or n == ln, 'all array bindings must have the same length'
            n = ln
        array_args = [i for i, f in enumerate(is_arr) if f]
        team = self.pick_team(1 if n is None else n, array_args)
        if any_torch:
            return self._eval_torch(args, array_args, n, reduce, team)
        # host path: the library stages through its own (256-byte aligned) device buffers
        vec = self.pick_vec(array_args, reduce, team, [])
        ks, prog = self.compiled(array_args, reduce=reduce, team=team, vec=vec)
        scalars = [float(args[a]) for a in ks.scalar_args]
        batched = n is not None
        n = 1 if n is None else n
        ins = [np.ascontiguousarray(np.asarray(args[a], dtype=dt).reshape(-1)) for a in ks.array_args]
        if reduce:
            sums = np.empty(ks.n_outputs, dtype=dt)
            prog.eval_host(scalars, [a.ctypes.data for a in ins], [sums.ctypes.data], n)
            return sums if self.out_is_tuple and ks.n_outputs > 1 else sums[0]
        out = np.empty((ks.n_outputs, n), dtype=dt)
        prog.eval_host(scalars, [a.ctypes.data for a in ins],
                       [out[k].ctypes.data for k in range(ks.n_outputs)], n)
        if not batched:
            vals = [float(v) for v in out[:, 0]]
            return vals if self.out_is_tuple and len(vals) > 1 else vals[0]
        return out if self.out_is_tuple and ks.n_outputs > 1 else out[0]

    def _eval_torch(self, args, array_args, n, reduce, team):
        import torch
        tdt = torch.float32 if self.float_width == 'float' else torch.float64
        dev = torch.device('cuda', self.device)
        ins = {}
        for a in array_args:
            t = args[a]
            if not _is_torch(t):
                t = torch.as_tensor(np.asarray(t), device=dev)
            ins[a] = t.to(device=dev, dtype=tdt).contiguous().view(-1)
        k_out = len(self.dag.outputs)
        out = None if reduce else torch.empty((k_out, n), dtype=tdt, device=dev)
        ptrs = [t.data_ptr() for t in ins.values()] + ([] if reduce else [out[k].data_ptr() for k in range(k_out)])
        vec = self.pick_vec(array_args, reduce, team, ptrs)
        ks, prog = self.compiled(array_args, reduce=reduce, team=team, vec=vec)
        scalars = [float(args[a]) for a in ks.scalar_args]
        ins = [ins[a] for a in ks.array_args]
        stream = torch.cuda.current_stream(dev).cuda_stream
        if reduce:
            sums = torch.empty(ks.n_outputs, dtype=tdt, device=dev)
            prog.launch(scalars, [t.data_ptr() for t in ins], [sums.data_ptr()], n, stream=stream)
            return sums if self.out_is_tuple and ks.n_outputs > 1 else sums[0]
        prog.launch(scalars, [t.data_ptr() for t in ins], [out[k].data_ptr() for k in range(ks.n_outputs)],
                    n, stream=stream)
        return out if self.out_is_tuple and ks.n_outputs > 1 else out[0]

    def grouped_program(self, box_vars, pixel_vars=(), reduce: bool = False, accumulate: bool = False,
                        pixel_diff=()):
        """(KernelSource, runtime.Program) of the grouped render variant ``render_groups`` launches for these box /
        per-pixel arguments (e.g. to attach peer mailboxes to it: ``PeerGroup.attach_groups``)."""
        pos = {lab: i for i, (lab, _) in enumerate(self.arglist)}
        grid_args = [pos[v if isinstance(v, str) else label_of(v)] for v in box_vars]
        pixel_args = sorted(pos[v if isinstance(v, str) else label_of(v)] for v in pixel_vars)
        diff = sorted(pos[v if isinstance(v, str) else label_of(v)] for v in pixel_diff)
        return self.compiled(grid_args, grid_args, reduce=reduce, grouped=True, pixel_args=pixel_args,
                             accumulate=accumulate, pixel_diff=diff)

    def render_groups(self, box_vars, res: Tuple[int, int], bounds, group_bindings, windows, max_rows: int,
                      max_cols: int, pixel_bindings=None, rows: Optional[Tuple[int, int]] = None, out=None,
                      reduce: bool = False, accumulate: bool = False, group_range: Optional[Tuple[int, int]] = None,
                      stream: int = 0, prepared: bool = False, prepare_only: bool = False, gather=None,
                      clear: bool = False):
        """One launch over many *groups* (e.g. triangles), each with its own parameter values and its own
        window of pixels of the tests/plot.py grid.

        group_bindings: {Var | label: torch CUDA tensor [G]} for every non-box, non-pixel argument.
        windows: int32 CUDA tensor [G, 4] = (row_begin, row_end, col_begin, col_end) per group, in pixels.
        max_rows / max_cols: upper bounds on the window sizes (launch geometry).
        pixel_bindings: {Var | label: image tensor [rows, res_y]} read per pixel (an upstream dL); a pair of images
        ``(a, b)`` binds the argument to ``a - b`` (formed on load: the difference image is never written).
        rows: the band of grid rows this device owns (multi-GPU sharding).
        Output: ``out`` images [k, rows, res_y] (assigned, or added to with ``accumulate`` -- the windows of
        one launch must then be disjoint), or with ``reduce`` a tensor [G_range, k] of per-group sums.
        prepare_only: run only the per-group prologue (uniform rows + tile classification for the groups' current
        parameters) for ``group_range``; later calls with ``prepared=True`` over sub-ranges then skip it (the same
        windows tensor, window bounds and rows; the parameters must not change in between).
        gather: ``(bin_begin, bin_groups)`` int32 CUDA tensors (``group_bins``): ALL groups in one launch over the image,
        every block walking the groups of its bin in ascending order -- for accumulating image programs whose windows
        overlap; the result is that of launching the groups one after the other.  ``clear``: the gather launch zeroes
        ``out`` itself (every block clears its pixels first) instead of adding to what it holds."""
        import torch
        pos = {lab: i for i, (lab, _) in enumerate(self.arglist)}
        pix = {(k if isinstance(k, str) else label_of(k)): v for k, v in (pixel_bindings or {}).items()}
        grp = {(k if isinstance(k, str) else label_of(k)): v for k, v in (group_bindings or {}).items()}
        # a pixel binding may be a pair (a, b) of images: the argument is then a[p] - b[p], formed on load
        ks, prog = self.grouped_program(box_vars, list(pix), reduce=reduce, accumulate=accumulate,
                                        pixel_diff=[lab for lab, v in pix.items() if isinstance(v, (tuple, list))])
        tdt = torch.float32 if self.float_width == 'float' else torch.float64
        dev = torch.device('cuda', self.device)
        G = int(windows.shape[0])
        g0, g1 = (0, G) if group_range is None else group_range
        gptrs, keep = [], []
        live = self.program.live
        for a in ks.scalar_args:
            lab, default = self.arglist[a]
            v = grp[lab] if lab in grp else (live[lab].value if lab in live else default)
            assert v is not None, f'No bindings found for {[lab]}'
            if not _is_torch(v) or v.dim() == 0:
                v = torch.full((G,), float(v), dtype=tdt, device=dev)
            assert v.is_cuda and v.dtype == tdt and v.numel() == G and v.is_contiguous(), \
                f'group binding {lab} must be a contiguous CUDA tensor [{G}] of the working precision'
            keep.append(v)
            gptrs.append(v.data_ptr())
        r0, r1 = (0, res[0]) if rows is None else rows
        pptrs = []
        for a in ks.pixel_args:
            bound = pix[self.arglist[a][0]]
            for img in (bound if isinstance(bound, (tuple, list)) else (bound,)):
                assert img.is_cuda and img.dtype == tdt and img.is_contiguous() and img.numel() == (r1 - r0) * res[1]
                pptrs.append(img.data_ptr())
                keep.append(img)
        assert windows.is_cuda and windows.dtype == torch.int32 and windows.is_contiguous() and windows.shape[1] == 4
        grid = runtime.Grid(float(bounds[0][0]), float(bounds[0][1]), float(bounds[1][0]), float(bounds[1][1]),
                            int(res[0]), int(res[1]), int(r0), int(r1))
        self._keepalive = keep
        if prepare_only:
            if g1 > g0 and r1 > r0:
                prog.prepare_groups(gptrs, g0, g1 - g0, grid, windows.data_ptr(), int(max_rows), int(max_cols),
                                    stream=stream)
            return None
        if out is None:
            out = (torch.empty((g1 - g0, ks.n_outputs), dtype=tdt, device=dev) if reduce
                   else torch.zeros((ks.n_outputs, r1 - r0, res[1]), dtype=tdt, device=dev))
        optrs = [out.data_ptr()] if reduce else [out[k].data_ptr() for k in range(ks.n_outputs)]
        if gather is not None:
            assert accumulate and not reduce and ks.gather_kernel and (g0, g1) == (0, G), \
                'gather launches: accumulating image programs, all groups at once'
            bb, bg = gather
            n_bins = -(-int(res[1]) // ks.block_size) * -(-(r1 - r0) // max(ks.tile_rows, 1))
            assert bb.is_cuda and bb.dtype == torch.int32 and bb.is_contiguous() and bb.numel() == n_bins + 1
            assert bg.is_cuda and bg.dtype == torch.int32 and bg.is_contiguous()
            self._keepalive = keep + [bb, bg]
            prog.render_groups_gather(gptrs, G, grid, windows.data_ptr(), int(max_rows), int(max_cols), bb.data_ptr(),
                                      bg.data_ptr(), pptrs, optrs, stream=stream, clear=clear)
            return out
        prog.render_groups(gptrs, g0, g1 - g0, grid, windows.data_ptr(), int(max_rows), int(max_cols), pptrs,
                           optrs, stream=stream, prepared=prepared)
        return out

    def prepare(self, bindings={}, reduce: bool = False):
        """Bind once, launch many times: resolves the arguments, moves array bindings to the device, allocates
        the output and returns a ``PreparedCall`` whose ``run()`` is a single C-ABI call (no per-call Python
        marshalling).  For optimisation loops and for timing the kernels rather than the interpreter."""
        import torch
        args = self._resolve(bindings)
        tdt = torch.float32 if self.float_width == 'float' else torch.float64
        dev = torch.device('cuda', self.device)
        n = None
        array_args = []
        for i, a in enumerate(args):
            ln = a.numel() if (_is_torch(a) and a.dim() > 0) else (int(np.size(a)) if (not _is_torch(a) and np.ndim(a) > 0) else None)
            if ln is None:
                continue
            assert n is None or n == ln, 'all array bindings must have the same length'
            n = ln
            array_args.append(i)
        n_eff = 1 if n is None else n
        team = self.pick_team(n_eff, array_args)
        tens = {}
        for a in array_args:
            t = args[a]
            if not _is_torch(t):
                t = torch.as_tensor(np.asarray(t))
            tens[a] = t.to(device=dev, dtype=tdt).contiguous().view(-1)
        k_out = len(self.dag.outputs)
        out = (torch.empty(k_out, dtype=tdt, device=dev) if reduce
               else torch.empty((k_out, n_eff), dtype=tdt, device=dev))
        ptrs = [t.data_ptr() for t in tens.values()] + ([] if reduce else [out[k].data_ptr() for k in range(k_out)])
        ks, prog = self.compiled(array_args, reduce=reduce, team=team,
                                 vec=self.pick_vec(array_args, reduce, team, ptrs), private=True)
        ins = [tens[a] for a in ks.array_args]
        scalars = [float(args[a]) for a in ks.scalar_args]
        return PreparedCall(self, ks, prog, scalars, ins, out, n_eff, reduce)

    # ------------------------------------------------------------------- render
    def prepare_render(self, box_vars, res: Tuple[int, int], bounds=((0, 1), (0, 1)), bindings=None,
                       rows: Optional[Tuple[int, int]] = None, out=None, reduce: bool = False,
                       interleave: Tuple[int, int] = (1, 0)) -> "PreparedRender":
        """``render`` bound once (arguments resolved, output allocated): the returned object's ``run(stream)`` costs
        one C call.  Same arguments and output conventions as ``render`` (bundles: ``out`` = (images, sums))."""
        import torch
        labels = [v if isinstance(v, str) else label_of(v) for v in box_vars]
        pos = {lab: i for i, (lab, _) in enumerate(self.arglist)}
        grid_args = [pos[lab] for lab in labels]
        # a prepared render owns its handle (its own uniform table and scratch): several of them -- also of the same
        # program with different scalars, on different streams or threads -- run concurrently
        ks, prog = self.compiled(grid_args, grid_args, reduce=reduce, private=True)
        scalars = self._resolve_scalars(bindings, ks.scalar_args)
        r0, r1 = (0, res[0]) if rows is None else rows
        tdt = torch.float32 if self.float_width == 'float' else torch.float64
        dev = torch.device('cuda', self.device)
        n_local = len(interleaved_rows(r1 - r0, ks.tile_rows, *interleave)) if tuple(interleave) != (1, 0) else r1 - r0
        nr = ks.reduce_outputs
        n_store = ks.n_outputs - nr
        if 0 < nr < ks.n_outputs:
            # bundle (value + derivative in ONE kernel): the leading outputs are images, the trailing ones sums
            imgs, sums = out if out is not None else (None, None)
            if imgs is None:
                imgs = torch.empty((n_store, n_local, res[1]), dtype=tdt, device=dev)
            if sums is None:
                sums = torch.empty(nr, dtype=tdt, device=dev)
            assert imgs.shape[1] == n_local and imgs.is_contiguous(), 'output image has the wrong rows'
            out = (imgs, sums)
            ptrs = [imgs[k].data_ptr() for k in range(n_store)] + [sums.data_ptr()]
        else:
            if out is None:
                out = (torch.empty(ks.n_outputs, dtype=tdt, device=dev) if nr
                       else torch.empty((ks.n_outputs, n_local, res[1]), dtype=tdt, device=dev))
            assert nr or (out.shape[1] == n_local and out.is_contiguous()), 'output image has the wrong rows'
            ptrs = [out.data_ptr()] if nr else [out[k].data_ptr() for k in range(ks.n_outputs)]
        grid = runtime.Grid(float(bounds[0][0]), float(bounds[0][1]), float(bounds[1][0]), float(bounds[1][1]),
                            int(res[0]), int(res[1]), int(r0), int(r1))
        return PreparedRender(self, ks, prog, scalars, grid, out, ptrs, interleave)

    def render(self, box_vars, res: Tuple[int, int], bounds=((0, 1), (0, 1)), bindings=None,
               rows: Optional[Tuple[int, int]] = None, out=None, stream: int = 0, reduce: bool = False,
               interleave: Tuple[int, int] = (1, 0)):
        """Evaluate over the pixel grid of tests/plot.py:8-29 with on-device pixel boxes.

        box_vars: (x_lb, x_ub, y_lb, y_ub) Vars or labels.  Returns a torch CUDA tensor
        [k, rows, res_y] (k = program outputs), or [k] sums over the pixels when ``reduce`` is True; with
        ``reduce=r`` (an int < k, for value + derivative bundles) the pair (images [k - r, rows, res_y],
        sums [r]).  Asynchronous on ``stream``.
        interleave=(stride, phase): only the tile rows phase, phase + stride, ... of the row range (interleaved
        sharding over ranks: every rank gets the same mix of tiles); the image then holds those rows, compact
        (``interleaved_rows`` lists them)."""
        import torch
        labels = [v if isinstance(v, str) else label_of(v) for v in box_vars]
        pos = {lab: i for i, (lab, _) in enumerate(self.arglist)}
        grid_args = [pos[lab] for lab in labels]
        ks, prog = self.compiled(grid_args, grid_args, reduce=reduce)
        scalars = self._resolve_scalars(bindings, ks.scalar_args)
        r0, r1 = (0, res[0]) if rows is None else rows
        tdt = torch.float32 if self.float_width == 'float' else torch.float64
        dev = torch.device('cuda', self.device)
        nr = ks.reduce_outputs
        n_store = ks.n_outputs - nr
        n_local = len(interleaved_rows(r1 - r0, ks.tile_rows, *interleave)) if tuple(interleave) != (1, 0) else r1 - r0
        grid = runtime.Grid(float(bounds[0][0]), float(bounds[0][1]), float(bounds[1][0]), float(bounds[1][1]),
                            int(res[0]), int(res[1]), int(r0), int(r1))
        if 0 < nr < ks.n_outputs:
            # bundle: the leading outputs are images, the trailing ``reduce`` outputs are summed over the pixels
            imgs, sums = out if out is not None else (None, None)
            if imgs is None:
                imgs = torch.empty((n_store, n_local, res[1]), dtype=tdt, device=dev)
            if sums is None:
                sums = torch.empty(nr, dtype=tdt, device=dev)
            prog.render(scalars, grid, [imgs[k].data_ptr() for k in range(n_store)] + [sums.data_ptr()], stream=stream,
                        interleave=interleave)
            return imgs, sums
        if out is None:
            out = (torch.empty(ks.n_outputs, dtype=tdt, device=dev) if nr
                   else torch.empty((ks.n_outputs, n_local, res[1]), dtype=tdt, device=dev))
        ptrs = [out.data_ptr()] if nr else [out[k].data_ptr() for k in range(ks.n_outputs)]
        prog.render(scalars, grid, ptrs, stream=stream, interleave=interleave)
        return out


def interleaved_rows(n_rows: int, tile_rows: int, stride: int, phase: int) -> List[int]:
    """Row indices (relative to the row range) rendered with ``interleave=(stride, phase)``: tile rows phase,
    phase + stride, ... of ``tile_rows`` pixel rows each, in the order of the compact output image."""
    out: List[int] = []
    n_tr = -(-n_rows // tile_rows)
    for t in range(phase, n_tr, stride):
        out.extend(range(t * tile_rows, min((t + 1) * tile_rows, n_rows)))
    return out


class PreparedRender:
    """A bound ``render`` (see ``CUDA_EvalMode.prepare_render``): ``run(stream)`` is one C-ABI call
    (``tegb_render_launch``) with pre-marshalled arguments; ``set_scalars`` rebinds the launch-wide arguments."""

    def __init__(self, mode, ks, prog, scalars, grid, out, ptrs, interleave=(1, 0)):
        self.mode, self.ks, self.prog, self.out = mode, ks, prog, out
        self._grid = grid
        self._scalars = runtime._dbl_array(scalars)
        self._ptrs = runtime._ptr_array(ptrs)
        self._fn = runtime.lib().tegb_render_launch_strided
        self._h = prog.handle
        self._stride, self._phase = int(interleave[0]), int(interleave[1])

    def set_scalars(self, values: Sequence[float]) -> None:
        """New values for the scalar arguments, in ``ks.scalar_args`` order."""
        self._scalars = runtime._dbl_array(values)

    def run(self, stream: int = 0):
        runtime.check(self._fn(self._h, self._scalars, ctypes.byref(self._grid), self._stride, self._phase, self._ptrs,
                               ctypes.c_void_p(stream)))
        return self.out


class PreparedCall:
    """A bound launch (see ``CUDA_EvalMode.prepare``).  ``run()`` re-launches on the current torch stream and
    returns the (reused) output tensor: [n] / [k, n], or [k] sums for ``reduce``."""

    def __init__(self, mode, ks, prog, scalars, ins, out, n, reduce):
        self.mode, self.ks, self.prog = mode, ks, prog
        self.ins, self.out, self.n, self.reduce = ins, out, n, reduce
        self._scalars = runtime._dbl_array(scalars)
        self._in = runtime._ptr_array([t.data_ptr() for t in ins])
        outs = [out.data_ptr()] if reduce else [out[k].data_ptr() for k in range(ks.n_outputs)]
        self._out = runtime._ptr_array(outs)
        self._fn = runtime.lib().tegb_program_launch
        self._h = prog.handle

    def set_scalars(self, values: Sequence[float]) -> None:
        """New values for the launch-wide (scalar-bound) arguments, in ``ks.scalar_args`` order."""
        self._scalars = runtime._dbl_array(values)

    def run(self, stream: Optional[int] = None):
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.out.device).cuda_stream
        runtime.check(self._fn(self._h, self._scalars, self._in, self._out, self.n, ctypes.c_void_p(stream)))
        squeeze = not (self.mode.out_is_tuple and self.ks.n_outputs > 1)
        return self.out[0] if squeeze else self.out


BACKENDS = {'CUDA': (CUDA_EvalMode, {})}


def evaluate(expr, bindings=None, backend=None, **kwargs):
    """reference teg/eval/backend.py:13-41, with 'CUDA' as the default backend of this package."""
    if not hasattr(expr, 'mode_cache'):
        setattr(expr, 'mode_cache', {})
        setattr(expr, 'mode_options', kwargs)
    bindings = {} if bindings is None else bindings
    backend = 'CUDA' if backend is None else backend
    if backend != 'CUDA':
        try:
            from teg.eval.backend import evaluate as ref_evaluate
        except ImportError as e:
            raise KeyError(backend) from e
        return ref_evaluate(expr, bindings, backend=backend, **kwargs)
    if backend not in expr.mode_cache.keys() or \
            not all([(key, value) in expr.mode_options.items() for key, value in kwargs.items()]):
        backend_cls, backend_kwargs = BACKENDS[backend]
        mode = backend_cls(expr, **kwargs, **backend_kwargs)
        expr.mode_cache[backend] = mode
        expr.mode_options = kwargs
    return expr.mode_cache[backend].eval(bindings=bindings)


def register() -> bool:
    """Add 'CUDA' to the reference's registry (teg/eval/backend.py:6-10) when ``teg`` is importable."""
    try:
        import teg.eval.backend as ref_backend
    except Exception:
        return False
    ref_backend.BACKENDS.setdefault('CUDA', (CUDA_EvalMode, {}))
    return True


def render_image(expr, variables=(), res=(64, 64), bounds=((0, 1), (0, 1)), bindings={}, num_samples=20,
                 float_width='float', device=0):
    """tests/plot.py:8-29 in one launch.  ``variables`` = ((x_lb, x_ub), (y_lb, y_ub)).

    Returns ``image[nx, ny]`` (numpy), or ``[k, nx, ny]`` for Tup-valued programs."""
    assert len(variables) == 2, 'Exactly two variable-pairs required'
    import torch
    if not hasattr(expr, 'mode_cache'):
        setattr(expr, 'mode_cache', {})
    key = ('CUDA_render', num_samples, float_width, device)
    mode = expr.mode_cache.get(key)
    if mode is None:
        mode = CUDA_EvalMode(expr, num_samples=num_samples, float_width=float_width, device=device)
        expr.mode_cache[key] = mode
    box = (variables[0][0], variables[0][1], variables[1][0], variables[1][1])
    out = mode.render(box, res, bounds, bindings)
    torch.cuda.synchronize(device)
    img = out.cpu().numpy()
    return img if (mode.out_is_tuple and img.shape[0] > 1) else img[0]
