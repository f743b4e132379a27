#!/usr/bin/env python
"""bench.py -- integral evaluations / s for the BASELINE.json workloads.

    python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, sm_100a)
    python bench.py --impl reference ...                    # the reference's C backend on the host cores

Headline workload (config.workload = "tri_raster_4096", BASELINE.json's target configuration): the per-pixel triangle
rasterization integral plus its reverse derivative w.r.t. the six vertex coordinates (oracle/scenes.py: tri_primal +
tri_grad, the program of SURVEY.md 8(d) config 2) on a 4096 x 4096 pixel grid per GPU, 16 midpoint samples per
integral (256 samples per pixel for the value, 16 per active delta term for the gradient).  One *step* = one pass over
the whole batch: value image, per-pixel gradients, and the per-parameter gradient sum (+ one all-reduce of the
6-vector when N > 1).  One *integral evaluation* = one pixel instance (value + its reverse derivative).
--scaling weak (default): the image grows to (4096 N) x 4096 pixels over the same domain and rank r owns the tile rows
r, r + N, ... (16 pixel rows each; --shard band: the contiguous rows [4096 r, 4096 (r+1))).  --scaling strong: the
image stays 4096 x 4096 and is split over the ranks.

The same JSON line also carries, under "workloads", the other driver-visible measurements (each with its own check):

  multi_tri_4096 (+ _strong)   BASELINE.json configs[4]: 4096 x 4096, 256 triangles, forward image + reverse-mode
                               [K, 7] gradients with an upstream dL = image - target (teg_b200/raster.py); the
                               gradient all-reduce is fused into the per-triangle column sums over NVLink peer memory
  tri_raster_4096_strong       (N > 1) the headline programs on a fixed 4096 x 4096 image split over the ranks
  suite                        (N = 1) configs 1-4 of SURVEY.md 8(d): roofline + reference-C baseline + bit_exact each
  existing_kernel              (N = 1) the reference's own --target CUDA_C output in a thread-per-pixel wrapper
                               (oracle/build_existing_kernel.py), timed on the same GPU: the bar SURVEY 2.1 names
  d2h_probe                    what bounds the end-to-end number: concurrent device-to-host copies of one image per rank

--workload multi_tri_4096 makes that workload the headline of the line instead (value / ms_per_step / roofline).

Nothing here reads /root/reference: programs come from tests/golden/programs/*.tegp, the CPU baseline from
oracle/_ref/*.so (the reference's own generated C, built where the reference exists) or, failing that, from the oracle
port.  oracle/ is executed only by the baseline legs (cpu_baseline, --impl reference, existing_kernel) and the checks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PROGRAM_DIR = os.path.join(ROOT, 'tests', 'golden', 'programs')
NUM_SAMPLES = 16
ANALYTIC_GRAD = [-0.3, -0.225, 0.35, -0.125, -0.05, 0.35]       # d area / d (x0 y0 x1 y1 x2 y2) of workloads.TRIANGLE
AREA = 0.2325


def load_text(scene: str) -> str:
    with open(os.path.join(PROGRAM_DIR, f'{scene}.tegp')) as f:
        return f.read()


def headline_config(workload: str, res: int, scaling: str):
    """config of the JSON line: identical in the CUDA arm and the reference arm (the driver compares them)"""
    if workload == 'multi_tri_4096':
        return {'workload': workload, 'res_per_gpu': [res, res] if scaling == 'weak' else None,
                'res_total': [res, res] if scaling == 'strong' else None, 'triangles_per_band': 256,
                'num_samples': NUM_SAMPLES, 'programs': ['tri_color_primal', 'tri_color_grad'], 'scaling': scaling}
    return {'workload': workload, 'res_per_gpu': [res, res] if scaling == 'weak' else None,
            'res_total': [res, res] if scaling == 'strong' else None, 'num_samples': NUM_SAMPLES,
            'programs': ['tri_primal', 'tri_grad'], 'scaling': scaling}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)'
    except Exception:
        return 6650.0, 'fallback stated in B200_PROFILING.md (MEASURED_PEAKS.json absent)'


# --------------------------------------------------------------------------------------------- stdout
class _StdoutToStderr:
    """File-descriptor level redirection of stdout to stderr: libraries that print to stdout from C (NCCL announces its
    version there when it initialises) must not get in front of the ONE JSON line this program prints."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)          # C stdio buffers too
        except Exception:
            pass
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


# ------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '100',
                 '-i', str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------- reference (CPU) arm
def cpu_runner(threads: int, scenes=('tri_primal', 'tri_grad'), ns: int = NUM_SAMPLES, dtype=np.float32):
    """Returns (kind, run(scene, inputs) -> ndarray).  'reference' = oracle/_ref (the reference's own C)."""
    from oracle import Oracle, RefLib, ref_lib_path
    if all(os.path.exists(ref_lib_path(s, dtype, ns)) for s in scenes):
        libs = {s: RefLib(s, dtype, ns) for s in scenes}
        return 'reference', lambda scene, inputs: libs[scene].eval(inputs, threads=threads)
    libs = {s: Oracle(load_text(s)) for s in scenes}
    return 'port', lambda scene, inputs: libs[scene].eval(inputs, ns, dtype, threads=threads)


def cpu_sample_inputs(res_x: int, res_y: int, n_rows: int):
    """Bounded sample of the tri_raster workload: n_rows pixel rows spread evenly over the res_x x res_y grid
    (n_rows >= res_x: the whole grid)."""
    from teg_b200.tegp import Program
    from teg_b200.workloads import TRIANGLE, bind, pixel_boxes
    prog = Program.loads(load_text('tri_primal'))
    if n_rows >= res_x:
        box = list(pixel_boxes((res_x, res_y)))
        n_rows = res_x
    else:
        stride = max(1, res_x // n_rows)
        rows = np.arange(0, res_x, stride)[:n_rows]
        parts = [pixel_boxes((res_x, res_y), rows=(int(r), int(r) + 1)) for r in rows]
        box = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        n_rows = len(rows)
    vals = dict(zip(('x_lb', 'x_ub', 'y_lb', 'y_ub'), box), **TRIANGLE)
    return bind(prog.args, vals), n_rows * res_y, n_rows


def multi_tri_cpu_sample(scene, res: int, every_tri: int = 8, every_row: int = 4):
    """Bounded sample of the multi_tri workload: every 8th triangle, every 4th row of its window (SoA columns in the
    argument order of tri_color_grad: dL first; tri_color_primal takes columns [1:])."""
    xs = np.linspace(0.0, 1.0, res + 1)
    cols = [[] for _ in range(12)]
    K = scene['verts'].shape[0]
    for k in range(0, K, every_tri):
        a0, a1, c0, c1 = scene['windows'][k]
        rr = np.arange(a0, min(a1, res), every_row)
        nx = np.repeat(rr, c1 - c0)
        ny = np.tile(np.arange(c0, c1), len(rr))
        vals = [np.full(nx.size, 0.1), xs[nx], xs[nx + 1], xs[ny], xs[ny + 1]] + \
               [np.full(nx.size, v) for v in scene['verts'][k]] + [np.full(nx.size, scene['col'][k])]
        for i, v in enumerate(vals):
            cols[i].append(v)
    cols = [np.concatenate(c) for c in cols]
    return cols, cols[0].size


def run_reference_arm(args) -> int:
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    world = args.gpus
    res = args.res
    if args.workload == 'multi_tri_4096':
        from teg_b200.workloads import multi_triangle_scene
        kind, run = cpu_runner(threads, ('tri_color_primal', 'tri_color_grad'))
        scene = multi_triangle_scene(bands=1, res=res)
        cols, n_inst = multi_tri_cpu_sample(scene, res, 8 if kind == 'reference' else 64, 4)
        sample = f'every {8 if kind == "reference" else 64}th triangle of one band, every 4th row of its window ({n_inst} pixel-triangle pairs) per step'

        def one():
            run('tri_color_primal', cols[1:])
            run('tri_color_grad', cols)
    else:
        kind, run = cpu_runner(threads)
        res_x = res * world if args.scaling == 'weak' else res
        # size the sample from a probe so that one step is about 2 s on this host: the whole grid where that fits
        probe, n_probe, _ = cpu_sample_inputs(res_x, res, 64 if kind == 'reference' else 4)
        run('tri_primal', probe)
        t0 = time.perf_counter()
        run('tri_primal', probe)
        run('tri_grad', probe)
        per_row = (time.perf_counter() - t0) / (n_probe / res)
        rows = int(max(8, min(res_x, args.cpu_seconds / max(per_row, 1e-9))))
        if args.cpu_rows:
            rows = min(res_x, args.cpu_rows)
        inputs, n_inst, rows = cpu_sample_inputs(res_x, res, rows)
        sample = (f'the whole {res_x}x{res} grid ({n_inst} instances) per step' if rows >= res_x else
                  f'{rows} evenly spaced pixel rows of the {res_x}x{res} grid ({n_inst} instances) per step')

        def one():
            run('tri_primal', inputs)
            run('tri_grad', inputs)
    for _ in range(max(args.warmup, 1)):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    value = n_inst * args.steps / dt
    line = {
        'impl': 'reference', 'metric': 'integral_evals_per_s', 'value': value, 'unit': 'integral evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': headline_config(args.workload, res, args.scaling),
        'cpu_baseline': {'value': value, 'unit': 'integral evals/s', 'cores': threads, 'kind': kind,
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'integral evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------- context
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        assert self.world == args.gpus or self.world == 1, f'--gpus {args.gpus} but WORLD_SIZE={self.world}'
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device -- the CUDA backend has no CPU fallback')
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            with _StdoutToStderr():
                dist.init_process_group('nccl', device_id=self.dev)
                dist.barrier()          # the communicator exists (and has said so) before stdout is restored
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.args = args

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def timed_steps(self, step, steps: int, warmup: int):
        """W untimed steps, then K steps timed with CUDA events on the launching stream (L2 flushed before each, outside
        the event pair), barrier + synchronize on both sides, max over ranks.  Returns (ms per step, launches)."""
        from teg_b200 import runtime
        torch = self.torch
        for _ in range(warmup):
            step(False)
        self.sync_all()
        l0 = runtime.launch_count()
        pairs = []
        for k in range(steps):
            self.flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            step(True)
            b.record(self.stream)
            pairs.append((a, b))
        self.sync_all()
        total_ms = self.max_over_ranks(float(sum(a.elapsed_time(b) for a, b in pairs)))
        return total_ms / steps, runtime.launch_count() - l0


# ------------------------------------------------------------------------------------ tri_raster_4096
class TriRaster:
    """value image + reverse derivative of every pixel summed per parameter inside the kernels + the one all-reduce
    when N > 1.  The two programs are independent: the gradient pipeline (classification, kernel, column sums,
    all-reduce) runs on a second stream next to the value pipeline, joined at the end."""

    def __init__(self, ctx: Ctx, scaling: str, allreduce: str = 'peer', fused: bool = False):
        from teg_b200 import CUDA_EvalMode
        from teg_b200.eval import interleaved_rows
        from teg_b200.tegp import Program
        from teg_b200.workloads import TRIANGLE
        torch, args = ctx.torch, ctx.args
        self.ctx, self.scaling = ctx, scaling
        world, rank, res = ctx.world, ctx.rank, args.res
        self.res = res
        self.res_x_total = res * world if scaling == 'weak' else res
        if world == 1:
            self.rows, self.inter = (0, res), (1, 0)
        elif args.shard == 'band':
            per = self.res_x_total // world
            self.rows, self.inter = (per * rank, per * (rank + 1)), (1, 0)
        else:
            assert self.res_x_total % (args.tile_rows * world) == 0, 'interleaved sharding: rows divisible by tile height x ranks'
            self.rows, self.inter = (0, self.res_x_total), (world, rank)
        self.n_local_rows = (self.rows[1] - self.rows[0]) if self.inter == (1, 0) else \
            len(interleaved_rows(self.res_x_total, args.tile_rows, world, rank))
        self.n_local = self.n_local_rows * res
        self.primal = Program.loads(load_text('tri_primal'))
        self.grad = Program.loads(load_text('tri_grad'))
        self.fw = 'float' if args.dtype == 'f32' else 'double'
        self.tdt = torch.float32 if args.dtype == 'f32' else torch.float64
        self.esz = 4 if args.dtype == 'f32' else 8
        xopts = dict(nvrtc_opts=f'--maxrregcount={args.maxreg}' if args.maxreg else '', tile_rows=args.tile_rows,
                     block_size=args.block, min_blocks=None if args.min_blocks < 0 else args.min_blocks,
                     coop_groups=args.coop_groups)
        self.m_primal = CUDA_EvalMode(self.primal, num_samples=NUM_SAMPLES, device=ctx.local, float_width=self.fw,
                                      cull=bool(args.cull), **xopts)
        self.m_grad = CUDA_EvalMode(self.grad, num_samples=NUM_SAMPLES, device=ctx.local, float_width=self.fw,
                                    cull=bool(args.cull), **xopts)
        self.box = self.primal.args[:4]
        self.params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in self.primal.args[4:]}
        self.gparams = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in self.grad.args[4:]}
        self.bounds = ((0.0, 1.0), (0.0, 1.0))
        dev = ctx.dev
        self.imgs = [torch.empty((1, self.n_local_rows, res), dtype=self.tdt, device=dev) for _ in range(2)]
        self.gsums = [torch.zeros(6, dtype=self.tdt, device=dev) for _ in range(2)]
        grid = (self.res_x_total, res)
        self.fused = bool(fused)
        if self.fused:
            # value and reverse derivative as ONE program (tegp.bundle): one classification, one main kernel that writes
            # the image and the per-tile gradient partials, one column sum
            self.m_fused = CUDA_EvalMode([self.primal, self.grad], num_samples=NUM_SAMPLES, device=ctx.local,
                                         float_width=self.fw, cull=bool(args.cull), **xopts)
            self.pr_fused = [self.m_fused.prepare_render(self.box, grid, self.bounds, self.gparams, rows=self.rows,
                                                         out=(self.imgs[b], self.gsums[b]), reduce=6,
                                                         interleave=self.inter) for b in range(2)]
            self.pr_grad = self.pr_fused
            self.pr_primal = self.pr_fused
        else:
            self.pr_grad = [self.m_grad.prepare_render(self.box, grid, self.bounds, self.gparams, rows=self.rows,
                                                       out=self.gsums[b], reduce=True, interleave=self.inter) for b in range(2)]
            self.pr_primal = [self.m_primal.prepare_render(self.box, grid, self.bounds, self.params, rows=self.rows,
                                                           out=self.imgs[b], interleave=self.inter) for b in range(2)]
        # the gradient pipeline is the longer one (and ends in the all-reduce): its blocks get priority
        self.side = torch.cuda.Stream(dev, priority=-1)
        self.forks = [torch.cuda.Event() for _ in range(2)]
        self.joins = [torch.cuda.Event() for _ in range(2)]
        self.peers, self.peer_note = None, None
        if world > 1 and allreduce == 'peer':
            from teg_b200.parallel import PeerGroup
            try:
                self.peers = PeerGroup(ctx.local, max_values=16)       # raises on every rank or on none
                for pr in self.pr_grad:
                    self.peers.attach(pr.prog)          # (fused: the bundle's handle)      # the last column-sum kernel of the gradient does the exchange itself
            except RuntimeError as e:
                self.peers, self.peer_note = None, f'peer mailboxes unavailable ({e}): NCCL all-reduce instead'
        self.ev = {'primal': [], 'grad': []}
        self.graph = None

    def capture(self):
        """The step as ONE CUDA graph (both streams, classification, kernels, column sums, the fused peer exchange --
        its sequence number lives on the device, so every replay is a new exchange).  A step is ~10 small launches plus
        stream bookkeeping: issued one by one from Python they cost more host time than the GPU needs to run them."""
        torch, ctx = self.ctx.torch, self.ctx
        if ctx.world > 1 and self.peers is None:
            return None                     # the NCCL fallback is not captured
        cap = torch.cuda.Stream(ctx.dev)
        cap.wait_stream(ctx.stream)
        with torch.cuda.stream(cap):
            self.step(False, 0, cap)
        ctx.stream.wait_stream(cap)
        ctx.sync_all()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            self.step(False, 0, torch.cuda.current_stream(ctx.dev))
        self.graph = g
        return g

    def step(self, record: bool, buf: int = 0, stream=None):
        if self.graph is not None and not record and buf == 0 and stream is None:
            self.graph.replay()             # one submission: both pipelines, the fused peer exchange included
            return
        ctx, torch = self.ctx, self.ctx.torch
        stream = ctx.stream if stream is None else stream
        side = self.side
        if self.fused:
            a, b = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if record else (None, None)
            if record:
                a.record(stream)
            self.pr_fused[buf].run(stream.cuda_stream)
            if ctx.world > 1 and self.peers is None:
                ctx.dist.all_reduce(self.gsums[buf])
            if record:
                b.record(stream)
                self.ev['primal'].append((a, b))
                self.ev['grad'].append((a, b))
            return
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        fork, join = self.forks[buf], self.joins[buf]
        fork.record(stream)
        side.wait_event(fork)
        if record:
            e[2].record(side)
        self.pr_grad[buf].run(side.cuda_stream)
        if ctx.world > 1 and self.peers is None:
            with torch.cuda.stream(side):
                ctx.dist.all_reduce(self.gsums[buf])        # enqueued behind the gradient on `side`
        if record:
            e[3].record(side)
        join.record(side)
        if record:
            e[0].record(stream)
        self.pr_primal[buf].run(stream.cuda_stream)
        if record:
            e[1].record(stream)
        stream.wait_event(join)
        if record:
            self.ev['primal'].append((e[0], e[1]))
            self.ev['grad'].append((e[2], e[3]))

    def check(self):
        """Sum of the value image over ALL ranks = triangle area; the all-reduced gradient = its analytic derivative
        (to quadrature error) and, for N > 1, = an NCCL all-reduce of the ranks' local sums."""
        ctx, torch = self.ctx, self.ctx.torch
        ctx.sync_all()
        area_local = float(self.imgs[0].double().sum().item())
        area = ctx.sum_over_ranks(area_local)
        g = self.gsums[0].double().cpu().numpy()
        out = {'sum_value_all_ranks': area, 'expected_area': AREA, 'area_ok': bool(abs(area - AREA) <= 2e-6),
               'grad_sum': g.tolist(), 'analytic_grad': ANALYTIC_GRAD,
               'grad_vs_analytic_ok': bool(np.max(np.abs(g - np.array(ANALYTIC_GRAD))) <= 0.03),
               'grad_vs_analytic_note': 'sanity bound 0.03: the 16-sample midpoint rule on the delta terms is not the exact '
                                        'boundary integral (the bit-level statement is the parity suite)'}
        if ctx.world > 1:
            # once, outside every timed region: the ranks' LOCAL sums (a handle without the peer group), all-reduced
            # by NCCL, against what the fused peer exchange delivered
            loc = torch.zeros(6, dtype=self.tdt, device=ctx.dev)
            pr = self.m_grad.prepare_render(self.box, (self.res_x_total, self.res), self.bounds, self.gparams,
                                            rows=self.rows, out=loc, reduce=True, interleave=self.inter)
            pr.run(ctx.stream.cuda_stream)
            ctx.dist.all_reduce(loc)
            torch.cuda.synchronize(ctx.dev)
            ref = loc.double().cpu().numpy()
            tol = 1e-6 if self.esz == 4 else 1e-13
            out['grad_nccl_of_local_sums'] = ref.tolist()
            out['peer_vs_nccl_ok'] = bool(np.max(np.abs(g - ref)) <= tol * max(1.0, float(np.max(np.abs(ref)))))
            same = [torch.zeros(6, dtype=self.tdt, device=ctx.dev) for _ in range(ctx.world)]
            ctx.dist.all_gather(same, self.gsums[0])
            out['same_bits_on_every_rank'] = bool(all(torch.equal(s, same[0]) for s in same))
        out['ok'] = bool(all(v for k, v in out.items() if k.endswith('_ok') or k == 'same_bits_on_every_rank'))
        return out

    def measure(self, steps: int, warmup: int, use_graph: bool = True):
        """phases: a few steps launch by launch with events around each pipeline; then the timed steps -- replays of the
        captured graph (use_graph) or launch-by-launch steps."""
        from teg_b200 import runtime
        ctx = self.ctx
        for _ in range(max(warmup, 3)):
            self.step(False)
        ctx.sync_all()
        for k in range(5):
            ctx.flush.fill_(k)
            self.step(True)
        ctx.sync_all()
        phase = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in self.ev.items() if v}
        l0 = runtime.launch_count()
        self.step(True)
        per_step = runtime.launch_count() - l0
        ctx.sync_all()
        if use_graph:
            self.capture()
        ms, launches = ctx.timed_steps(lambda rec: self.step(False), steps, max(warmup, 3))
        if self.graph is not None:
            launches = per_step * steps            # a replay runs the step's kernels without passing the launch counter
        n_total = ctx.sum_over_ranks(float(self.n_local))
        return {'ms_per_step': ms, 'value': n_total / (ms * 1e-3), 'unit': 'integral evals/s', 'gpu_launches': int(launches),
                'phase_ms': phase, 'instances_all_ranks': int(n_total), 'cuda_graph': self.graph is not None,
                'phase_note': 'phases timed launch by launch before the timed region (the two pipelines run concurrently '
                              'on two streams: their times overlap)' + ('; the timed steps replay one CUDA graph'
                                                                         if self.graph is not None else '')}

    def close(self):
        if self.peers is not None:
            self.peers.close()          # checks the time-out flag, then a barrier before the mailboxes go away
            self.peers = None


# ------------------------------------------------------------------------------------- multi_tri_4096
class MultiTri:
    """BASELINE.json configs[4]: forward (ONE gather launch: every image block adds the triangles of its bin) + backward with the upstream
    gradient dL = image - target (one grouped launch, per-triangle sums reduced in-kernel, all-reduced over the peer
    mailboxes when N > 1).
    An *evaluation* = one (pixel, triangle) pair inside a triangle's window: its value integral and its reverse
    derivative.  weak: one unit square / band of 4096 rows / 256 triangles per GPU.  strong: ONE band split by rows."""

    def __init__(self, ctx: Ctx, scaling: str):
        from teg_b200.raster import MultiTriangleRasterizer
        from teg_b200.tegp import Program
        from teg_b200.workloads import multi_triangle_scene
        torch, args = ctx.torch, ctx.args
        self.ctx, self.scaling = ctx, scaling
        world, rank, res = ctx.world, ctx.rank, args.res
        primal = Program.loads(load_text('tri_color_primal'))
        grad = Program.loads(load_text('tri_color_grad'))
        self.primal = primal
        mb = (4, 4)
        if scaling == 'weak':
            self.scene = multi_triangle_scene(bands=world, res=res)
            self.r = MultiTriangleRasterizer(primal, grad, res, bands=world, rank=rank, device=ctx.local, min_blocks=mb)
            br = self.scene['band_ranges']
            active = (br[max(rank - 1, 0)][0], br[min(rank + 1, world - 1)][1])      # own band and its neighbours
            rows = (res * rank, res * (rank + 1))
        else:
            from teg_b200.parallel import shard_rows
            self.scene = multi_triangle_scene(bands=1, res=res)
            rows = shard_rows(res, world, rank)
            self.r = MultiTriangleRasterizer(primal, grad, res, bands=1, rank=0, device=ctx.local, min_blocks=mb, rows=rows)
            active = None
        sc = self.scene
        self.K = sc['verts'].shape[0]
        self.r.set_scene(sc['verts'], sc['col'], sc['windows'], sc['classes'], sc['max_rows'], sc['max_cols'], active=active)
        w = sc['windows'].astype(np.int64)
        r0, r1 = np.clip(w[:, 0], rows[0], rows[1]), np.clip(w[:, 1], rows[0], rows[1])
        self.pairs_local = int(((r1 - r0) * (w[:, 3] - w[:, 2])).sum())
        self.pairs_total = int(((w[:, 1] - w[:, 0]) * (w[:, 3] - w[:, 2])).sum())
        n_rows = rows[1] - rows[0]
        dev = ctx.dev
        self.img = torch.empty((1, n_rows, res), dtype=torch.float32, device=dev)
        self.target = torch.full((n_rows, res), 0.25, dtype=torch.float32, device=dev)
        self.dL = torch.empty((n_rows, res), dtype=torch.float32, device=dev)
        self.g = torch.empty((self.K, 7), dtype=torch.float32, device=dev)
        self.peers, self.peer_note = None, None
        if world > 1:
            from teg_b200.parallel import PeerGroup
            try:
                self.peers = PeerGroup(ctx.local, max_values=7 * self.K)
                self.r.set_peers(self.peers, diff=True)
            except RuntimeError as e:
                self.peers, self.peer_note = None, f'peer mailboxes unavailable ({e}): NCCL all-reduce instead'
        self.ev = {'forward': [], 'backward': []}
        self.graph = None

    def body(self, sp):
        # the upstream gradient dL = image - target of the L2 loss is formed inside the backward kernel (a pixel argument
        # bound to a pair of images): the same subtraction, no dL image written and read back
        self.r.forward(out=self.img, stream=sp)
        self.r.backward((self.img[0], self.target), out=self.g, stream=sp)

    def step_launches(self, record: bool):
        ctx, torch = self.ctx, self.ctx.torch
        stream = ctx.stream
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            e[0].record(stream)
        self.r.forward(out=self.img, stream=stream.cuda_stream)
        if record:
            e[1].record(stream)
            e[2].record(stream)
        self.r.backward((self.img[0], self.target), out=self.g, stream=stream.cuda_stream)
        if record:
            e[3].record(stream)
        if ctx.world > 1 and self.peers is None:
            ctx.dist.all_reduce(self.g)
        if record:
            self.ev['forward'].append((e[0], e[1]))
            self.ev['backward'].append((e[2], e[3]))

    def step(self, record: bool):
        if self.graph is None:
            return self.step_launches(record)
        self.graph.replay()                 # one submission: the whole step, the fused peer exchange included
        if self.ctx.world > 1 and self.peers is None:
            self.ctx.dist.all_reduce(self.g)

    def measure(self, steps: int, warmup: int, use_graph: bool = True):
        ctx = self.ctx
        for _ in range(3):
            self.step_launches(False)
        ctx.sync_all()
        for _ in range(3):
            self.step_launches(True)              # per-phase times, launch by launch
        ctx.sync_all()
        phase = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in self.ev.items()}
        from teg_b200 import runtime
        l0 = runtime.launch_count()
        self.step_launches(False)
        per_step = runtime.launch_count() - l0          # kernels of this library in one step (a graph replays them)
        if use_graph:
            self.graph = self.r.capture(self.body)
        ms, launches = ctx.timed_steps(lambda rec: self.step(rec), steps, warmup)
        if use_graph:
            launches = per_step * steps
        pairs_all = ctx.sum_over_ranks(float(self.pairs_local))
        return {'ms_per_step': ms, 'value': pairs_all / (ms * 1e-3), 'unit': 'integral evals/s (pixel-triangle pairs)',
                'gpu_launches': int(launches), 'phase_ms': phase, 'pairs_all_ranks': int(pairs_all),
                'triangles': int(self.K), 'cuda_graph': bool(use_graph),
                'phase_note': 'phases timed launch by launch before the timed region'
                + ('; the timed steps replay one CUDA graph' if use_graph else '')}

    def check(self):
        """size-independent properties: image sum and gradient checksums are all-reduced, so a strong-scaling run must
        print the N = 1 values at every N; for N > 1 the fused peer exchange is compared with an NCCL all-reduce of the
        ranks' local [K, 7] sums once, outside every timed region."""
        ctx, torch = self.ctx, self.ctx.torch
        ctx.sync_all()
        self.step_launches(False)
        ctx.sync_all()
        out = {'sum_image_all_ranks': ctx.sum_over_ranks(float(self.img.double().sum().item())),
               'grad_l1': float(self.g.double().abs().sum().item()), 'grad_l2': float(self.g.double().norm().item()),
               'finite': bool(torch.isfinite(self.g).all().item() and torch.isfinite(self.img).all().item())}
        # dL as an explicit image through the other compiled variant (no peers attached): the same bits as the in-kernel
        # difference at N = 1; for N > 1 its local sums, all-reduced by NCCL, against the fused peer exchange
        torch.sub(self.img[0], self.target, out=self.dL)
        r = self.r
        saved, r.peers = r.peers, None
        loc = r.backward(self.dL)
        r.peers = saved
        if ctx.world > 1:
            ctx.dist.all_reduce(loc)
        torch.cuda.synchronize(ctx.dev)
        ref, got = loc.double(), self.g.double()
        scale = float(ref.abs().max().item())
        out['explicit_dL_vs_in_kernel_difference_max_abs'] = float((ref - got).abs().max().item())
        if ctx.world == 1:
            out['explicit_dL_equals_in_kernel_difference'] = bool(torch.equal(loc, self.g))
        elif self.peers is not None:
            out['peer_vs_nccl_max_abs'] = out['explicit_dL_vs_in_kernel_difference_max_abs']
            out['peer_vs_nccl_ok'] = bool(out['peer_vs_nccl_max_abs'] <= 2e-6 * max(scale, 1e-30))
            same = [torch.empty_like(self.g) for _ in range(ctx.world)]
            ctx.dist.all_gather(same, self.g)
            out['same_bits_on_every_rank'] = bool(all(torch.equal(s, same[0]) for s in same))
        out['ok'] = bool(out['finite'] and out.get('peer_vs_nccl_ok', True) and out.get('same_bits_on_every_rank', True)
                         and out.get('explicit_dL_equals_in_kernel_difference', True))
        return out

    def close(self):
        self.graph = None
        if self.peers is not None:
            self.peers.close()
            self.peers = None


# ------------------------------------------------------------------------------------------ the suite
def gpu_time_ms(ctx: Ctx, fn, reps: int = 10, flush: bool = True) -> float:
    """mean CUDA-event time of fn() on the current stream over `reps` launches after 3 warm-ups, L2 flushed before each"""
    torch = ctx.torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize(ctx.dev)
    pairs = []
    for k in range(reps):
        if flush:
            ctx.flush.fill_(k & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ctx.stream)
        fn()
        b.record(ctx.stream)
        pairs.append((a, b))
    torch.cuda.synchronize(ctx.dev)
    return float(np.mean([a.elapsed_time(b) for a, b in pairs]))


def suite_entry(ctx: Ctx, peak_tflops: float, config: str, scene: str, ns: int, res=None, n=None, bound: str = 'fp32',
                cpu_instances: int = 1 << 16):
    """One program of configs 1-4 through the array API (device buffers, a prepared call = ONE C-ABI call per launch):
    time, roofline (HBM: algorithmic bytes = 4 x (array inputs + outputs) per instance; FP32: the reference program's
    operation count), the reference's own C on the host cores over a bounded sample, and bit_exact on it."""
    from oracle import Oracle, RefLib, ref_lib_path
    from teg_b200 import CUDA_EvalMode
    from teg_b200.opcount import count as op_count
    from teg_b200.tegp import Program
    from teg_b200.workloads import bind, scene_values
    torch = ctx.torch
    text = load_text(scene)
    prog = Program.loads(text)
    inputs = bind(prog.args, scene_values(scene, res=res or (256, 256), n=n))
    count = max(int(np.size(v)) for v in inputs)
    mode = CUDA_EvalMode(prog, num_samples=ns, device=ctx.local)
    b = {lab: (torch.as_tensor(np.asarray(v, dtype=np.float32), device=ctx.dev) if np.ndim(v) else v)
         for lab, v in zip(prog.args, inputs)}
    arr = [i for i, v in enumerate(inputs) if np.ndim(v) > 0]
    call = mode.prepare(b)
    sp = ctx.stream.cuda_stream
    big = count * 4 * (len(arr) + len(mode.dag.outputs)) > (200 << 20)        # working set beyond L2: no flush needed
    ms = gpu_time_ms(ctx, lambda: call.run(sp), reps=10, flush=not big)
    out = call.run(sp)
    torch.cuda.synchronize(ctx.dev)
    k_out = len(mode.dag.outputs)
    # parity + CPU baseline on a bounded, evenly spaced sample
    m = min(count, cpu_instances)
    idx = np.linspace(0, count - 1, m).astype(np.int64)
    sub = [np.asarray(v)[idx] if np.ndim(v) else v for v in inputs]
    threads = os.cpu_count() or 1
    if os.path.exists(ref_lib_path(scene, np.float32, ns)):
        lib, kind = RefLib(scene, np.float32, ns), 'reference'
        run = lambda: lib.eval(sub, threads=threads)          # noqa: E731
    else:
        lib, kind = Oracle(text), 'port'
        run = lambda: lib.eval(sub, ns, np.float32, threads=threads)          # noqa: E731
    want = run()
    t0 = time.perf_counter()
    run()
    cpu_s = time.perf_counter() - t0
    got = out.detach().cpu().numpy().reshape(k_out, count)[:, idx]
    exact = bool(np.array_equal(got, want, equal_nan=True))
    scale = np.maximum(np.abs(want), np.max(np.abs(want)) * 1e-3 + 1e-30)
    oc = op_count(mode.reference_schedule(arr), ns)
    io_bytes = 4 * (len(arr) + k_out)
    peak_hbm, src = hbm_peak()
    if bound == 'hbm':
        ach = io_bytes * count / (ms * 1e-3) / 1e9
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': peak_hbm, 'unit': 'GB/s', 'frac': ach / peak_hbm,
                'traffic': None, 'algorithmic_bytes_per_instance': io_bytes, 'peak_source': src}
    else:
        ach = oc.unguarded * count / (ms * 1e-3) / 1e12
        roof = {'bound': 'fp32_fma', 'achieved': ach, 'peak': peak_tflops, 'unit': 'TFLOP/s', 'frac': ach / peak_tflops,
                'traffic': None, 'algorithmic_ops_per_instance_guards_false': oc.unguarded,
                'peak_source': 'FFMA micro-benchmark (tegb_fma_peak) in this run',
                'note': 'reference-program operations (SURVEY 8(d) rule, guards false) over the measured time: culled '
                        'instances skip that work, so this is a lower bound for guarded programs and can exceed 1 where '
                        'culling resolves most instances'}
    return {'config': config, 'program': scene, 'num_samples': ns, 'instances': count, 'team': mode.pick_team(count, arr),
            'ms_per_launch': ms, 'instances_per_s': count / (ms * 1e-3), 'l2': 'working set > L2' if big else 'L2 flushed before each launch',
            'roofline': roof, 'hbm_GBps_streams': io_bytes * count / (ms * 1e-3) / 1e9,
            'cpu_baseline': {'value': m / cpu_s, 'unit': 'instances/s', 'cores': threads, 'kind': kind,
                             'sample': f'{m} evenly spaced instances of the batch'},
            'speedup_vs_cpu': (count / (ms * 1e-3)) / (m / cpu_s),
            'bit_exact': exact, 'max_scaled_err': float(np.nanmax(np.abs(got - want) / scale)) if want.size else 0.0}


def run_suite(ctx: Ctx, peak_tflops: float):
    out = []
    plan = [
        ('cfg1_readme_1Mi', 'readme_deriv', 50, dict(n=1 << 20, bound='hbm')),
        ('cfg1_readme_64Mi_hbm', 'readme_deriv', 50, dict(n=1 << 26, bound='hbm')),
        ('cfg1_readme_1Mi', 'readme_primal', 50, dict(n=1 << 20)),
        ('cfg2_triangle_256', 'tri_primal', 16, dict(res=(256, 256))),
        ('cfg2_triangle_256', 'tri_grad', 16, dict(res=(256, 256))),
        ('cfg3_shaded_1024', 'shaded_primal', 16, dict(res=(1024, 1024))),
        ('cfg3_shaded_1024', 'shaded_grad', 16, dict(res=(1024, 1024))),
        ('cfg4_nested2d_N1024', 'nested2d_value', 1024, dict(n=4096, cpu_instances=64)),
        ('cfg4_nested2d_N1024', 'nested2d_grad', 1024, dict(n=4096, cpu_instances=1024)),
        ('cfg4_nested3d_N128', 'nested3d_value', 128, dict(n=4096, cpu_instances=32)),
    ]
    for config, scene, ns, kw in plan:
        try:
            out.append(suite_entry(ctx, peak_tflops, config, scene, ns, **kw))
        except Exception as e:          # a suite entry must not take the headline down with it
            out.append({'config': config, 'program': scene, 'error': f'{type(e).__name__}: {e}'})
    return out


# ----------------------------------------------------------------------------------- existing kernel
def existing_kernel(ctx: Ctx, tri: TriRaster):
    """The reference's own ``--target CUDA_C`` text (teg/__main__.py:100-109) in a thread-per-pixel wrapper, compiled by
    nvcc with its defaults (oracle/build_existing_kernel.py -> oracle/_ref/existing_*.cubin): the kernel a Teg
    maintainer would get today, launched through the same C ABI on the same pixel grid, timed like roofline_dense."""
    from teg_b200 import runtime
    from teg_b200.codegen import KernelSource
    torch = ctx.torch
    out = {}
    res = tri.res
    for scene in ('tri_primal', 'tri_grad'):
        cub = os.path.join(ROOT, 'oracle', '_ref', f'existing_{scene}_f32_N{NUM_SAMPLES}.cubin')
        meta = cub[:-6] + '.json'
        if not (os.path.exists(cub) and os.path.exists(meta)):
            return {'unavailable': 'oracle/_ref/existing_*.cubin not built (needs /root/reference at build time)'}
        with open(meta) as f:
            mt = json.load(f)
        with open(cub, 'rb') as f:
            mod = runtime.Module.from_cubin(f.read())
        k = int(mt['n_outputs'])
        ks = KernelSource(name=scene, ctype='float', num_samples=NUM_SAMPLES, source='', uniform_kernel=f'{scene}_uniform',
                          main_kernel=f'{scene}_main', n_uniform=int(mt['n_scalar_args']),
                          scalar_args=list(range(4, 4 + int(mt['n_scalar_args']))), array_args=[], grid_args=[0, 1, 2, 3],
                          n_outputs=k, block_size=int(mt['block_size']), reduce_outputs=0, team=1, tile_rows=1)
        prog = runtime.Program(mod, ks, device=ctx.local)
        imgs = torch.empty((k, res, res), dtype=torch.float32, device=ctx.dev)
        from teg_b200.workloads import TRIANGLE
        scal = [TRIANGLE[lab.rsplit('_', 1)[0]] for lab in mt['scalar_args']]
        grid = runtime.Grid(0.0, 1.0, 0.0, 1.0, res, res, 0, res)
        ptrs = [imgs[j].data_ptr() for j in range(k)]
        sp = ctx.stream.cuda_stream
        ms = gpu_time_ms(ctx, lambda: prog.render(scal, grid, ptrs, stream=sp), reps=5)
        torch.cuda.synchronize(ctx.dev)
        rec = {'ms': ms, 'evals_per_s': res * res / (ms * 1e-3), 'ptxas': mt.get('ptxas'), 'flags': mt.get('flags'),
               'reference_cuda_c_lines': mt.get('lines_of_reference_cuda_c')}
        if scene == 'tri_primal':
            ours = tri.imgs[0][0]
            diff = (imgs[0] - ours).abs()
            rec['pixels_differing_from_ours'] = int((diff > 0).sum().item())
            rec['max_abs_diff_vs_ours'] = float(diff.max().item())
            rec['sum_value'] = float(imgs[0].double().sum().item())
        else:
            rec['grad_sum'] = imgs.double().sum(dim=(1, 2)).cpu().numpy().tolist()
            rec['outputs'] = 'six per-pixel gradient images (the wrapper has no in-kernel reduction)'
        out[scene] = rec
    return out


# ------------------------------------------------------------------------------------------ d2h probe
def d2h_probe(ctx: Ctx, nbytes: int):
    """What bounds the end-to-end figure: every rank copies one image (plain cudaMemcpyAsync through torch) from its GPU
    to pinned host memory, all ranks at the same time.  Per-rank GB/s (min over ranks) and the aggregate."""
    torch = ctx.torch
    src = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    ctx.sync_all()
    reps = 6
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(ctx.dev)
    dt = time.perf_counter() - t0
    mine = nbytes * reps / dt / 1e9
    slowest = -ctx.max_over_ranks(-mine)
    agg = ctx.sum_over_ranks(mine)
    return {'bytes_per_copy': nbytes, 'ranks_copying_concurrently': ctx.world, 'GBps_per_rank_min': slowest,
            'GBps_aggregate': agg, 'how': 'tensor.copy_(device tensor, non_blocking=True) into pinned memory, 6 copies per '
            'rank back to back, all ranks concurrently after a barrier'}


# ------------------------------------------------------------------------------------------- our arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='tri_raster_4096', choices=['tri_raster_4096', 'multi_tri_4096'],
                    help='the headline of the JSON line (the other measurements ride along under "workloads")')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--res', type=int, default=4096)
    ap.add_argument('--cpu-rows', type=int, default=0, help='pixel rows in the bounded CPU sample (0: sized from a probe)')
    ap.add_argument('--cpu-seconds', type=float, default=2.5, help='target duration of one reference-arm step')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--fused', type=int, default=0, help='1: value and reverse derivative as ONE program bundle (one kernel)')
    ap.add_argument('--graph', type=int, default=1, help='1: the timed steps replay ONE captured CUDA graph of the step')
    ap.add_argument('--no-extras', action='store_true', help='headline only: no "workloads" records (profiling runs)')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'], help='working precision of the programs')
    ap.add_argument('--maxreg', type=int, default=0, help='experiment: --maxrregcount for the generated kernels')
    ap.add_argument('--min-blocks', type=int, default=-1, help='experiment: __launch_bounds__(block, N) for render kernels')
    ap.add_argument('--block', type=int, default=128, help='experiment: threads per block of the generated kernels')
    ap.add_argument('--tile-rows', type=int, default=16, help='pixel rows per render block / classification tile')
    ap.add_argument('--coop-groups', type=int, default=4, help='experiment: owners served per cooperative pass, at most')
    ap.add_argument('--allreduce', default='peer', choices=['peer', 'nccl'], help='N > 1: gradient all-reduce fused '
                    'into the column-sum kernel over NVLink peer memory (default), or a separate NCCL all-reduce')
    ap.add_argument('--shard', default='interleave', choices=['interleave', 'band'],
                    help='N > 1: how the pixel rows are split over the ranks')
    ap.add_argument('--cull', type=int, default=1, help='0: dense kernels (every comparison at every sample); '
                    '1: domain culling (teg_b200/cull.py, bit-identical results)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        return run_reference_arm(args)

    from teg_b200 import CUDA_EvalMode, runtime
    from teg_b200.opcount import count as op_count

    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    res = args.res
    peak_hbm, hbm_src = hbm_peak()
    sampler = ClockSampler(ctx.local)

    # ------------------------------------------------------------------ headline
    if args.workload == 'tri_raster_4096':
        head = TriRaster(ctx, args.scaling, args.allreduce, fused=bool(args.fused))
        for _ in range(args.warmup):
            head.step(False)
        ctx.sync_all()
        if rank == 0:
            sampler.start()
        hm = head.measure(args.steps, args.warmup, use_graph=bool(args.graph))
        # the timed region lasts only milliseconds: keep the same steps running (untimed) for ~1.2 s so that nvidia-smi
        # (100 ms period) sees the clocks and throttle reasons under this very load.  The number of extra steps is
        # derived from the all-reduced time, i.e. it is the same on every rank (the steps contain a collective).
        n_hold = int(min(60000, max(20, 1200.0 / max(hm['ms_per_step'], 1e-3))))
        for k in range(n_hold):
            head.step(False)
            if k % 64 == 63:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks['window'] = f'timed steps + {n_hold} untimed repetitions of the same step (~1.2 s)'
        check = head.check()
        tri = head
    else:
        head = MultiTri(ctx, args.scaling)
        for _ in range(args.warmup):
            head.step_launches(False)
        ctx.sync_all()
        if rank == 0:
            sampler.start()
        hm = head.measure(args.steps, args.warmup, bool(args.graph))
        n_hold = int(min(5000, max(20, 1200.0 / max(hm['ms_per_step'], 1e-3))))
        for k in range(n_hold):
            head.step(False)
        torch.cuda.synchronize(dev)
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks['window'] = f'timed steps + {n_hold} untimed repetitions of the same step (~1.2 s)'
        check = head.check()
        tri = None
    ms_per_step, value = hm['ms_per_step'], hm['value']

    # ---------------------------------------------------------------- e2e: host buffers, copies inside
    e2e, e2e_grad = None, None
    if not args.no_e2e and tri is not None:
        tdt, esz = tri.tdt, tri.esz
        stream = ctx.stream
        h_imgs = [torch.empty((tri.n_local_rows, res), dtype=tdt).pin_memory() for _ in range(2)]
        h_gs = [torch.empty(6, dtype=tdt).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        done = [torch.cuda.Event() for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]

        def e2e_steps(k_steps: int, image: bool):
            """Every step: inputs (6 vertex coordinates, by value in the launch parameters; the pixel grid is index
            arithmetic) go to the device, the image and the gradient vector come back to pinned host memory.  Two
            device buffers: the copy of step k overlaps the kernels of step k+1."""
            for k in range(k_steps):
                b = k & 1
                if k >= 2:
                    stream.wait_event(copied[b])        # buffer b is free again
                tri.step(False, b)
                done[b].record(stream)
                copy_stream.wait_event(done[b])
                with torch.cuda.stream(copy_stream):
                    if image:
                        h_imgs[b].copy_(tri.imgs[b][0], non_blocking=True)
                    h_gs[b].copy_(tri.gsums[b], non_blocking=True)
                    copied[b].record(copy_stream)
            copy_stream.synchronize()
            stream.synchronize()

        def run_e2e(image: bool):
            e2e_steps(3, image)
            ctx.sync_all()
            k_e2e = max(4, min(args.steps, 20))
            t0 = time.perf_counter()
            e2e_steps(k_e2e, image)
            ctx.sync_all()
            dt = ctx.max_over_ranks(time.perf_counter() - t0)
            return k_e2e, dt

        k_e2e, dt = run_e2e(True)
        assert abs(float(h_imgs[(k_e2e - 1) & 1].double().sum()) - float(tri.imgs[(k_e2e - 1) & 1].double().sum())) < 1e-9
        n_all = hm['instances_all_ranks']
        e2e = {'value': n_all * k_e2e / dt, 'unit': 'integral evals/s', 'h2d_bytes_per_step': 2 * 6 * esz,
               'd2h_bytes_per_step': int(h_imgs[0].numel() * esz + h_gs[0].numel() * esz),
               'api': 'CUDA_EvalMode.prepare_render(...).run() for the value image and for the gradient (reduce=True): '
                      'one tegb_render_launch_strided C-ABI call each (pixel boxes generated on device); image + '
                      'gradient vector copied to pinned host memory every step on a copy stream, double-buffered',
               'steps': k_e2e, 'ms_per_step': dt / k_e2e * 1e3,
               'd2h_GBps_per_rank': h_imgs[0].numel() * esz * k_e2e / dt / 1e9}
        # the caller pattern of an optimisation loop (README.md:3, reverse_deriv's out_deriv_vals) needs the gradient
        # (and a loss) on the host, not the image: the same step with only the [6] vector copied back
        k2, dt2 = run_e2e(False)
        e2e_grad = {'value': n_all * k2 / dt2, 'unit': 'integral evals/s', 'h2d_bytes_per_step': 2 * 6 * esz,
                    'd2h_bytes_per_step': int(h_gs[0].numel() * esz), 'steps': k2, 'ms_per_step': dt2 / k2 * 1e3,
                    'what': 'same step; the image stays on the device, only the all-reduced gradient vector is copied to '
                            'pinned host memory every step'}

    if not args.no_e2e and tri is None:
        # multi_tri headline end to end: the step through MultiTriangleRasterizer.forward / backward (grouped C-ABI
        # launches), the triangle parameters go to the device and the image band + the [K, 7] gradient come back to
        # pinned host memory, every step
        mt = head
        h_img = torch.empty((mt.img.shape[1], res), dtype=torch.float32).pin_memory()
        h_g = torch.empty((mt.K, 7), dtype=torch.float32).pin_memory()
        h_par = torch.as_tensor(np.concatenate([mt.scene['verts'], mt.scene['col'][:, None]], axis=1),
                                dtype=torch.float32).pin_memory()
        d_par = torch.empty((mt.K, 7), dtype=torch.float32, device=dev)

        def mt_e2e(k_steps):
            for _ in range(k_steps):
                d_par.copy_(h_par, non_blocking=True)
                mt.r.update_params(d_par[:, :6], d_par[:, 6].contiguous())
                mt.step(False)
                h_img.copy_(mt.img[0], non_blocking=True)
                h_g.copy_(mt.g, non_blocking=True)
            torch.cuda.synchronize(dev)
        mt_e2e(2)
        ctx.sync_all()
        k_e2e = max(4, min(args.steps, 10))
        t0 = time.perf_counter()
        mt_e2e(k_e2e)
        ctx.sync_all()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {'value': hm['pairs_all_ranks'] * k_e2e / dt, 'unit': 'integral evals/s', 'h2d_bytes_per_step': int(h_par.numel() * 4),
               'd2h_bytes_per_step': int(h_img.numel() * 4 + h_g.numel() * 4), 'steps': k_e2e, 'ms_per_step': dt / k_e2e * 1e3,
               'api': 'MultiTriangleRasterizer.update_params / forward / backward (tegb_render_groups_* C-ABI calls, one CUDA '
                      'graph replay per step); parameters copied in, image band + [K, 7] gradient copied to pinned host '
                      'memory every step'}

    # ----------------------------------------------------------------------------------- other workloads
    workloads = {}
    if not args.no_extras:
        k_x = max(5, min(args.steps, 10))
        d2h = d2h_probe(ctx, res * res * 4)
        workloads['d2h_probe'] = d2h
        if tri is not None:
            try:
                mt = MultiTri(ctx, 'weak')
                rec = mt.measure(k_x, 3, bool(args.graph))
                rec['check'] = mt.check()
                rec['scaling'] = 'weak'
                rec['allreduce'] = 'fused into the per-triangle column sums over NVLink peer memory' if mt.peers is not None \
                    else ('none (1 GPU)' if world == 1 else f'NCCL ({mt.peer_note})')
                mt_weak = mt
                workloads['multi_tri_4096'] = rec
            except Exception as e:
                workloads['multi_tri_4096'] = {'error': f'{type(e).__name__}: {e}'}
                mt_weak = None
            if world > 1:
                if mt_weak is not None:
                    mt_weak.close()
                    mt_weak = None
                for name, ctor in (('tri_raster_4096_strong', lambda: TriRaster(ctx, 'strong', args.allreduce)),
                                   ('multi_tri_4096_strong', lambda: MultiTri(ctx, 'strong'))):
                    try:
                        w = ctor()
                        rec = w.measure(k_x, 3, bool(args.graph))
                        rec['check'] = w.check()
                        rec['scaling'] = 'strong'
                        rec['res_total'] = [res, res]
                        w.close()
                        workloads[name] = rec
                    except Exception as e:
                        workloads[name] = {'error': f'{type(e).__name__}: {e}'}
        else:
            mt_weak = None
    else:
        mt_weak = None

    if rank != 0:
        if tri is not None:
            tri.close()
        if mt_weak is not None:
            mt_weak.close()
        if args.workload != 'tri_raster_4096':
            head.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------------------------- roofline
    esz = 4 if args.dtype == 'f32' else 8
    peak_tflops, clk_mhz = runtime.fma_peak(esz, ctx.local)
    nominal = 2 * (128 if esz == 4 else 64) * 148 * 1965e6 / 1e12
    traffic_tab = {}
    try:
        import glob
        latest = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_traffic.json')))[-1]
        with open(latest) as f:
            traffic_tab = json.load(f)
        traffic_src = os.path.relpath(latest, ROOT)
    except Exception:
        traffic_src = None

    def traffic_of(key):
        tj = traffic_tab.get(key)
        return None if (tj is None or esz != 4) else tj['dram_bytes_read'] + tj['dram_bytes_write']

    roofline, roofline_dense, existing = None, None, None
    stream = ctx.stream
    sptr = stream.cuda_stream
    if tri is not None:
        # The reference program's operation count (SURVEY 8(d) rule) is defined on the un-culled program.
        sch = tri.m_primal.reference_schedule([0, 1, 2, 3])
        oc = op_count(sch, NUM_SAMPLES)
        n_local = tri.n_local
        flops_per_launch = oc.unguarded * n_local           # primal has no guards: exact
        # dense leg (always measured, same run): the value kernel that evaluates every comparison at every sample --
        # the FP32 roofline statement of BASELINE.json's north_star
        m_dense = tri.m_primal if not args.cull else CUDA_EvalMode(tri.primal, num_samples=NUM_SAMPLES, device=ctx.local,
                                                                 float_width=tri.fw, cull=False)
        grid = (tri.res_x_total, res)
        dense_ms = gpu_time_ms(ctx, lambda: m_dense.render(tri.box, grid, tri.bounds, tri.params, rows=tri.rows,
                                                            out=tri.imgs[1], stream=sptr, interleave=tri.inter),
                               reps=max(5, min(args.steps, 20)))
        tri.pr_primal[0].run(sptr)
        torch.cuda.synchronize(dev)
        dense_equal = bool(torch.equal(tri.imgs[1], tri.imgs[0]))
        assert dense_equal, 'dense and culled value images differ'
        dense_achieved = flops_per_launch / (dense_ms * 1e-3) / 1e12
        roofline_dense = {
            'bound': 'fp32_fma' if esz == 4 else 'fp64_fma', 'kernel': 'tri_primal_main (cull=0)', 'achieved': dense_achieved,
            'peak': peak_tflops, 'unit': 'TFLOP/s', 'frac': dense_achieved / peak_tflops, 'traffic': traffic_of('tri_primal_main_dense'),
            'peak_source': 'FFMA micro-benchmark (tegb_fma_peak) measured in this run; MEASURED_PEAKS.json has no fp32 figure',
            'peak_nominal': nominal, 'algorithmic_ops_per_instance': oc.unguarded, 'ops_per_inner_sample': oc.per_sample_inner,
            'kernel_ms': dense_ms, 'image_equals_culled_image': dense_equal,
            'note': 'achieved counts the reference program\'s operations (SURVEY 8(d) rule: 17 per inner sample); ptxas '
                    'tabulates the inner-variable-only terms, so fewer instructions execute and the fraction can exceed 1; '
                    'the kernel is bound by the ALU pipe (3 FSETP per sample; ncu pipe figures and SASS under profiles/)'}
        if args.cull:
            # default path: after culling almost no sample is evaluated; what bounds the value kernel is its output
            # stream.  Timed alone (in the step it shares the GPU with the gradient pipeline), L2 flushed: the main
            # kernel itself between two CUDA events the runtime records around its launch on the launching stream
            # (tegb_program_enable_timing), and the whole call (classification + main kernel) between events here.
            prm = tri.pr_primal[1]
            prm.prog.enable_timing(True)
            reps = max(5, min(args.steps, 20))
            main_ms = []
            for k in range(3 + reps):
                ctx.flush.fill_(k & 0xff)
                prm.run(sptr)
                if k >= 3:
                    main_ms.append(prm.prog.main_kernel_ms())
            prm.prog.enable_timing(False)
            kern_ms = float(np.mean(main_ms))
            call_ms = gpu_time_ms(ctx, lambda: prm.run(sptr), reps=reps)
            bytes_per_launch = n_local * esz
            ach = bytes_per_launch / (kern_ms * 1e-3) / 1e9
            roofline = {
                'bound': 'hbm', 'kernel': 'tri_primal_main (cull=1)', 'achieved': ach, 'peak': peak_hbm,
                'unit': 'GB/s', 'frac': ach / peak_hbm, 'traffic': traffic_of('tri_primal_main'),
                'traffic_unit': 'bytes per launch (dram read + write, ncu)', 'traffic_source': traffic_src,
                'algorithmic_bytes': bytes_per_launch, 'peak_source': hbm_src, 'kernel_ms': kern_ms,
                'launches_timed': len(main_ms), 'call_ms_classification_plus_main': call_ms,
                'frac_with_classification': bytes_per_launch / (call_ms * 1e-3) / 1e9 / peak_hbm,
                'algorithmic_tflops': flops_per_launch / (kern_ms * 1e-3) / 1e12,
                'algorithmic_frac_of_fp32_peak': flops_per_launch / (kern_ms * 1e-3) / 1e12 / peak_tflops,
                'note': 'culled kernels (teg_b200/cull.py) evaluate the comparisons only where an edge crosses a pixel: '
                        'the algorithmic bytes are the value image (one store per pixel, no input stream); kernel_ms is '
                        'the main kernel alone (CUDA events recorded by the runtime around its launch, L2 flushed before '
                        'each call), call_ms adds the tile classification kernel and the launch gap; the image fits the '
                        '126 MB L2, so DRAM traffic (ncu) is below the algorithmic bytes: the bound is the L2 write path, '
                        'a plain fill of the same image takes ~15 us; algorithmic_tflops applies the reference program\'s '
                        'operation count to this time (the culled kernel skips that work, so the figure exceeds the FFMA '
                        'peak); the FP32 statement for the sample loop itself is roofline_dense'}
        else:
            roofline = dict(roofline_dense, kernel='tri_primal_main')
        roofline['hbm_GBps_out'] = (n_local * esz) / (ms_per_step * 1e-3) / 1e9
        if world == 1 and not args.no_extras and esz == 4:
            try:
                existing = existing_kernel(ctx, tri)
                if 'tri_primal' in existing:
                    ek = existing['tri_primal']['ms']
                    existing['ratios'] = {'dense_vs_existing_value_kernel': ek / dense_ms,
                                          'culled_vs_existing_value_kernel': ek / roofline['kernel_ms'],
                                          'step_vs_existing_value_plus_grad': (ek + existing['tri_grad']['ms']) / ms_per_step}
            except Exception as e:
                existing = {'error': f'{type(e).__name__}: {e}'}
    else:
        # multi_tri headline: the forward kernels against the FP32 roofline (reference-program operations per pair)
        sch = head.r.m_primal.reference_schedule([0, 1, 2, 3])
        oc = op_count(sch, NUM_SAMPLES)
        fwd_ms = hm['phase_ms']['forward']
        ach = oc.unguarded * head.pairs_local / (fwd_ms * 1e-3) / 1e12
        roofline = {'bound': 'fp32_fma', 'kernel': 'tri_color_primal_gather (one launch over the image, every block walks the triangles of its bin)', 'achieved': ach,
                    'peak': peak_tflops, 'unit': 'TFLOP/s', 'frac': ach / peak_tflops, 'traffic': None,
                    'algorithmic_ops_per_instance': oc.unguarded, 'kernel_ms': fwd_ms,
                    'note': 'reference-program operations over the forward phase; culled kernels skip most of them'}
    if 'multi_tri_4096' in workloads and 'error' not in workloads['multi_tri_4096'] and mt_weak is not None:
        sch = mt_weak.r.m_primal.reference_schedule([0, 1, 2, 3])
        oc5 = op_count(sch, NUM_SAMPLES)
        rec = workloads['multi_tri_4096']
        fwd_ms = rec['phase_ms']['forward']
        ach = oc5.unguarded * mt_weak.pairs_local / (fwd_ms * 1e-3) / 1e12
        rec['roofline'] = {'bound': 'fp32_fma', 'kernel': 'tri_color_primal_gather (one launch over the image, every block walks the triangles of its bin)',
                           'achieved': ach, 'peak': peak_tflops, 'unit': 'TFLOP/s', 'frac': ach / peak_tflops,
                           'traffic': None, 'algorithmic_ops_per_instance': oc5.unguarded, 'kernel_ms': fwd_ms,
                           'note': 'reference-program operations over the forward phase; culled kernels skip most of them'}

    # ----------------------------------------------------------------------------------- cpu baseline
    cpu = None
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        if tri is not None:
            kind, run = cpu_runner(threads)
            inputs, n_inst, rows = cpu_sample_inputs(res, res, 8 if kind == 'port' else 256)

            def both(r, x):
                r('tri_primal', x)
                r('tri_grad', x)
            both(run, inputs)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                both(run, inputs)
                best = min(best, time.perf_counter() - t0)
            kind1, run1 = cpu_runner(1)
            inputs1, n1, _ = cpu_sample_inputs(res, res, max(8, rows // 16))
            t0 = time.perf_counter()
            both(run1, inputs1)
            best1 = time.perf_counter() - t0
            cpu = {'value': n_inst / best, 'unit': 'integral evals/s', 'cores': threads, 'kind': kind,
                   'sample': f'{rows} evenly spaced pixel rows of the {res}x{res} grid ({n_inst} instances), best of 3',
                   'value_1core': n1 / best1}
        else:
            kind, run = cpu_runner(threads, ('tri_color_primal', 'tri_color_grad'))
            cols, n_s = multi_tri_cpu_sample(head.scene, res, 8 if kind == 'reference' else 64, 4)
            run('tri_color_primal', cols[1:])
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                run('tri_color_primal', cols[1:])
                run('tri_color_grad', cols)
                best = min(best, time.perf_counter() - t0)
            cpu = {'value': n_s / best, 'unit': 'integral evals/s', 'cores': threads, 'kind': kind,
                   'sample': f'every 8th triangle, every 4th row of its window ({n_s} pixel-triangle pairs), best of 3'}
        if 'multi_tri_4096' in workloads and 'error' not in workloads['multi_tri_4096'] and mt_weak is not None:
            kind, run = cpu_runner(threads, ('tri_color_primal', 'tri_color_grad'))
            cols, n_s = multi_tri_cpu_sample(mt_weak.scene, res, 8 if kind == 'reference' else 64, 4)
            run('tri_color_primal', cols[1:])
            t0 = time.perf_counter()
            run('tri_color_primal', cols[1:])
            run('tri_color_grad', cols)
            dt = time.perf_counter() - t0
            workloads['multi_tri_4096']['cpu_baseline'] = {
                'value': n_s / dt, 'unit': 'integral evals/s', 'cores': threads, 'kind': kind,
                'sample': f'every 8th triangle, every 4th row of its window ({n_s} pixel-triangle pairs)'}
    if world == 1 and not args.no_extras and not args.no_cpu and esz == 4:
        workloads['suite'] = run_suite(ctx, peak_tflops)
    if existing is not None:
        workloads['existing_kernel'] = existing

    peers_used = (tri is not None and tri.peers is not None) or (tri is None and head.peers is not None)
    allreduce_txt = ('fused into the column-sum kernel over NVLink peer memory' if peers_used else
                     ('none (1 GPU)' if world == 1 else 'NCCL'))
    line = {
        'metric': 'integral_evals_per_s', 'value': value, 'unit': 'integral evals/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': headline_config(args.workload, res, args.scaling),
        'impl_config': {'cull': bool(args.cull),
                        'outputs': 'value image + per-parameter gradient sums (reduced in-kernel)',
                        'l2': 'L2 flushed between timed steps by writing a 256 MB buffer (outside the per-step event '
                              'pairs); inputs are generated from the pixel index',
                        'parallelism': f'rows sharded over {world} GPU(s) ({args.shard if world > 1 else "all"}), one '
                                       f'all-reduce of the gradient per step ({allreduce_txt})'},
        'allreduce_note': (tri.peer_note if tri is not None else head.peer_note), 'phase_ms': hm['phase_ms'],
        'phase_note': hm.get('phase_note'), 'cuda_graph': hm.get('cuda_graph'),
        'roofline': roofline, 'roofline_dense': roofline_dense, 'cpu_baseline': cpu, 'e2e': e2e, 'e2e_grad_only': e2e_grad,
        'gpu_launches': int(hm['gpu_launches']), 'clocks': clocks, 'check': check, 'workloads': workloads,
    }
    print(json.dumps(line))
    if tri is not None:
        tri.close()
    if mt_weak is not None:
        mt_weak.close()
    if args.workload != 'tri_raster_4096':
        head.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
