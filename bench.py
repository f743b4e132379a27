#!/usr/bin/env python
"""bench.py -- integral evaluations / s for the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, sm_100a)
    python bench.py --impl reference ...                    # the reference's C backend on the host cores

Workload (config.workload = "tri_raster_4096"): the per-pixel triangle rasterization integral plus its
reverse derivative w.r.t. the six vertex coordinates (oracle/scenes.py: tri_primal + tri_grad, the
program of SURVEY.md §8(d) config 2) on a 4096 x 4096 pixel grid per GPU, 16 midpoint samples per
integral (256 samples per pixel for the value, 16 per active delta term for the gradient).
One *step* = one pass over the whole batch: value image, per-pixel gradients, and the per-parameter
gradient sum (+ one NCCL all-reduce of the 6-vector when N > 1).  One *integral evaluation* = one
pixel instance (value + its reverse derivative).  Multi-GPU is weak scaling: the image grows to
(4096 N) x 4096 pixels over the same domain and rank r owns the tile rows r, r + N, ... (16 pixel rows each;
--shard band gives it the contiguous rows [4096 r, 4096 (r+1)) instead).

Nothing here reads /root/reference: programs come from tests/golden/programs/*.tegp, the CPU
baseline from oracle/_ref/*.so (the reference's own generated C, built where the reference exists)
or, failing that, from the oracle port.
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
WORKLOAD = 'tri_raster_4096'


def load_text(scene: str) -> str:
    with open(os.path.join(PROGRAM_DIR, f'{scene}.tegp')) as f:
        return f.read()


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
def cpu_runner(threads: int):
    """Returns (kind, run(scene, inputs) -> ndarray).  'reference' = oracle/_ref (the reference's own C)."""
    from oracle import Oracle, RefLib, ref_lib_path
    if all(os.path.exists(ref_lib_path(s, np.float32, NUM_SAMPLES)) for s in ('tri_primal', 'tri_grad')):
        libs = {s: RefLib(s, np.float32, NUM_SAMPLES) for s in ('tri_primal', 'tri_grad')}
        return 'reference', lambda scene, inputs: libs[scene].eval(inputs, threads=threads)
    libs = {s: Oracle(load_text(s)) for s in ('tri_primal', 'tri_grad')}
    return 'port', lambda scene, inputs: libs[scene].eval(inputs, NUM_SAMPLES, np.float32, threads=threads)


def cpu_sample_inputs(res: int, n_rows: int):
    """Bounded sample of the workload: n_rows pixel rows spread evenly over the res x res grid."""
    from teg_b200.tegp import Program
    from teg_b200.workloads import TRIANGLE, bind, pixel_boxes
    prog = Program.loads(load_text('tri_primal'))
    stride = max(1, res // n_rows)
    rows = np.arange(0, res, stride)[:n_rows]
    parts = [pixel_boxes((res, res), rows=(int(r), int(r) + 1)) for r in rows]
    box = [np.concatenate([p[k] for p in parts]) for k in range(4)]
    vals = dict(zip(('x_lb', 'x_ub', 'y_lb', 'y_ub'), box), **TRIANGLE)
    return bind(prog.args, vals), len(rows) * res


def time_cpu(run, inputs, reps: int = 3) -> float:
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        run('tri_primal', inputs)
        run('tri_grad', inputs)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference_arm(args) -> int:
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    kind, run = cpu_runner(threads)
    if kind == 'port':          # the interpreter is ~50x slower than compiled C: keep the sample bounded
        args.cpu_rows = min(args.cpu_rows, 8)
    inputs, n_inst = cpu_sample_inputs(args.res, args.cpu_rows)
    for _ in range(max(args.warmup, 1)):
        run('tri_primal', inputs)
        run('tri_grad', inputs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run('tri_primal', inputs)
        run('tri_grad', inputs)
    dt = time.perf_counter() - t0
    value = n_inst * args.steps / dt
    sample = f'{args.cpu_rows} evenly spaced pixel rows of the {args.res}x{args.res} grid ({n_inst} instances) per step'
    line = {
        'impl': 'reference', 'metric': 'integral_evals_per_s', 'value': value, 'unit': 'integral evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'res': args.res, 'num_samples': NUM_SAMPLES,
                   'programs': ['tri_primal', 'tri_grad']},
        'cpu_baseline': {'value': value, 'unit': 'integral evals/s', 'cores': threads, 'kind': kind,
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'integral evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------- our arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--res', type=int, default=4096)
    ap.add_argument('--cpu-rows', type=int, default=256, help='pixel rows in the bounded CPU sample')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--fused', type=int, default=0, help='1: value and reverse derivative in ONE kernel (program bundle)')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'], help='working precision of the programs')
    ap.add_argument('--maxreg', type=int, default=0, help='experiment: --maxrregcount for the generated kernels')
    ap.add_argument('--min-blocks', type=int, default=-1, help='experiment: __launch_bounds__(block, N) for render kernels')
    ap.add_argument('--block', type=int, default=128, help='experiment: threads per block of the generated kernels')
    ap.add_argument('--tile-rows', type=int, default=16, help='pixel rows per render block / classification tile')
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

    import torch
    import torch.distributed as dist
    from teg_b200 import CUDA_EvalMode, runtime
    from teg_b200.opcount import count as op_count
    from teg_b200.tegp import Program
    from teg_b200.workloads import TRIANGLE

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the CUDA backend has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        with _StdoutToStderr():
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()              # the communicator exists (and has said so) before stdout is restored

    res = args.res
    res_x_total = res * world                  # weak scaling: the image grows to (res * world) x res pixels
    # --shard band: rank r owns rows [res*r, res*(r+1)); --shard interleave (default): rank r owns the tile rows
    # r, r + world, ... (16 pixel rows each), so every rank gets the same mix of interior / exterior / edge tiles
    if args.shard == 'band' or world == 1:
        rows, inter = (res * rank, res * (rank + 1)), (1, 0)
    else:
        assert res % args.tile_rows == 0, 'interleaved sharding wants res divisible by the tile height'
        rows, inter = (0, res_x_total), (world, rank)
    n_local = res * res
    primal = Program.loads(load_text('tri_primal'))
    grad = Program.loads(load_text('tri_grad'))
    fw = 'float' if args.dtype == 'f32' else 'double'
    tdt = torch.float32 if args.dtype == 'f32' else torch.float64
    esz = 4 if args.dtype == 'f32' else 8
    xopts = dict(nvrtc_opts=f'--maxrregcount={args.maxreg}' if args.maxreg else '', tile_rows=args.tile_rows,
                 block_size=args.block, min_blocks=None if args.min_blocks < 0 else args.min_blocks)
    m_primal = CUDA_EvalMode(primal, num_samples=NUM_SAMPLES, device=local_rank, float_width=fw, cull=bool(args.cull),
                             **xopts)
    m_grad = CUDA_EvalMode(grad, num_samples=NUM_SAMPLES, device=local_rank, float_width=fw, cull=bool(args.cull),
                           **xopts)
    m_fused = CUDA_EvalMode([primal, grad], num_samples=NUM_SAMPLES, device=local_rank, float_width=fw,
                            cull=bool(args.cull)) if args.fused else None
    box = primal.args[:4]
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in primal.args[4:]}
    gparams = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in grad.args[4:]}
    bounds = ((0.0, 1.0), (0.0, 1.0))

    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream
    imgs = [torch.empty((1, res, res), dtype=tdt, device=dev) for _ in range(2)]
    gsums = [torch.zeros(6, dtype=tdt, device=dev) for _ in range(2)]
    img, gsum = imgs[0], gsums[0]

    ev = {k: [] for k in ('primal', 'grad')}

    # bound launches (one C-ABI call each); the fused bundle keeps the generic render() call
    pr_grad = [m_grad.prepare_render(box, (res_x_total, res), bounds, gparams, rows=rows, out=gsums[b], reduce=True,
                                     interleave=inter)
               for b in range(2)]
    pr_primal = [m_primal.prepare_render(box, (res_x_total, res), bounds, params, rows=rows, out=imgs[b], interleave=inter)
                 for b in range(2)]
    # the gradient pipeline is the longer one (and ends in the all-reduce): its blocks get priority
    side = torch.cuda.Stream(dev, priority=-1)
    forks = [torch.cuda.Event() for _ in range(2)]
    joins = [torch.cuda.Event() for _ in range(2)]
    side_ptr = side.cuda_stream
    peers = None
    peer_note = None
    if world > 1 and args.allreduce == 'peer' and m_fused is None:
        from teg_b200.parallel import PeerGroup
        try:
            peers = PeerGroup(local_rank, max_values=16)       # raises on every rank or on none
            for pr in pr_grad:
                peers.attach(pr.prog)        # the last column-sum kernel of the gradient does the exchange itself
        except RuntimeError as e:
            peers, peer_note = None, f'peer mailboxes unavailable ({e}): NCCL all-reduce instead'

    def step(record: bool, buf: int = 0):
        """value image + reverse derivative of every pixel summed per parameter inside the kernels + the one
        all-reduce when N > 1.  The two programs are independent: the gradient pipeline (classification, kernel,
        column sums, all-reduce) runs on a second stream next to the value pipeline, joined at the end."""
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if m_fused is not None:
            if record:
                e[0].record(stream)
            m_fused.render(box, (res_x_total, res), bounds, gparams, rows=rows, out=(imgs[buf], gsums[buf]),
                           stream=sptr, reduce=6, interleave=inter)
            if record:
                e[1].record(stream)
            if world > 1:
                dist.all_reduce(gsums[buf])
            if record:
                ev['primal'].append((e[0], e[1]))
                ev['grad'].append((e[0], e[1]))
            return
        fork, join = forks[buf], joins[buf]
        fork.record(stream)
        side.wait_event(fork)
        if record:
            e[2].record(side)
        pr_grad[buf].run(side_ptr)
        if world > 1 and peers is None:
            with torch.cuda.stream(side):
                dist.all_reduce(gsums[buf])        # enqueued behind the gradient on `side`
        if record:
            e[3].record(side)
        join.record(side)
        if record:
            e[0].record(stream)
        pr_primal[buf].run(sptr)
        if record:
            e[1].record(stream)
        stream.wait_event(join)
        if record:
            ev['primal'].append((e[0], e[1]))
            ev['grad'].append((e[2], e[3]))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    launches0 = runtime.launch_count()
    for _ in range(args.warmup):
        step(False)
    sync_all()
    launches_warm = runtime.launch_count()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    step_ev = []
    sync_all()
    for k in range(args.steps):
        flush.fill_(k & 0xff)                                           # evict L2 (not timed)
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step(True)
        b.record(stream)
        step_ev.append((a, b))
    sync_all()
    total_ms = float(sum(a.elapsed_time(b) for a, b in step_ev))
    launches_timed = runtime.launch_count() - launches_warm
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # the timed region lasts only milliseconds: keep the same steps running (untimed) for ~1.2 s so that nvidia-smi
    # (100 ms period) sees the clocks and throttle reasons under this very load.  The number of extra steps is derived
    # from the all-reduced time, i.e. it is the same on every rank (the steps contain a collective).
    n_hold = int(min(60000, max(20, 1200.0 / max(total_ms / args.steps, 1e-3))))
    for k in range(n_hold):
        step(False)
        if k % 64 == 63:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks['window'] = f'timed steps + {n_hold} untimed repetitions of the same step (~1.2 s)'
    sync_all()
    ms_per_step = total_ms / args.steps
    value = n_local * world * args.steps / (total_ms * 1e-3)

    phase_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ev.items()}

    # sanity: the summed image is the triangle area, the summed gradient its analytic derivative
    area = float(img.double().sum().item())
    gvec = gsum.cpu().numpy().astype(np.float64)

    # ---------------------------------------------------------------- e2e: host buffers, copies inside
    e2e = None
    if not args.no_e2e:
        h_imgs = [torch.empty((res, res), dtype=tdt).pin_memory() for _ in range(2)]
        h_gs = [torch.empty(6, dtype=tdt).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        done = [torch.cuda.Event() for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]

        def e2e_steps(k_steps: int):
            """Every step: inputs (6 vertex coordinates, by value in the launch parameters; the pixel grid is
            index arithmetic) go to the device, the image and the gradient vector come back to pinned host
            memory.  Two device buffers: the copy of step k overlaps the kernels of step k+1."""
            for k in range(k_steps):
                b = k & 1
                if k >= 2:
                    stream.wait_event(copied[b])        # buffer b is free again
                step(False, b)
                done[b].record(stream)
                copy_stream.wait_event(done[b])
                with torch.cuda.stream(copy_stream):
                    h_imgs[b].copy_(imgs[b][0], non_blocking=True)
                    h_gs[b].copy_(gsums[b], non_blocking=True)
                    copied[b].record(copy_stream)
            copy_stream.synchronize()
            stream.synchronize()

        e2e_steps(3)
        sync_all()
        k_e2e = max(4, min(args.steps, 20))
        t0 = time.perf_counter()
        e2e_steps(k_e2e)
        sync_all()
        dt = time.perf_counter() - t0
        assert abs(float(h_imgs[(k_e2e - 1) & 1].double().sum()) - float(imgs[(k_e2e - 1) & 1].double().sum())) < 1e-9
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': n_local * world * k_e2e / float(te.item()), 'unit': 'integral evals/s',
               'h2d_bytes_per_step': 2 * 6 * esz,
               'd2h_bytes_per_step': int(h_imgs[0].numel() * esz + h_gs[0].numel() * esz),
               'api': 'CUDA_EvalMode.prepare_render(...).run() for the value image and for the gradient (reduce=True): '
                      'one tegb_render_launch_strided C-ABI call each (pixel boxes generated on device); image + '
                      'gradient vector copied to pinned host memory every step on a copy stream, double-buffered',
               'steps': k_e2e,
               'ms_per_step': float(te.item()) / k_e2e * 1e3}

    if peers is not None:
        peers.close()                # checks the time-out flag, then a barrier before the mailboxes go away
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------------------------- roofline
    # The reference program's operation count (SURVEY 8(d) rule) is defined on the un-culled program.
    sch = m_primal.reference_schedule([0, 1, 2, 3])
    oc = op_count(sch, NUM_SAMPLES)
    flops_per_launch = oc.unguarded * n_local           # primal has no guards: exact
    peak_tflops, clk_mhz = runtime.fma_peak(esz, local_rank)
    nominal = 2 * (128 if esz == 4 else 64) * 148 * 1965e6 / 1e12
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            hbm_peak, hbm_src = float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)'
    except Exception:
        hbm_peak, hbm_src = 6650.0, 'fallback stated in B200_PROFILING.md (MEASURED_PEAKS.json absent)'
    traffic_tab = {}
    try:
        import glob
        latest = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_traffic.json')))[-1]
        with open(latest) as f:
            traffic_tab = json.load(f)
    except Exception:
        pass

    def traffic_of(key):
        tj = traffic_tab.get(key)
        return None if (tj is None or esz != 4) else tj['dram_bytes_read'] + tj['dram_bytes_write']

    # dense leg (always measured, same run): the value kernel that evaluates every comparison at every sample --
    # the FP32 roofline statement of BASELINE.json's north_star
    m_dense = m_primal if not args.cull else CUDA_EvalMode(primal, num_samples=NUM_SAMPLES, device=local_rank,
                                                           float_width=fw, cull=False)
    dense_ev = []
    for k in range(3 + max(5, min(args.steps, 20))):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(k & 0xff)
        a.record(stream)
        m_dense.render(box, (res_x_total, res), bounds, params, rows=rows, out=imgs[1], stream=sptr, interleave=inter)
        b.record(stream)
        if k >= 3:
            dense_ev.append((a, b))
    torch.cuda.synchronize(dev)
    dense_ms = float(np.mean([a.elapsed_time(b) for a, b in dense_ev]))
    assert torch.equal(imgs[1], imgs[0]), 'dense and culled value images differ'
    dense_achieved = flops_per_launch / (dense_ms * 1e-3) / 1e12
    roofline_dense = {
        'bound': 'fp32_fma' if esz == 4 else 'fp64_fma', 'kernel': 'tri_primal_main (cull=0)', 'achieved': dense_achieved,
        'peak': peak_tflops, 'unit': 'TFLOP/s', 'frac': dense_achieved / peak_tflops, 'traffic': traffic_of('tri_primal_main_dense'),
        'peak_source': 'FFMA micro-benchmark (tegb_fma_peak) measured in this run; MEASURED_PEAKS.json has no fp32 figure',
        'peak_nominal': nominal, 'algorithmic_ops_per_instance': oc.unguarded, 'ops_per_inner_sample': oc.per_sample_inner,
        'kernel_ms': dense_ms, 'launches_timed': len(dense_ev),
        'note': 'achieved counts the reference program\'s operations (SURVEY 8(d) rule: 17 per inner sample); ptxas '
                'tabulates the inner-variable-only terms, so fewer instructions execute and the fraction can exceed 1; '
                'the kernel is bound by the ALU pipe (3 FSETP per sample; ncu: alu 81 %, fma 42 %, issue 87 %; profiles/)'}
    if args.cull:
        # default path: after culling almost no sample is evaluated; what bounds the value kernel is its output stream.
        # Timed alone (in the step it shares the GPU with the gradient pipeline), L2 flushed, CUDA events.
        solo_ev = []
        for k in range(3 + max(5, min(args.steps, 20))):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(k & 0xff)
            a.record(stream)
            pr_primal[1].run(sptr)
            b.record(stream)
            if k >= 3:
                solo_ev.append((a, b))
        torch.cuda.synchronize(dev)
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in solo_ev]))
        bytes_per_launch = n_local * esz
        ach = bytes_per_launch / (kern_ms * 1e-3) / 1e9
        roofline = {
            'bound': 'hbm', 'kernel': 'tri_primal_tiles + tri_primal_main (cull=1)', 'achieved': ach, 'peak': hbm_peak,
            'unit': 'GB/s', 'frac': ach / hbm_peak, 'traffic': traffic_of('tri_primal_main'),
            'traffic_unit': 'bytes per launch (dram read + write, ncu)', 'algorithmic_bytes': bytes_per_launch,
            'peak_source': hbm_src, 'kernel_ms': kern_ms,
            'algorithmic_tflops': flops_per_launch / (kern_ms * 1e-3) / 1e12,
            'algorithmic_frac_of_fp32_peak': flops_per_launch / (kern_ms * 1e-3) / 1e12 / peak_tflops,
            'note': 'culled kernels (teg_b200/cull.py) evaluate the comparisons only where an edge crosses a pixel: '
                    'the algorithmic bytes are the value image (one store per pixel, no input stream); kernel_ms is '
                    'the CUDA-event time of the tile classification + main kernel launched alone (20 launches after '
                    'the timed steps, L2 flushed); algorithmic_tflops applies the '
                    'reference program\'s operation count to this time (the culled kernel skips that work, so the '
                    'figure exceeds the FFMA peak); the FP32 statement for the sample loop itself is roofline_dense'}
    else:
        roofline = dict(roofline_dense, kernel='tri_primal_main')
    roofline['hbm_GBps_out'] = (n_local * esz) / (ms_per_step * 1e-3) / 1e9

    # ----------------------------------------------------------------------------------- cpu baseline
    cpu = None
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        kind, run = cpu_runner(threads)
        if kind == 'port':
            args.cpu_rows = min(args.cpu_rows, 8)
        inputs, n_inst = cpu_sample_inputs(res, args.cpu_rows)
        run('tri_primal', inputs)
        best = time_cpu(run, inputs, reps=3)
        kind1, run1 = cpu_runner(1)
        inputs1, n1 = cpu_sample_inputs(res, max(8, args.cpu_rows // 16))
        best1 = time_cpu(run1, inputs1, reps=2)
        cpu = {'value': n_inst / best, 'unit': 'integral evals/s', 'cores': threads, 'kind': kind,
               'sample': f'{args.cpu_rows} evenly spaced pixel rows of the {res}x{res} grid ({n_inst} instances), '
                         f'best of 3', 'value_1core': n1 / best1}

    line = {
        'metric': 'integral_evals_per_s', 'value': value, 'unit': 'integral evals/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'res_per_gpu': [res, res], 'num_samples': NUM_SAMPLES,
                   'programs': ['tri_primal', 'tri_grad'], 'fused_kernel': bool(args.fused), 'cull': bool(args.cull), 'outputs': 'value image [4096,4096] + per-parameter '
                   'gradient sum [6] (reduced in-kernel)', 'l2': 'L2 flushed between timed steps by writing a '
                   '256 MB buffer (outside the per-step event pairs); inputs are generated from the pixel index',
                   'parallelism': f'rows sharded over {world} GPU(s) ({args.shard if world > 1 else "all"}), one all-reduce of the [6] gradient per step ('
                                  + ('fused into the column-sum kernel over NVLink peer memory' if peers is not None else 'NCCL')
                                  + '), on the gradient stream next to the value kernel'},
        'allreduce_note': peer_note, 'phase_ms': phase_ms, 'phase_note': 'the value and the gradient pipelines run concurrently on two streams: '
        'their phase times overlap and each is longer than the same pipeline alone', 'roofline': roofline, 'roofline_dense': roofline_dense, 'cpu_baseline': cpu, 'e2e': e2e,
        'gpu_launches': int(launches_timed), 'clocks': clocks,
        'check': {'sum_value': area, 'expected_area_per_rank': 0.2325 if world == 1 else None,
                  'grad_sum': gvec.tolist()},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
