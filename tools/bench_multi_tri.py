#!/usr/bin/env python
"""Config 5: 4096 x 4096 per GPU, 256 flat-coloured triangles per band, forward image +
reverse-mode gradients [K, 7] with an upstream dL = image - target, sharded over N GPUs (weak scaling: one
unit square / band of 4096 pixel rows / 256 triangles per GPU), one NCCL all-reduce of the [K, 7] gradient.

    python tools/bench_multi_tri.py --steps 20                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N tools/bench_multi_tri.py --gpus N

One step = forward (9 window-disjoint classes added into the image) + dL (elementwise) + backward (one
grouped launch, per-triangle sums reduced in-kernel) + all-reduce.  An *evaluation* = one (pixel, triangle)
pair inside a triangle's window: its value integral and its reverse derivative.  Prints one JSON line in
bench.py's format (this is a secondary workload; bench.py's default stays the headline).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--res', type=int, default=4096)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cull', type=int, default=1)
    ap.add_argument('--graph', type=int, default=1, help='1: replay the step as one CUDA graph (parameters live in '
                    'device arrays, so the graph stays valid while they change)')
    ap.add_argument('--min-blocks', default='4,4', help='experiment: "a,b" = __launch_bounds__ resident blocks per SM of '
                    'the grouped value / gradient kernels (0 = the compiler\'s choice)')
    ap.add_argument('--tile-rows', default='32,32', help='pixel rows per tile of the value / gradient kernels')
    ap.add_argument('--coop-groups', type=int, default=4)
    ap.add_argument('--gather', type=int, default=1, help='1: forward as ONE launch over the image (every block walks the '
                    'triangles of its bin); 0: one accumulating launch per class of window-disjoint triangles')
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from teg_b200 import runtime
    from teg_b200.opcount import count as op_count
    from teg_b200.raster import MultiTriangleRasterizer
    from teg_b200.tegp import Program
    from teg_b200.workloads import multi_triangle_scene

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    res = args.res
    pdir = os.path.join(ROOT, 'tests', 'golden', 'programs')
    primal = Program.load(os.path.join(pdir, 'tri_color_primal.tegp'))
    grad = Program.load(os.path.join(pdir, 'tri_color_grad.tegp'))
    scene = multi_triangle_scene(bands=world, res=res)
    K = scene['verts'].shape[0]
    mb = tuple((int(x) or None) for x in args.min_blocks.split(',')) if args.min_blocks else (None, None)
    r = MultiTriangleRasterizer(primal, grad, res, bands=world, rank=rank, device=local, cull=bool(args.cull),
                                min_blocks=mb, tile_rows=tuple(int(x) for x in args.tile_rows.split(',')), coop_groups=args.coop_groups, gather=bool(args.gather))
    br = scene['band_ranges']
    active = (br[max(rank - 1, 0)][0], br[min(rank + 1, world - 1)][1])      # own band and its neighbours
    r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'],
                scene['max_cols'], active=active)
    w = scene['windows'].astype(np.int64)
    r0 = np.clip(w[:, 0], res * rank, res * (rank + 1))
    r1 = np.clip(w[:, 1], res * rank, res * (rank + 1))
    pairs_local = int(((r1 - r0) * (w[:, 3] - w[:, 2])).sum())
    pairs_total = int(((w[:, 1] - w[:, 0]) * (w[:, 3] - w[:, 2])).sum())

    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream
    img = torch.empty((1, res, res), dtype=torch.float32, device=dev)
    target = torch.full((res, res), 0.25, dtype=torch.float32, device=dev)
    dL = torch.empty((res, res), dtype=torch.float32, device=dev)
    g = torch.empty((K, 7), dtype=torch.float32, device=dev)
    ev = {'forward': [], 'backward': []}

    def step(record):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            e[0].record(stream)
        r.forward(out=img, stream=sptr)
        if record:
            e[1].record(stream)
        torch.sub(img[0], target, out=dL)
        if record:
            e[2].record(stream)
        r.backward(dL, out=g, stream=sptr)
        if record:
            e[3].record(stream)
        if world > 1:
            dist.all_reduce(g)
        if record:
            ev['forward'].append((e[0], e[1]))
            ev['backward'].append((e[2], e[3]))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step(False)
    sync_all()
    for _ in range(3):
        step(True)              # per-phase times, launch by launch
    sync_all()
    phase = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ev.items()}
    graph = None
    if args.graph:
        def body(sp):
            r.forward(out=img, stream=sp)
            torch.sub(img[0], target, out=dL)
            r.backward(dL, out=g, stream=sp)
        graph = r.capture(body)

        def step(record):           # noqa: F811
            graph.replay()
            if world > 1:
                dist.all_reduce(g)
        for _ in range(3):
            step(False)
        sync_all()
    l0 = runtime.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pairs = []
    sync_all()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step(True)
        b.record(stream)
        pairs.append((a, b))
    sync_all()
    total_ms = float(sum(a.elapsed_time(b) for a, b in pairs))
    clocks = sampler.stop() if rank == 0 else None
    launches = runtime.launch_count() - l0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    coverage = float(img.double().sum().item()) / (res * res)
    gnorm = float(g.double().norm().item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    sch = r.m_primal.reference_schedule([0, 1, 2, 3])
    oc = op_count(sch, 16)
    peak, _ = runtime.fma_peak(4, local)
    achieved = oc.unguarded * pairs_local / (phase['forward'] * 1e-3) / 1e12
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import RefLib, ref_lib_path
        if os.path.exists(ref_lib_path('tri_color_primal', np.float32, 16)):
            threads = os.cpu_count() or 1
            libs = {s: RefLib(s, np.float32, 16) for s in ('tri_color_primal', 'tri_color_grad')}
            # bounded sample: every 8th triangle, every 4th row of its window
            xs = np.linspace(0.0, 1.0, res + 1)
            cols = [[] for _ in range(12)]
            for k in range(0, K, 8):
                a0, a1, c0, c1 = scene['windows'][k]
                nx = np.repeat(np.arange(a0, a1, 4), c1 - c0)
                ny = np.tile(np.arange(c0, c1), len(range(a0, a1, 4)))
                vals = [np.full(nx.size, 0.1), xs[nx], xs[nx + 1], xs[ny], xs[ny + 1]] + \
                       [np.full(nx.size, v) for v in scene['verts'][k]] + [np.full(nx.size, scene['col'][k])]
                for i, v in enumerate(vals):
                    cols[i].append(v)
            cols = [np.concatenate(c) for c in cols]
            n_s = cols[0].size
            libs['tri_color_primal'].eval(cols[1:], threads=threads)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                libs['tri_color_primal'].eval(cols[1:], threads=threads)
                libs['tri_color_grad'].eval(cols, threads=threads)
                best = min(best, time.perf_counter() - t0)
            cpu = {'value': n_s / best, 'unit': 'integral evals/s', 'cores': threads, 'kind': 'reference',
                   'sample': f'every 8th triangle, every 4th row of its window ({n_s} pixel-triangle pairs), best of 3'}
    line = {
        'metric': 'integral_evals_per_s', 'value': pairs_total * args.steps / (total_ms * 1e-3),
        'unit': 'integral evals/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'multi_tri_4096', 'res_per_gpu': [res, res], 'triangles': K, 'num_samples': 16,
                   'pixel_triangle_pairs': pairs_total, 'programs': ['tri_color_primal', 'tri_color_grad'], 'cull': bool(args.cull), 'cuda_graph': bool(args.graph),
                   'l2': 'L2 flushed between timed steps (256 MB write, outside the event pairs)',
                   'parallelism': f'pixel rows sharded over {world} GPU(s); one NCCL all-reduce of the [{K},7] gradient'},
        'phase_ms': phase, 'phase_note': 'phases timed launch by launch before the timed region'
        + ('; the timed steps replay one CUDA graph' if graph is not None else ''),
        'roofline': {'bound': 'fp32_fma', 'kernel': 'tri_color_primal_main (grouped)', 'achieved': achieved,
                     'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                     'algorithmic_ops_per_instance': oc.unguarded},
        'cpu_baseline': cpu, 'gpu_launches': int(launches), 'clocks': clocks,
        'check': {'mean_image': coverage, 'grad_norm': gnorm},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
