#!/usr/bin/env python
"""Throughput of every BASELINE.json config on one GPU, next to the reference C on the host cores.

    python tools/sweep.py [--quick] > profiles/rNN_sweep.jsonl

One JSON line per (config, program, precision, num_samples): instances/s on the GPU (device buffers,
CUDA events, best of 5 after warm-up), the kernel variant used (thread per instance or team), the max
error against the oracle on a sample of the batch, and -- where oracle/_ref has the same variant compiled --
the reference's own C on all host cores.  These are the parity-test configurations of SURVEY.md §8(d)
(configs 1-4); the headline number is bench.py's.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import Oracle, RefLib, ref_lib_path  # noqa: E402
from teg_b200 import CUDA_EvalMode  # noqa: E402
from teg_b200.opcount import count as op_count  # noqa: E402
from teg_b200.tegp import Program  # noqa: E402
from teg_b200.workloads import bind, scene_values  # noqa: E402

PROG = os.path.join(ROOT, 'tests', 'golden', 'programs')


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3


CULL = True


def run(config, scene, ns, dtype, res=None, n=None, team=1, check=2048):
    text = open(os.path.join(PROG, f'{scene}.tegp')).read()
    prog = Program.loads(text)
    vals = scene_values(scene, res=res or (256, 256), n=n)
    inputs = bind(prog.args, vals)
    count = max(np.size(v) for v in inputs)
    fw = 'float' if dtype == np.float32 else 'double'
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    mode = CUDA_EvalMode(prog, num_samples=ns, float_width=fw, team=team, cull=CULL)
    b = {lab: (torch.as_tensor(np.asarray(v, dtype=dtype), device='cuda') if np.ndim(v) else v)
         for lab, v in zip(prog.args, inputs)}
    arr = [i for i, v in enumerate(inputs) if np.ndim(v) > 0]
    used_team = mode.pick_team(count, arr)
    out = mode.eval(b)
    call = mode.prepare(b)
    stream = torch.cuda.current_stream().cuda_stream
    t = gpu_time(lambda: call.run(stream))
    t_eval = gpu_time(lambda: mode.eval(b))
    # parity on a sample
    idx = np.linspace(0, count - 1, min(check, count)).astype(np.int64)
    sub = [np.asarray(v)[idx] if np.ndim(v) else v for v in inputs]
    want = Oracle(text).eval(sub, ns, dtype, threads=os.cpu_count() or 1)
    got = out.detach().cpu().numpy().reshape(-1, count)[:, idx]
    scale = np.maximum(np.abs(want), np.max(np.abs(want)) * 1e-3 + 1e-30)
    err = float(np.nanmax(np.abs(got - want) / scale))
    sch = mode.reference_schedule(arr)          # the op count is defined on the reference's program
    oc = op_count(sch, ns)
    line = {'config': config, 'program': scene, 'dtype': 'f32' if dtype == np.float32 else 'f64', 'num_samples': ns,
            'instances': int(count), 'team': used_team, 'cull': bool(getattr(mode.sched_dag, 'culled', [])), 'gpu_s': t, 'gpu_s_via_eval': t_eval, 'gpu_instances_per_s': count / t,
            'max_scaled_err_vs_oracle': err, 'bit_exact': bool(np.array_equal(got, want, equal_nan=True)),
            'alg_ops_per_instance_unguarded': oc.unguarded, 'alg_ops_per_instance_all_guards': oc.all_true,
            'alg_tflops_lower_bound': oc.unguarded * count / t / 1e12,
            'io_bytes_per_instance': (len(arr) + len(mode.dag.outputs)) * (4 if dtype == np.float32 else 8)}
    line['hbm_GBps'] = line['io_bytes_per_instance'] * count / t / 1e9
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            line['hbm_frac_of_measured_peak'] = line['hbm_GBps'] / json.load(f)['hbm_gbs']
    except Exception:
        line['hbm_frac_of_measured_peak'] = line['hbm_GBps'] / 6650.0        # fallback stated in the profiling guide
    path = ref_lib_path(scene, dtype, ns)
    if os.path.exists(path):
        ref = RefLib(scene, dtype, ns)
        threads = os.cpu_count() or 1
        m = min(count, 1 << 18)
        sub = [np.asarray(v)[:m] if np.ndim(v) else v for v in inputs]
        ref.eval(sub, threads=threads)
        t0 = time.perf_counter()
        ref.eval(sub, threads=threads)
        dt = time.perf_counter() - t0
        line['cpu_reference_instances_per_s'] = m / dt
        line['cpu_cores'] = threads
        line['speedup'] = line['gpu_instances_per_s'] / line['cpu_reference_instances_per_s']
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--cull', type=int, default=1, help='0: dense kernels; 1: domain culling (default)')
    args = ap.parse_args()
    global CULL
    CULL = bool(args.cull)
    f32, f64 = np.float32, np.float64
    # config 1: README example, 1 Mi t-values
    for scene in ('readme_deriv', 'readme_primal'):
        for dt in (f32, f64):
            run('cfg1_readme_1Mi', scene, 50, dt, n=1 << 20)
    # config 1 at a size where the HBM streams, not the launch latency, set the time (HBM roofline)
    for dt in (f32, f64):
        run('cfg1_readme_64Mi_hbm', 'readme_deriv', 50, dt, n=1 << 26, check=4096)
    # config 2: triangle 256 x 256
    for scene in ('tri_primal', 'tri_grad', 'raster_theta'):
        ns = 10 if scene == 'raster_theta' else 16
        for dt in (f32, f64):
            run('cfg2_triangle_256', scene, ns, dt, res=(256, 256))
    # config 3: shaded triangle + render_test scenes, 1024 x 1024
    r3 = (512, 512) if args.quick else (1024, 1024)
    for scene, ns in (('shaded_primal', 16), ('shaded_grad', 16), ('simple_line', 20), ('simple_line_grad', 20),
                      ('circle', 20), ('circle_dt', 20), ('rect_hyperbola', 20), ('rect_hyperbola_grad', 20),
                      ('bilinear_lerp', 20), ('bilinear_lerp_grad', 20)):
        run('cfg3_render_1024', scene, ns, f32, res=r3)
    run('cfg3_render_1024', 'shaded_grad', 16, f64, res=r3)
    # config 4: nested 2-D / 3-D, 4096 instances, sample-count sweep; teams as chosen by the 'auto' policy
    sweep2 = (16, 64, 256) if args.quick else (16, 64, 256, 1024, 4096)
    for ns in sweep2:
        run('cfg4_nested2d_sweep', 'nested2d_value', ns, f32, n=4096, team='auto', check=64 if ns > 256 else 1024)
        run('cfg4_nested2d_sweep', 'nested2d_grad', ns, f32, n=4096, team='auto', check=256)
    sweep3 = (16, 32, 64) if args.quick else (16, 32, 64, 128, 256)
    for ns in sweep3:
        run('cfg4_nested3d_sweep', 'nested3d_value', ns, f32, n=4096, team='auto', check=16 if ns > 64 else 256)
        run('cfg4_nested3d_sweep', 'nested3d_grad', ns, f32, n=4096, team='auto', check=64 if ns > 64 else 256)
    run('cfg4_nested2d_sweep', 'nested2d_value', 256, f64, n=4096, team='auto', check=256)


if __name__ == '__main__':
    main()
