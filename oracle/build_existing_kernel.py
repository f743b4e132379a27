"""The "existing kernel" bar: the reference's OWN ``--target CUDA_C`` output in the thread-per-instance
wrapper a Teg maintainer would write around it.  MEASUREMENT INFRASTRUCTURE; runs only where /root/reference exists.

The reference can already emit its scalar functions with a ``__device__`` decorator (``python -m teg --compile --target
CUDA_C``: teg/__main__.py:100-109, ``device_code=True`` in teg/ir/passes/to_c.py:402,440,557,598) but ships no kernel,
no launcher and no batching.  Here that text -- produced by the reference's own ``to_c`` exactly as its CLI drives it --
is wrapped in

    extern "C" __global__ <scene>_main(xe, ye, res_y, row0, n_rows, tile_rows, row_stride, row_phase, out_k...)

(one thread per pixel of the tests/plot.py grid, pixel boxes from the two edge tables, one output image per component of
the result; the launch-wide arguments arrive through the ``tegU`` table), i.e. the launch interface of this repository's
un-tiled render programs, so that bench.py can time it through the same C ABI (``tegb_render_launch``) on the same box.
It is compiled with nvcc's defaults (``-O3``, FMA contraction on: what `nvcc file.cu` gives a maintainer) for sm_100a
into ``oracle/_ref/existing_<scene>_f32_N<samples>.cubin`` (git-ignored, shipped to the GPU box like the _ref libraries).
Nothing from /root/reference is copied into the repository: the generated text lives in a temp dir.

Usage: python oracle/build_existing_kernel.py [--force]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ROOT = os.environ.get('TEG_REFERENCE_PATH', '/root/reference')
REF_DIR = os.path.join(HERE, '_ref')

# (scene, num_samples): the headline workload of bench.py
KERNELS = [('tri_primal', 16), ('tri_grad', 16)]
BLOCK = 128


def cubin_path(scene: str, ns: int) -> str:
    return os.path.join(REF_DIR, f'existing_{scene}_f32_N{ns}.cubin')


def meta_path(scene: str, ns: int) -> str:
    return os.path.join(REF_DIR, f'existing_{scene}_f32_N{ns}.json')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'teg'))


def _wrapper(scene: str, entry: str, c_arg_labels, prog_args, out_size: int) -> str:
    """c_arg_labels: the C entry's parameters in order; prog_args: TEGP argument order (x_lb, x_ub, y_lb, y_ub first)."""
    box = {prog_args[0]: 'xe[nx]', prog_args[1]: 'xe[nx + 1]', prog_args[2]: 'ye_lo', prog_args[3]: 'ye_hi'}
    scal = [a for a in prog_args if a not in box]
    slot = {a: j for j, a in enumerate(scal)}
    call = ', '.join(box[a] if a in box else f'tegU[{slot[a]}]' for a in c_arg_labels)
    outs = ', '.join(f'float* __restrict__ out{k}' for k in range(out_size))
    if out_size == 1:
        store = f'out0[i] = {entry}({call});'
    else:
        store = (f'generic_array<float,{out_size}> g_ = {entry}({call}); ' +
                 ' '.join(f'out{k}[i] = g_[{k}];' for k in range(out_size)))
    uparams = ', '.join([f'const float a{j}' for j in range(len(scal))] + ['float* __restrict__ U'])
    ubody = ' '.join(f'U[{j}] = a{j};' for j in range(len(scal)))
    return f'''
__constant__ float tegU[{max(1, len(scal))}];
extern "C" __global__ void {scene}_uniform({uparams}) {{ if (blockIdx.x == 0 && threadIdx.x == 0) {{ {ubody} }} }}
extern "C" __global__ void {scene}_main(const float* __restrict__ xe, const float* __restrict__ ye, const int res_y,
                                        const int row0, const int n_rows, const int tile_rows, const int row_stride,
                                        const int row_phase, {outs}) {{
    const int ny = blockIdx.x * blockDim.x + threadIdx.x;
    if (ny >= res_y) return;
    const float ye_lo = ye[ny], ye_hi = ye[ny + 1];
    for (int r = 0; r < tile_rows; ++r) {{
        const int rel = (blockIdx.y * row_stride + row_phase) * tile_rows + r;
        if (rel >= n_rows) break;
        const int nx = row0 + rel;
        const long long i = (long long)(blockIdx.y * tile_rows + r) * res_y + ny;
        {store}
    }}
}}
'''


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    args = ap.parse_args(argv)
    if not reference_available():
        print(f'build_existing_kernel: {REF_ROOT} not present, nothing to do')
        return 0
    if os.environ.get('PYTHONHASHSEED') != '0':
        env = dict(os.environ, PYTHONHASHSEED='0', PYTHONDONTWRITEBYTECODE='1')
        return subprocess.call([sys.executable, os.path.abspath(__file__), *(argv or sys.argv[1:])], env=env)
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, ROOT)
    sys.setrecursionlimit(200000)
    from teg.ir.passes.to_c import to_c
    from teg.lang.base import Var
    from teg.passes.compile import to_ir
    from oracle.scenes import SCENES
    from teg_b200.tegp import from_teg

    os.makedirs(REF_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='teg_existing_kernel_')
    for scene, ns in KERNELS:
        out, meta = cubin_path(scene, ns), meta_path(scene, ns)
        if not args.force and os.path.exists(out) and os.path.exists(meta):
            continue
        Var.global_uid = 1000
        program, arglist = SCENES[scene]()
        prog = from_teg(program, arglist, name=scene)
        # driven like teg/__main__.py:81-109 (--target CUDA_C -f single -s <ns> -m <scene>_)
        ir_func = to_ir(program)
        ir_func.func_name = f'{scene}_'
        labelled = {s.label: s for s in ir_func.inputs}
        ordered = [labelled[f'{v.name}_{v.uid}'] for v in arglist if f'{v.name}_{v.uid}' in labelled]
        ir_func.inputs = ordered + [s for s in ir_func.inputs if s not in ordered]
        funclist, c_args = to_c(ir_func, device_code=True, float_width='float', num_samples=ns, fn_name_prefix=f'{scene}_')
        out_size = ir_func.output.irtype().size
        code = '#include "teg_cuda_runtime.h"\n' + '\n'.join(body for _, body in funclist) + '\n'
        code += _wrapper(scene, f'{scene}_', [a[0] for a in c_args], prog.args, out_size)
        src = os.path.join(tmp, f'existing_{scene}.cu')
        with open(src, 'w') as f:
            f.write(code)
        cmd = ['nvcc', '-cubin', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++14',
               '-Xptxas', '-v', f'-I{os.path.join(REF_ROOT, "teg", "include")}', src, '-o', out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(f'[existing {scene}] nvcc FAILED:\n{r.stderr[-3000:]}')
            return 1
        lines = r.stderr.split('\n')
        at = [j for j, ln in enumerate(lines) if f"'{scene}_main'" in ln]
        usage = [ln for ln in lines[at[0]:at[0] + 4] if 'registers' in ln] if at else []
        n_scal = len(prog.args) - 4
        with open(meta, 'w') as f:
            json.dump({'scene': scene, 'num_samples': ns, 'n_outputs': out_size, 'n_scalar_args': n_scal,
                       'scalar_args': prog.args[4:], 'block_size': BLOCK, 'lines_of_reference_cuda_c': code.count('\n'),
                       'flags': 'nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (FMA contraction on)',
                       'ptxas': usage[-1].strip() if usage else ''}, f, indent=1)
        print(f'[existing {scene}] {code.count(chr(10))} lines of reference CUDA_C -> {out}: {usage[-1].strip() if usage else ""}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
