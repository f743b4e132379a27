"""Build the reference-side artefacts.  TEST INFRASTRUCTURE; runs only where /root/reference exists.

For every scene in oracle/scenes.py:

  1. build the program with the reference frontend (deterministic: PYTHONHASHSEED=0, Var uid reset);
  2. write its TEGP text to tests/golden/programs/<scene>.tegp  (committed fixture; what the CUDA
     backend and the interpreter oracle consume on machines without ``teg``);
  3. emit the reference's OWN C for it -- ``to_ir`` (teg/passes/compile.py:51) ->
     ``to_c(float_width, num_samples)`` (teg/ir/passes/to_c.py:635) -- wrap the entry function in an
     ``extern "C"`` SoA batch loop and compile it with the reference's JIT flags
     (``g++ -O3 -std=c++11 -fPIC``, teg/eval/c_eval.py:119; plus -shared -fopenmp for the batch loop)
     into oracle/_ref/<scene>_<f32|f64>_N<samples>.so.

Nothing from /root/reference is copied into the repository: the generated C lives in a temp dir and
only the compiled .so (git-ignored, but shipped to the GPU box) is kept.

Usage: python oracle/build_ref.py [--only scene[,scene]] [--jobs J] [--force]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ROOT = os.environ.get('TEG_REFERENCE_PATH', '/root/reference')
REF_DIR = os.path.join(HERE, '_ref')
PROG_DIR = os.path.join(ROOT, 'tests', 'golden', 'programs')

# (precision, num_samples) variants compiled per scene
VARIANTS = {
    'readme_primal': [('f32', 50), ('f64', 50)],
    'readme_deriv': [('f32', 50), ('f64', 50)],
    'tri_primal': [('f32', 16), ('f64', 16)],
    'tri_grad': [('f32', 16), ('f64', 16)],
    'raster_theta_primal': [('f32', 10)],
    'raster_theta': [('f32', 10), ('f64', 10)],
    'shaded_primal': [('f32', 16), ('f64', 16)],
    'shaded_grad': [('f32', 16), ('f64', 16)],
    'tri_color_primal': [('f32', 16), ('f64', 16)],
    'tri_color_grad': [('f32', 16), ('f64', 16)],
    # config 4: the top of the sample-count sweep as well -- N = 1024 / 4096 in 2-D, 128 / 256 in 3-D
    'nested2d_value': [('f32', 16), ('f32', 64), ('f64', 16), ('f32', 1024), ('f32', 4096), ('f64', 1024)],
    'nested2d_grad': [('f32', 16), ('f32', 64), ('f64', 16), ('f32', 1024), ('f32', 4096), ('f64', 1024)],
    'nested3d_value': [('f32', 16), ('f64', 16), ('f32', 128), ('f32', 256)],
    'nested3d_grad': [('f32', 16), ('f64', 16), ('f32', 128), ('f32', 256)],
    'simple_line': [('f32', 20)],
    'simple_line_grad': [('f32', 20), ('f64', 20)],
    'circle': [('f32', 20)],
    'circle_dt': [('f32', 20), ('f64', 20)],
    'rect_hyperbola': [('f32', 20)],
    'rect_hyperbola_grad': [('f32', 20), ('f64', 20)],
    'bilinear_lerp': [('f32', 20)],
    'bilinear_lerp_grad': [('f32', 20), ('f64', 20)],
}


# variants whose golden vectors hold fewer instances (N^2 / N^3 samples each): tests/golden/vectors_high_n/
HIGH_N_INSTANCES = 64


def is_high_n(scene: str, ns: int) -> bool:
    return scene.startswith('nested') and ns >= 128


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'teg'))


def _wrapper(entry: str, call_slots, n_args: int, out_size: int, ctype: str) -> str:
    """call_slots[j] = which input slot (TEGP ``args`` order) feeds the j-th C parameter."""
    call = ', '.join(f'in[{a}L * n + i]' for a in call_slots)
    if out_size == 1:
        store = f'{ctype} r = {entry}({call}); out[i] = r;'
    else:
        store = (f'generic_array<{ctype},{out_size}> r = {entry}({call}); '
                 f'for (int j = 0; j < {out_size}; j++) out[(long)j * n + i] = r[j];')
    return f'''
extern "C" int teg_ref_num_args() {{ return {n_args}; }}
extern "C" int teg_ref_out_size() {{ return {out_size}; }}
extern "C" void teg_ref_run(const {ctype}* in, {ctype}* out, long n, int threads) {{
    if (threads < 1) threads = 1;
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (long i = 0; i < n; i++) {{ {store} }}
}}
'''


def emit_reference_c(program, arglist, ctype: str, num_samples: int, prefix: str) -> str:
    """The reference's own codegen, driven exactly like teg/__main__.py:81-97."""
    from teg.passes.compile import to_ir
    from teg.ir.passes.to_c import to_c

    ir_func = to_ir(program)
    ir_func.func_name = prefix
    labelled = {s.label: s for s in ir_func.inputs}
    ordered = [labelled[f'{v.name}_{v.uid}'] for v in arglist if f'{v.name}_{v.uid}' in labelled]
    rest = [s for s in ir_func.inputs if s not in ordered]
    ir_func.inputs = ordered + rest
    funclist, c_args = to_c(ir_func, float_width=ctype, num_samples=num_samples, fn_name_prefix=prefix)
    out_size = ir_func.output.irtype().size
    code = '#include <math.h>\n#include "teg_c_runtime.h"\n'
    for _, body in funclist:
        code += body + '\n'
    return code, [a[0] for a in c_args], out_size


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='')
    ap.add_argument('--jobs', type=int, default=max(1, (os.cpu_count() or 2)))
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--programs-only', action='store_true', help='write TEGP fixtures, skip the g++ builds')
    args = ap.parse_args(argv)

    if not reference_available():
        print(f'build_ref: {REF_ROOT} not present, nothing to do')
        return 0
    if os.environ.get('PYTHONHASHSEED') != '0':
        env = dict(os.environ, PYTHONHASHSEED='0', PYTHONDONTWRITEBYTECODE='1')
        return subprocess.call([sys.executable, os.path.abspath(__file__), *(argv or sys.argv[1:])], env=env)

    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, ROOT)
    sys.setrecursionlimit(200000)
    import teg
    from teg.lang.base import Var
    from oracle.scenes import SCENES
    from teg_b200.tegp import from_teg

    os.makedirs(REF_DIR, exist_ok=True)
    os.makedirs(PROG_DIR, exist_ok=True)
    manifest_path = os.path.join(REF_DIR, 'MANIFEST.json')
    manifest = {}
    if os.path.exists(manifest_path):
        with open(manifest_path) as f:
            manifest = json.load(f)

    only = [s for s in args.only.split(',') if s]
    tmp = tempfile.mkdtemp(prefix='teg_ref_build_')
    jobs = []
    for scene, builder in SCENES.items():
        if only and scene not in only:
            continue
        path = os.path.join(PROG_DIR, f'{scene}.tegp')
        if not args.force and not args.programs_only and os.path.exists(path):
            # fast path: every variant was already built from exactly this committed TEGP text
            cur = hashlib.sha256(open(path).read().encode()).hexdigest()
            keys = [f'{scene}_{prec}_N{ns}' for prec, ns in VARIANTS.get(scene, [])]
            if all(manifest.get(k, {}).get('tegp_sha256') == cur and
                   os.path.exists(os.path.join(REF_DIR, k + '.so')) for k in keys):
                continue
        t0 = time.time()
        Var.global_uid = 1000          # reproducible labels, clear of the module-level constants
        program, arglist = builder()
        prog = from_teg(program, arglist, name=scene)
        text = prog.dumps()
        sha = hashlib.sha256(text.encode()).hexdigest()
        old = open(path).read() if os.path.exists(path) else None
        if old != text:
            with open(path, 'w') as f:
                f.write(text)
        print(f'[{scene}] frontend {time.time() - t0:.1f}s  tegp nodes={len(prog.nodes)} '
              f'tree={prog.tree_size()} args={len(prog.args)} {"(changed)" if old != text else ""}', flush=True)
        if args.programs_only:
            continue
        for prec, ns in VARIANTS.get(scene, []):
            so = os.path.join(REF_DIR, f'{scene}_{prec}_N{ns}.so')
            key = f'{scene}_{prec}_N{ns}'
            if not args.force and os.path.exists(so) and manifest.get(key, {}).get('tegp_sha256') == sha:
                continue
            ctype = 'float' if prec == 'f32' else 'double'
            t1 = time.time()
            code, c_args, out_size = emit_reference_c(program, arglist, ctype, ns, prefix=f'{scene}_')
            # the C entry takes its own (used-only) argument list; the batch takes TEGP ``args`` order,
            # in which arguments the program never reads still occupy a slot
            slot = {lab: i for i, lab in enumerate(prog.args)}
            wrapper = _wrapper(f'{scene}_', [slot[lab] for lab in c_args], len(prog.args), out_size, ctype)
            src = os.path.join(tmp, f'{key}.cc')
            with open(src, 'w') as f:
                f.write(code + wrapper)
            print(f'[{key}] reference to_c {time.time() - t1:.1f}s, {code.count(chr(10))} lines', flush=True)
            jobs.append((key, src, so, sha))

    def compile_one(job):
        key, src, so, sha = job
        t1 = time.time()
        cmd = ['g++', '-O3', '-std=c++11', '-fPIC', '-shared', '-fopenmp',
               f'-I{os.path.join(REF_ROOT, "teg", "include")}', src, '-o', so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(f'[{key}] g++ FAILED:\n{r.stderr[-2000:]}', flush=True)
            return key, None
        print(f'[{key}] g++ {time.time() - t1:.1f}s', flush=True)
        return key, sha

    with ThreadPoolExecutor(max_workers=args.jobs) as ex:
        for key, sha in ex.map(compile_one, jobs):
            if sha is not None:
                manifest[key] = {'tegp_sha256': sha, 'flags': 'g++ -O3 -std=c++11 -fPIC -shared -fopenmp'}
    with open(manifest_path, 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
