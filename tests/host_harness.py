"""TEST-ONLY: compile the generated kernel text for the HOST and run it in a loop.

The CUDA emitter writes the per-instance body as ``TEGB_FN`` functions; here ``TEGB_FN`` is re-defined
to ``static inline`` so g++ can compile the very same statements.  This lets the CPU test-suite check the
backend's *lowering* (DAG, scheduling, guards, loop fusion, uniform hoisting) against the oracle on
machines without a GPU.  It is a checker living under tests/: the product has no host evaluation path
(teg_b200_runtime.cuh #errors without nvcc/NVRTC).
"""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_DIR = os.path.join(tempfile.gettempdir(), 'tegb200_host_harness')
_INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'teg_b200', 'include')


def compile_for_host(ks):
    T = ks.ctype
    pre = ('#include <math.h>\n#define TEGB_FN static inline\n'
           '#define TEGB_UNIFORM_TABLE(T, name, n) static T name[n]\n#define __restrict__\n')
    uargs = ', '.join([f'sc[{j}]' for j in range(len(ks.scalar_args))] + ['tegU'])
    varying = list(ks.array_args) + list(ks.grid_args)
    n_in = len(varying) + int(getattr(ks, 'tile_conds', 0))        # tile tri-states come last (pseudo arguments)
    margs = ', '.join([f'in[{j}][i]' for j in range(n_in)] +
                      (['(unsigned long long)(inst_base + i)'] if getattr(ks, 'takes_instance_index', False) else []) + ['res'])
    stores = ' '.join(f'out[{k}][i] = res[{k}];' for k in range(ks.n_outputs))
    # tile-culled units: pick the body the kernel would pick for the tile states of this instance
    n_var = len(varying)
    call = ''
    for label, match in (getattr(ks, 'tile_bodies', None) or []):
        cond = ' && '.join('(' + ' || '.join(f'in[{n_var + k}][i] == {v}' for v in match[k]) + ')' for k in sorted(match))
        call += f'if ({cond}) {ks.name}_tile_{label}({margs}); else '
    call += f'{ks.name}_instance({margs});'
    drv = (f'\nstatic long long inst_base = 0;\nextern "C" void set_inst_base(long long b) {{ inst_base = b; }}\n'
           f'extern "C" void run(const {T}* sc, const {T}* const* in, {T}* const* out, long long n) {{\n'
           f'  {ks.name}_uniform_body({uargs});\n'
           f'  for (long long i = 0; i < n; i++) {{ {T} res[{ks.n_outputs}]; {call} {stores} }}\n}}\n')
    text = pre + ks.source + drv
    os.makedirs(_DIR, exist_ok=True)
    key = hashlib.sha256(text.encode()).hexdigest()[:20]
    src, so = os.path.join(_DIR, f'{key}.cc'), os.path.join(_DIR, f'{key}.so')
    if not os.path.exists(so):
        with open(src, 'w') as f:
            f.write(text)
        subprocess.run(['g++', '-O1', '-std=c++11', '-fPIC', '-shared', '-ffp-contract=off', f'-I{_INCLUDE}',
                        src, '-o', so + '.tmp'], check=True)
        os.replace(so + '.tmp', so)
    return ctypes.CDLL(so)


def run_on_host(ks, inputs, n, dtype, tile_inputs=(), inst_base=0):
    """tile_inputs: one per-instance array of tri-states (0 / 1 / 2) per tile condition of a tiled render unit."""
    lib = compile_for_host(ks)
    assert len(tile_inputs) == int(getattr(ks, 'tile_conds', 0))
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    P = ctypes.POINTER(ct)
    sc = np.array([inputs[a] for a in ks.scalar_args] + [0], dtype=dtype)
    varying = sorted(list(ks.array_args) + list(ks.grid_args))
    arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(inputs[a], dtype=dtype), (n,))) for a in varying]
    arrs += [np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=dtype), (n,))) for t in tile_inputs]
    outs = [np.empty(n, dtype=dtype) for _ in range(ks.n_outputs)]
    inp = (P * max(1, len(arrs)))(*[a.ctypes.data_as(P) for a in arrs])
    outp = (P * len(outs))(*[o.ctypes.data_as(P) for o in outs])
    lib.set_inst_base(ctypes.c_longlong(inst_base))
    lib.run(sc.ctypes.data_as(P), inp, outp, ctypes.c_longlong(n))
    return np.stack(outs)
