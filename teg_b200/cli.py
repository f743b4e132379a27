"""``python -m teg_b200`` -- the reference CLI surface (teg/__main__.py:11-113) with a ``CUDA`` target.

    python -m teg_b200 --compile --target CUDA -f single -s 16 -m NAME -o out.cu [--cubin] [--arrays a,b] prog.py

``prog.py`` exposes ``__PROGRAM__`` (a reduced expression built with ``teg`` or ``teg_b200.lang``) and an
optional ``__ARGLIST__`` fixing the argument order (teg/__main__.py:63-88); a ``.tegp`` file is accepted
too.  ``--includes`` / ``--include-options`` print the runtime include directory like the reference.
The ``CUDA`` target writes a complete translation unit: ``<NAME>_uniform`` + ``<NAME>_main`` batched
kernels (thread per instance, SoA).  ``--arrays`` names the arguments that are per-instance arrays
(default: all of them); the rest are launch-wide scalars evaluated once by the uniform prologue.
``C`` writes compact C++ from the backend's own IR (``emit_c``; set TEGB200_C_TARGET=reference to forward to the
reference CLI instead); ``CUDA_C`` is forwarded to the reference CLI when ``teg`` is importable.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from optparse import OptionParser

INCLUDE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'include')


def emit_cuda(program, arglist=(), float_type='float', samples=50, method='teg_method', arrays=None):
    """Lower a reduced program to the CUDA translation unit.  Returns codegen.KernelSource."""
    from .codegen import generate
    from .dag import build_dag
    from .schedule import make_schedule
    from .tegp import Program, from_teg
    from .workloads import base_name
    prog = program if isinstance(program, Program) else from_teg(program, arglist, name=method)
    dag = build_dag(prog)
    if arrays is None:
        array_args = list(range(len(prog.args)))
    else:
        names = set(arrays)
        array_args = [i for i, lab in enumerate(prog.args) if lab in names or base_name(lab) in names]
    sch = make_schedule(dag, array_args)
    return generate(sch, float_type, int(samples), name=method)


def emit_c(program, arglist=(), float_type='float', samples=50, method='teg_method') -> str:
    """Compact C++ for the reference's ``C`` target, from the backend's own IR.

    Same entry-point contract as teg/ir/passes/to_c.py:635-651 -- a function ``<method>(T arg0, T arg1, ...)``
    taking the arguments in ``__ARGLIST__`` order and returning ``T`` (or ``generic_array<T,k>`` for Tup
    programs, include teg_c_runtime.h) -- but after hash-consing, let-elimination, guard merging and loop
    fusion: a triangle reverse derivative is ~550 lines instead of 33.5 k, so g++ (and the reference's own C
    backend) compiles it in a fraction of the time.  Arithmetic is unchanged (one operation per reference
    operation, same order), so the values are bit-identical to the reference's C.  This is a *code generator
    target*; nothing in teg_b200 evaluates through it."""
    from .codegen import generate
    from .dag import build_dag
    from .schedule import make_schedule
    from .tegp import Program, from_teg
    prog = program if isinstance(program, Program) else from_teg(program, arglist, name=method)
    dag = build_dag(prog)
    n_args = len(prog.args)
    sch = make_schedule(dag, list(range(n_args)), uniform_hoist=False)      # every argument is a plain scalar
    ks = generate(sch, float_type, int(samples), name=method, use_asm=False, coop=False)
    k = ks.n_outputs
    T = float_type
    body = ks.source[:ks.source.index('#ifdef __CUDACC__')]
    body = body.replace('#include "teg_b200_runtime.cuh"\n', '')
    params = ', '.join(f'{T} a{a}' for a in range(n_args))
    call = ', '.join([f'a{a}' for a in range(n_args)] + ['res'])
    if k == 1:
        ret, tail = T, 'return res[0];'
    else:
        ret = f'generic_array<{T},{k}>'
        tail = f'{ret} out; for (int i = 0; i < {k}; i++) out[i] = res[i]; return out;'
    head = ('#include <math.h>\n#ifndef TEGB_FN\n#define TEGB_FN static inline\n#endif\n'
            '#ifndef TEGB_UNIFORM_TABLE\n#define TEGB_UNIFORM_TABLE(T, name, n) static T name[n]\n#endif\n'
            '#ifndef __CUDACC__\n#define __restrict__\n#endif\n')
    entry = (f'{ret} {method}({params}) {{\n    {T} res[{k}];\n    {method}_instance({call});\n    {tail}\n}}\n')
    return head + body + entry


def main(argv=None) -> int:
    parser = OptionParser()
    parser.add_option('-I', '--includes', dest='includes', action='store_true', default=False,
                      help='Returns C/C++ include directory for the CUDA runtime header')
    parser.add_option('-i', '--include-options', dest='include_options', action='store_true', default=False)
    parser.add_option('-c', '--compile', dest='compile', action='store_true', default=False)
    parser.add_option('-o', '--output', dest='output', default=None)
    parser.add_option('-t', '--target', dest='target', default='CUDA')
    parser.add_option('-f', '--fprecision', dest='precision', default='single')
    parser.add_option('-m', '--method', dest='method', default='teg_method')
    parser.add_option('-s', '--samples', dest='samples', default=50)
    parser.add_option('--arrays', dest='arrays', default=None,
                      help='comma-separated argument names that are per-instance arrays (default: all)')
    parser.add_option('--cubin', dest='cubin', action='store_true', default=False,
                      help='also JIT-compile for sm_100a and write <output>.cubin')
    options, args = parser.parse_args(argv)

    assert [options.includes, options.include_options, options.compile].count(True) == 1, \
        'Exactly one of --includes, --include-options and --compile can be specified'
    if options.include_options:
        print(f'-I"{INCLUDE_PATH}"' if ' ' in INCLUDE_PATH else f'-I{INCLUDE_PATH}')
        return 0
    if options.includes:
        print(INCLUDE_PATH)
        return 0

    assert options.precision in ['single', 'double'], \
        f'Invalid precision setting {options.precision}. Specify either single or double'
    float_type = {'single': 'float', 'double': 'double'}[options.precision]
    assert options.target in ['C', 'CUDA_C', 'CUDA'], 'No other backends are supported at the moment'
    input_file = args[0]
    if options.target == 'C' and os.environ.get('TEGB200_C_TARGET', 'compact') == 'compact':
        pass        # handled below: compact C from the backend's IR
    elif options.target in ('C', 'CUDA_C'):
        import runpy
        sys.argv = ['teg', *(argv if argv is not None else sys.argv[1:])]
        runpy.run_module('teg', run_name='__main__')
        return 0

    print(f'Building program {input_file}')
    if input_file.endswith('.tegp'):
        from .tegp import Program
        program, arglist = Program.load(input_file), []
    else:
        spec = importlib.util.spec_from_file_location('module.name', input_file)
        code_module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(code_module)
        assert '__PROGRAM__' in dir(code_module), f'Unable to find "__PROGRAM__" variable in {input_file}'
        program = code_module.__PROGRAM__
        arglist = getattr(code_module, '__ARGLIST__', [])
        if type(program).__module__.startswith('teg.'):
            from teg.passes.simplify import simplify          # teg/__main__.py:66-67
            print(f'Simplifying program {input_file}')
            program = simplify(program)
    if options.target == 'C':
        code = emit_c(program, arglist, float_type, options.samples, options.method)
        out = options.output or (options.method + '.h')
        with open(out, 'w') as f:
            f.write(code)
        print(f'Generated C code in {out} ({code.count(chr(10))} lines)')
        return 0
    arrays = None if options.arrays is None else [a for a in options.arrays.split(',') if a]
    ks = emit_cuda(program, arglist, float_type, options.samples, options.method, arrays)
    out = options.output or (options.method + '.cu')
    print(f'Converting to CUDA file {out}')
    with open(out, 'w') as f:
        f.write(ks.source)
    if options.cubin:
        from . import runtime
        blob = runtime.Module(ks.source, use_cache=False).cubin()
        with open(os.path.splitext(out)[0] + '.cubin', 'wb') as f:
            f.write(blob)
    print(f'Generated CUDA code in {out}: kernels {ks.uniform_kernel}, {ks.main_kernel}; '
          f'{len(ks.array_args)} array argument(s), {len(ks.scalar_args)} scalar, {ks.n_outputs} output(s)')
    return 0
