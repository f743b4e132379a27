"""``python -m teg_b200`` -- the reference CLI surface (teg/__main__.py:11-113) with a ``CUDA`` target.

    python -m teg_b200 --compile --target CUDA -f single -s 16 -m NAME -o out.cu [--cubin] [--arrays a,b] prog.py

``prog.py`` exposes ``__PROGRAM__`` (a reduced expression built with ``teg`` or ``teg_b200.lang``) and an
optional ``__ARGLIST__`` fixing the argument order (teg/__main__.py:63-88); a ``.tegp`` file is accepted
too.  ``--includes`` / ``--include-options`` print the runtime include directory like the reference.
The ``CUDA`` target writes a complete translation unit: ``<NAME>_uniform`` + ``<NAME>_main`` batched
kernels (thread per instance, SoA).  ``--arrays`` names the arguments that are per-instance arrays
(default: all of them); the rest are launch-wide scalars evaluated once by the uniform prologue.
``C`` and ``CUDA_C`` are forwarded to the reference CLI when ``teg`` is importable.
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
    if options.target in ('C', 'CUDA_C'):
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
