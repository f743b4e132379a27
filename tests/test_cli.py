"""`python -m teg_b200 --compile --target CUDA ...` -- the reference CLI flags (teg/__main__.py:11-29)."""
import os
import subprocess
import sys

from conftest import ROOT

PROG = '''
from teg_b200.lang import TegVar, Var, Teg, IfElse
x, t = TegVar('x'), Var('t', 0.5)
__PROGRAM__ = Teg(0, 1, IfElse(x < t, 1, 0), x)
__ARGLIST__ = [t]
'''


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, '-m', 'teg_b200', *args], cwd=cwd, env=env, capture_output=True, text=True)


def test_compile_target_cuda_writes_kernels_and_cubin(tmp_path):
    (tmp_path / 'prog.py').write_text(PROG)
    out = tmp_path / 'readme.cu'
    r = _run(['-c', '-t', 'CUDA', '-f', 'double', '-s', '8', '-m', 'readme', '-o', str(out), '--cubin', 'prog.py'],
             tmp_path)
    assert r.returncode == 0, r.stderr
    text = out.read_text()
    assert 'typedef double T;' in text and 'readme_main' in text and 'readme_uniform' in text
    assert '< 8u' in text                                  # 8 midpoint samples
    assert (tmp_path / 'readme.cubin').stat().st_size > 1000


def test_includes_flags(tmp_path):
    r = _run(['-I'], tmp_path)
    assert r.returncode == 0 and os.path.exists(os.path.join(r.stdout.strip(), 'teg_b200_runtime.cuh'))
    r = _run(['-i'], tmp_path)
    assert r.stdout.strip().startswith('-I')


def test_exactly_one_mode_and_valid_precision(tmp_path):
    (tmp_path / 'prog.py').write_text(PROG)
    assert _run([], tmp_path).returncode != 0
    assert _run(['-c', '-f', 'half', 'prog.py'], tmp_path).returncode != 0
    assert _run(['-c', '-t', 'OPENCL', 'prog.py'], tmp_path).returncode != 0


def test_compact_c_target_is_bit_identical_to_the_oracle(tmp_path):
    """--target C from the backend's IR: same entry contract as to_c.py, ~60x less code,
    same bits."""
    import ctypes
    import numpy as np
    from conftest import load_program, program_text
    from oracle import Oracle
    from teg_b200.cli import emit_c
    from teg_b200.workloads import bind, scene_values
    prog = load_program('tri_grad')
    code = emit_c(prog, float_type='float', samples=16, method='tri_grad')
    assert code.count('\n') < 1200                       # the reference emits 33.5 k lines for this program
    n_args = len(prog.args)
    params = ', '.join(f'in[{a} * n + i]' for a in range(n_args))
    code += ('\nextern "C" void run(const float* in, float* out, long n) { for (long i = 0; i < n; i++) { '
             f'generic_array<float,6> r = tri_grad({params}); for (int j = 0; j < 6; j++) out[j * n + i] = r[j]; }} }}\n')
    src = tmp_path / 't.cc'
    src.write_text('template<typename T, int N> struct generic_array { T values[N]; T& operator[](int i) '
                   '{ return values[i]; } };\n' + code)
    so = tmp_path / 't.so'
    subprocess.run(['g++', '-O2', '-std=c++11', '-fPIC', '-shared', '-ffp-contract=off', str(src), '-o', str(so)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    inputs = bind(prog.args, scene_values('tri_grad', res=(24, 24)))
    n = 576
    soa = np.empty((n_args, n), dtype=np.float32)
    for i, v in enumerate(inputs):
        soa[i] = v
    out = np.empty((6, n), dtype=np.float32)
    P = ctypes.POINTER(ctypes.c_float)
    lib.run(soa.ctypes.data_as(P), out.ctypes.data_as(P), ctypes.c_long(n))
    assert np.array_equal(out, Oracle(program_text('tri_grad')).eval(inputs, 16, np.float32))


def test_integration_mode_flag_selects_the_trapezoid_rule(tmp_path):
    """--integration-mode takes the names of to_c.py:389-393; unknown names are refused."""
    (tmp_path / 'prog.py').write_text(PROG)
    out = tmp_path / 'trap.cu'
    r = _run(['-c', '-t', 'CUDA', '-s', '9', '-m', 'trap', '-o', str(out), '--integration-mode',
              'trapezoidal_quadrature', 'prog.py'], tmp_path)
    assert r.returncode == 0, r.stderr
    text = out.read_text()
    assert '/ (T)8;' in text and '== 8u) ? hi' in text        # step = (hi - lo) / (N - 1); last point = hi
    assert _run(['-c', '--integration-mode', 'simpson', 'prog.py'], tmp_path).returncode != 0


def test_monte_carlo_flag(tmp_path):
    (tmp_path / 'prog.py').write_text(PROG)
    out = tmp_path / 'mc.cu'
    r = _run(['-c', '-t', 'CUDA', '-s', '32', '-m', 'mc', '-o', str(out), '--integration-mode', 'monte_carlo_uniform',
              '--seed', '42', 'prog.py'], tmp_path)
    assert r.returncode == 0, r.stderr
    text = out.read_text()
    assert '#define TEGB_MC_SEED 42ull' in text and 'tegb_uniform(TEGB_MC_SEED, tegb_inst, 0u' in text


def test_bench_keeps_library_chatter_off_stdout():
    """bench.py prints ONE JSON line: text written to stdout from C while NCCL initialises is sent to stderr."""
    code = ('import sys, ctypes; sys.path.insert(0, %r); import bench\n'
            'libc = ctypes.CDLL(None)\n'
            'with bench._StdoutToStderr():\n'
            '    libc.puts(b"NCCL version (from C)")\n'
            'print(\'{"json": 1}\')\n') % ROOT
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == '{"json": 1}\n' and 'NCCL version (from C)' in r.stderr


def test_bench_reference_arm_prints_one_json_line_without_a_gpu():
    """`bench.py --impl reference`: the reference's own C (oracle/_ref) or the oracle port on the host cores, same
    metric / unit / workload keys as the CUDA arm, plus cpu_baseline and e2e objects; rank != 0 prints nothing."""
    import json
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1', '--cpu-rows', '4'], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'integral_evals_per_s' and d['unit'] == 'integral evals/s'
    assert d['config']['workload'] == 'tri_raster_4096' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                         '--warmup', '1', '--cpu-rows', '4'], capture_output=True, text=True, env=dict(env, RANK='1'))
    assert r1.returncode == 0 and r1.stdout.strip() == ''


def test_bench_reference_arm_of_the_multi_triangle_workload():
    """`--workload multi_tri_4096 --impl reference`: config 5 on the host cores, the same config dict as the CUDA arm."""
    import json
    import bench
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0', '--workload', 'multi_tri_4096', '--res', '1024'], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d['impl'] == 'reference' and d['value'] > 0 and d['scaling'] == 'weak'
    assert d['config'] == bench.headline_config('multi_tri_4096', 1024, 'weak')
    assert d['config']['programs'] == ['tri_color_primal', 'tri_color_grad']
    # the two arms of the headline workload describe it with the same dict too (the driver compares them)
    assert bench.headline_config('tri_raster_4096', 4096, 'weak')['workload'] == 'tri_raster_4096'
