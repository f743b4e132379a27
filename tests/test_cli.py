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
