"""The C ABI used from plain C (examples/c_abi_demo.c): CLI-generated unit + companion desc header, gcc, run."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

PROG = '''
from teg_b200.lang import TegVar, Var, Teg, IfElse
x, t = TegVar('x'), Var('t', 0.5)
__PROGRAM__ = Teg(0, 1, IfElse(x < t, 1, 0), x)
__ARGLIST__ = [t]
'''


def test_c_program_drives_the_library(cuda_device, tmp_path):
    (tmp_path / 'prog.py').write_text(PROG)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'teg_b200', '-c', '-t', 'CUDA', '-f', 'single', '-s', '50', '-m', 'readme',
                        '-o', str(tmp_path / 'readme.cu'), 'prog.py'], cwd=tmp_path, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / 'readme_desc.h').exists()
    exe = tmp_path / 'demo'
    lib_dir = os.path.join(ROOT, 'teg_b200')
    subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), '-I', str(tmp_path),
                    os.path.join(ROOT, 'examples', 'c_abi_demo.c'), '-o', str(exe), '-L', lib_dir, '-ltegb200',
                    f'-Wl,-rpath,{lib_dir}', '-lm'], check=True)
    r = subprocess.run([str(exe), str(tmp_path / 'readme.cu')], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'c_abi_demo: 1000 instances' in r.stdout
