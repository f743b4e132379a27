"""The C-ABI shared library: loads without a GPU, exports every symbol include/tegb200.h declares,
and reports errors through status codes + tegb_last_error (never abort)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, 'include', 'tegb200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tegb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from teg_b200 import runtime
    lib = runtime.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'libtegb200.so lacks {n}'
    assert set(runtime.SYMBOLS) == set(names)
    assert lib.tegb_version() >= 100


def test_invalid_arguments_return_status_and_message():
    from teg_b200 import runtime
    lib = runtime.lib()
    out = ctypes.c_void_p()
    assert lib.tegb_compile(None, b'sm_100a', None, ctypes.byref(out)) == 1
    assert b'null' in lib.tegb_last_error()
    assert lib.tegb_module_from_cubin(None, 0, ctypes.byref(out)) == 1
    assert lib.tegb_reduce_sum(None, 0, 0, 4, None, None, 0, None) == 1
    assert lib.tegb_allreduce_sum(None, 0, 4, None, None) == 1
    assert lib.tegb_device_count() >= 0


def test_program_create_without_device_fails_loudly(tmp_path, monkeypatch):
    """On a machine without a CUDA device the product path must raise, not fall back."""
    import pytest
    from teg_b200 import runtime
    if runtime.device_count() > 0:
        pytest.skip('a CUDA device is present')
    monkeypatch.setenv('TEGB200_CACHE_DIR', str(tmp_path))
    from conftest import load_program
    from teg_b200 import CUDA_EvalMode
    prog = load_program('readme_primal')
    with pytest.raises(runtime.TegbError):
        CUDA_EvalMode(prog, num_samples=8).eval({prog.args[0]: 0.3})


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirrors of tegb_program_desc / tegb_grid (runtime.py) against the compiler's layout of
    include/tegb200.h: same size, same field offsets, same field order."""
    import subprocess
    from teg_b200 import runtime
    fields = [n for n, _ in runtime.ProgramDesc._fields_]
    gfields = [n for n, _ in runtime.Grid._fields_]
    prints = ''.join(f'    printf("%zu\\n", offsetof(tegb_program_desc, {n}));\n' for n in fields)
    gprints = ''.join(f'    printf("%zu\\n", offsetof(tegb_grid, {n}));\n' for n in gfields)
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tegb200.h"\nint main(void) {\n'
                   '    printf("%zu\\n", sizeof(tegb_program_desc));\n' + prints +
                   '    printf("%zu\\n", sizeof(tegb_grid));\n' + gprints + '    return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', f'-I{os.path.join(ROOT, "include")}', str(src), '-o', str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(runtime.ProgramDesc)
    assert vals[1:1 + len(fields)] == [getattr(runtime.ProgramDesc, n).offset for n in fields]
    rest = vals[1 + len(fields):]
    assert rest[0] == ctypes.sizeof(runtime.Grid)
    assert rest[1:] == [getattr(runtime.Grid, n).offset for n in gfields]
    # INTEGRATION.md shows the same field list
    with open(os.path.join(ROOT, 'INTEGRATION.md')) as f:
        doc = f.read()
    stub = doc[doc.index('class Desc(ctypes.Structure)'):doc.index('mod, prog = ctypes.c_void_p()')]
    assert re.findall(r"\('([a-z_]+)', ctypes", stub) == fields
