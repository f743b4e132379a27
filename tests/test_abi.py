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
