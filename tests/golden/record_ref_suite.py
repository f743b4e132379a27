"""Record every C-backend ``evaluate`` call of the reference's own test-suite as a fixture.

Run where /root/reference exists (takes ~7 min: the reference JIT-compiles and spawns a process per call):

    PYTHONHASHSEED=0 python tests/golden/record_ref_suite.py

Loads /root/reference/tests/test.py unmodified, with ``teg.eval.evaluate`` wrapped so that each call made
with the default / 'C' backend is stored as (program in TEGP form, bindings by label, num_samples, the
value the reference's C backend returned).  Output: tests/golden/ref_suite.json.gz.  The replay tests
(tests/test_ref_suite_replay.py) push the same programs and bindings through the oracle (CPU) and the
CUDA backend (GPU box) and compare with what the reference returned -- SURVEY.md §8(c): "re-running
tests/test.py with the backend defaulted to the new CUDA mode is the acceptance test for drop-in
behaviour", made portable to machines without the ``teg`` frontend.
"""
import gzip
import hashlib
import importlib.util
import inspect
import json
import os
import sys
import unittest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('TEG_REFERENCE_PATH', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'tests'))
sys.path.insert(0, ROOT)
sys.setrecursionlimit(200000)

import teg.eval  # noqa: E402
import teg.eval.backend as ref_backend  # noqa: E402
from teg_b200.tegp import from_teg, label_of  # noqa: E402

real_evaluate = ref_backend.evaluate
programs = {}
calls = []


def _caller():
    for fr in inspect.stack()[2:]:
        slf = fr.frame.f_locals.get('self')
        if isinstance(slf, unittest.TestCase):
            return slf.id().split('.', 1)[-1]
    return '?'


def recording_evaluate(expr, bindings=None, backend=None, **kwargs):
    result = real_evaluate(expr, bindings, backend=backend, **kwargs)
    if backend in (None, 'C'):
        try:
            prog = from_teg(expr, name='ref_suite')
            text = prog.dumps()
            key = hashlib.sha256(text.encode()).hexdigest()[:16]
            programs[key] = text
            b = {label_of(k): float(v) for k, v in (bindings or {}).items()}
            res = [float(x) for x in result] if isinstance(result, (list, tuple)) else float(result)
            calls.append({'test': _caller(), 'program': key, 'bindings': b,
                          'num_samples': int(kwargs.get('num_samples', 50)), 'result': res})
        except Exception as e:   # a call the exchange format cannot express is reported, not hidden
            calls.append({'test': _caller(), 'error': f'{type(e).__name__}: {e}'})
    return result


teg.eval.evaluate = recording_evaluate
ref_backend.evaluate = recording_evaluate

spec = importlib.util.spec_from_file_location('ref_test', os.path.join(REF, 'tests', 'test.py'))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
mod.evaluate = recording_evaluate

suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
res = unittest.TextTestRunner(verbosity=1).run(suite)
out = {'reference_tests_run': res.testsRun, 'failures': len(res.failures), 'errors': len(res.errors),
       'programs': programs, 'calls': calls}
path = os.path.join(HERE, 'ref_suite.json.gz')
with gzip.open(path, 'wt') as f:
    json.dump(out, f)
print(f'{len(calls)} calls, {len(programs)} programs, {sum(1 for c in calls if "error" in c)} unrecordable -> {path}')
