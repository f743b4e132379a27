"""CPU oracle for the Teg evaluation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``teg_b200`` (the product) never does.

Two checkers live here:

* ``Oracle`` -- ctypes wrapper over ``libteg_oracle.so`` (oracle/teg_oracle.c), a plain-C
  interpreter restating the reference C backend (citations in the C file).  kind = "port".
* ``RefLib`` -- ctypes wrapper over ``oracle/_ref/<prog>_<prec>_N<samples>.so``: the reference's
  OWN generated C (``teg.ir.passes.to_c`` run from /root/reference by oracle/build_ref.py)
  compiled with the reference's own flags (``g++ -O3 -std=c++11 -fPIC``, c_eval.py:119).
  kind = "reference".  Built only where /root/reference exists; the .so files travel.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libteg_oracle.so')
REF_DIR = os.path.join(HERE, '_ref')

_lib = None


def build(force: bool = False) -> str:
    """gcc-compile the interpreter (plain C99, OpenMP for the batch loop, no FMA contraction)."""
    srcs = [os.path.join(HERE, 'teg_oracle.c'), os.path.join(HERE, 'teg_oracle_impl.h')]
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ['gcc', '-O2', '-std=gnu99', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off',
           srcs[0], '-o', LIB_PATH, '-lm']
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.tego_load.restype = ctypes.c_void_p
        lib.tego_load.argtypes = [ctypes.c_char_p]
        lib.tego_free.argtypes = [ctypes.c_void_p]
        lib.tego_last_error.restype = ctypes.c_char_p
        lib.tego_num_args.argtypes = [ctypes.c_void_p]
        lib.tego_out_size.argtypes = [ctypes.c_void_p]
        for name, ct in (('tego_eval_f32_q', ctypes.c_float), ('tego_eval_f64_q', ctypes.c_double)):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ct), ctypes.c_long, ctypes.c_int,
                           ctypes.POINTER(ct), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong,
                           ctypes.c_long]
        _lib = lib
    return _lib


def _as_soa(inputs: Sequence, n: Optional[int], dtype) -> np.ndarray:
    cols = [np.asarray(v, dtype=dtype) for v in inputs]
    if n is None:
        n = max([c.size for c in cols if c.ndim > 0] + [1])
    out = np.empty((len(cols), n), dtype=dtype)
    for i, c in enumerate(cols):
        out[i, :] = c           # scalars broadcast
    return out


class Oracle:
    """Interpreter-backed evaluation of one TEGP program over a batch of instances."""

    kind = 'port'

    def __init__(self, tegp_text: str):
        lib = _load()
        self._lib = lib
        self._h = lib.tego_load(tegp_text.encode())
        if not self._h:
            raise ValueError(f'oracle: {lib.tego_last_error().decode()}')
        self.n_args = lib.tego_num_args(self._h)
        self.out_size = lib.tego_out_size(self._h)
        if self.out_size < 0:
            raise ValueError(f'oracle: {lib.tego_last_error().decode()}')

    def __del__(self):
        if getattr(self, '_h', None):
            self._lib.tego_free(self._h)
            self._h = None

    def eval(self, inputs: Sequence, num_samples: int, dtype=np.float32, n: Optional[int] = None,
             threads: int = 1, quadrature: str = 'midpoint', seed: int = 0, inst_base: int = 0) -> np.ndarray:
        """inputs: one scalar or length-n array per program argument (in ``args`` order).
        quadrature: 'midpoint' (the C backend's rule), 'trapezoid' (the numpy backend's, numpy_eval.py:57-68) or
        'monte_carlo' (uniform samples from Philox4x32-10 keyed by ``seed``; ``inst_base`` = index of the first
        instance, teg_oracle_impl.h case 'T').
        Returns an array [out_size, n] in ``dtype``."""
        dtype = np.dtype(dtype)
        assert len(inputs) == self.n_args, f'expected {self.n_args} inputs, got {len(inputs)}'
        soa = _as_soa(inputs, n, dtype)
        n = soa.shape[1]
        out = np.empty((self.out_size, n), dtype=dtype)
        if dtype == np.float32:
            fn, ct = self._lib.tego_eval_f32_q, ctypes.c_float
        elif dtype == np.float64:
            fn, ct = self._lib.tego_eval_f64_q, ctypes.c_double
        else:
            raise ValueError('dtype must be float32 or float64')
        q = {'midpoint': 0, 'trapezoid': 1, 'monte_carlo': 2}[quadrature]
        rc = fn(self._h, soa.ctypes.data_as(ctypes.POINTER(ct)), n, int(num_samples),
                out.ctypes.data_as(ctypes.POINTER(ct)), self.out_size, int(threads), q,
                ctypes.c_ulonglong(int(seed) & 0xFFFFFFFFFFFFFFFF), ctypes.c_long(int(inst_base)))
        if rc != 0:
            raise RuntimeError(f'oracle: {self._lib.tego_last_error().decode()}')
        return out


def ref_lib_path(name: str, dtype, num_samples: int) -> str:
    prec = 'f32' if np.dtype(dtype) == np.float32 else 'f64'
    return os.path.join(REF_DIR, f'{name}_{prec}_N{num_samples}.so')


class RefLib:
    """The reference's own generated C for one program, compiled by oracle/build_ref.py."""

    kind = 'reference'

    def __init__(self, name: str, dtype, num_samples: int):
        self.path = ref_lib_path(name, dtype, num_samples)
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.dtype = np.dtype(dtype)
        self.num_samples = num_samples
        self._lib = ctypes.CDLL(self.path)
        self._lib.teg_ref_num_args.restype = ctypes.c_int
        self._lib.teg_ref_out_size.restype = ctypes.c_int
        self.n_args = self._lib.teg_ref_num_args()
        self.out_size = self._lib.teg_ref_out_size()
        ct = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
        self._ct = ct
        self._lib.teg_ref_run.argtypes = [ctypes.POINTER(ct), ctypes.POINTER(ct), ctypes.c_long, ctypes.c_int]
        self._lib.teg_ref_run.restype = None

    def eval(self, inputs: Sequence, num_samples: Optional[int] = None, dtype=None, n: Optional[int] = None,
             threads: int = 1) -> np.ndarray:
        assert num_samples in (None, self.num_samples)
        assert dtype is None or np.dtype(dtype) == self.dtype
        assert len(inputs) == self.n_args, f'expected {self.n_args} inputs, got {len(inputs)}'
        soa = _as_soa(inputs, n, self.dtype)
        n = soa.shape[1]
        out = np.empty((self.out_size, n), dtype=self.dtype)
        self._lib.teg_ref_run(soa.ctypes.data_as(ctypes.POINTER(self._ct)),
                              out.ctypes.data_as(ctypes.POINTER(self._ct)), n, int(threads))
        return out
