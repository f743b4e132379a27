"""ctypes binding of libtegb200.so (include/tegb200.h).

There is no fallback: if the library is missing or a device is unavailable the calls raise.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
from typing import Sequence

from . import build as _build

_lib = None


class TegbError(RuntimeError):
    pass


class ProgramDesc(ctypes.Structure):
    _fields_ = [('uniform_kernel', ctypes.c_char_p), ('main_kernel', ctypes.c_char_p),
                ('uniform_symbol', ctypes.c_char_p), ('elem_size', ctypes.c_int), ('n_uniform', ctypes.c_int),
                ('n_scalar_args', ctypes.c_int), ('n_array_args', ctypes.c_int), ('n_grid_args', ctypes.c_int),
                ('n_outputs', ctypes.c_int), ('block_size', ctypes.c_int), ('reduce_outputs', ctypes.c_int), ('team', ctypes.c_int), ('vec', ctypes.c_int), ('grouped', ctypes.c_int),
                ('n_pixel_args', ctypes.c_int), ('accumulate', ctypes.c_int), ('n_tile_conds', ctypes.c_int),
                ('tile_rows', ctypes.c_int), ('tile_cols', ctypes.c_int), ('tile_samples', ctypes.c_int), ('tile_kernel', ctypes.c_char_p)]


class Grid(ctypes.Structure):
    _fields_ = [('x_lo', ctypes.c_double), ('x_hi', ctypes.c_double), ('y_lo', ctypes.c_double),
                ('y_hi', ctypes.c_double), ('res_x', ctypes.c_int32), ('res_y', ctypes.c_int32),
                ('row_begin', ctypes.c_int32), ('row_end', ctypes.c_int32)]


SYMBOLS = [
    'tegb_last_error', 'tegb_version', 'tegb_runtime_header_hash', 'tegb_nvrtc_version', 'tegb_device_count', 'tegb_compile', 'tegb_module_from_cubin',
    'tegb_module_cubin', 'tegb_module_log', 'tegb_module_destroy', 'tegb_program_create',
    'tegb_program_destroy', 'tegb_program_launch', 'tegb_program_eval_host', 'tegb_render_launch',
    'tegb_render_launch_strided', 'tegb_render_groups_launch', 'tegb_render_groups_prepare',
    'tegb_render_groups_launch_prepared', 'tegb_render_groups_gather', 'tegb_program_tile_table', 'tegb_peer_create', 'tegb_peer_connect',
    'tegb_peer_allreduce_sum', 'tegb_peer_error', 'tegb_peer_destroy', 'tegb_program_set_peer',
    'tegb_program_set_peer_groups',
    'tegb_launch_count', 'tegb_program_enable_timing', 'tegb_program_main_kernel_ms', 'tegb_reduce_workspace_bytes', 'tegb_reduce_sum', 'tegb_nccl_unique_id',
    'tegb_nccl_comm_create', 'tegb_nccl_comm_destroy', 'tegb_allreduce_sum', 'tegb_fma_peak',
]


def lib():
    """Load (building first if the sources are newer) libtegb200.so.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.needs_build():
        try:
            _build.build_runtime()
        except Exception as e:   # no nvcc and no prebuilt library: loud failure, never a fallback
            if not os.path.exists(path):
                raise TegbError(f'libtegb200.so is missing and could not be built: {e}') from e
    L = ctypes.CDLL(path)
    vp, cp, i32, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int64
    L.tegb_last_error.restype = cp
    L.tegb_version.restype = i32
    L.tegb_runtime_header_hash.restype = ctypes.c_uint64
    L.tegb_device_count.restype = i32
    L.tegb_compile.argtypes = [cp, cp, cp, ctypes.POINTER(vp)]
    L.tegb_module_from_cubin.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    L.tegb_module_cubin.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    L.tegb_module_log.argtypes = [vp]
    L.tegb_module_log.restype = cp
    L.tegb_module_destroy.argtypes = [vp]
    L.tegb_module_destroy.restype = None
    L.tegb_program_create.argtypes = [vp, ctypes.POINTER(ProgramDesc), i32, ctypes.POINTER(vp)]
    L.tegb_program_destroy.argtypes = [vp]
    L.tegb_program_destroy.restype = None
    L.tegb_program_launch.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(vp),
                                      ctypes.POINTER(vp), i64, vp]
    L.tegb_program_eval_host.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(vp),
                                         ctypes.POINTER(vp), i64]
    L.tegb_render_launch.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(Grid),
                                     ctypes.POINTER(vp), vp]
    L.tegb_render_groups_launch.argtypes = [vp, ctypes.POINTER(vp), i32, i32, ctypes.POINTER(Grid), vp, i32, i32,
                                            ctypes.POINTER(vp), ctypes.POINTER(vp), vp]
    L.tegb_render_groups_launch_prepared.argtypes = L.tegb_render_groups_launch.argtypes
    L.tegb_render_groups_prepare.argtypes = [vp, ctypes.POINTER(vp), i32, i32, ctypes.POINTER(Grid), vp, i32, i32, vp]
    L.tegb_render_groups_gather.argtypes = [vp, ctypes.POINTER(vp), i32, ctypes.POINTER(Grid), vp, i32, i32, vp, vp, i32,
                                            ctypes.POINTER(vp), ctypes.POINTER(vp), vp]
    L.tegb_program_tile_table.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.tegb_render_launch_strided.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(Grid), i32, i32,
                                             ctypes.POINTER(vp), vp]
    L.tegb_peer_create.argtypes = [i32, i32, i32, i32, ctypes.POINTER(vp), vp]
    L.tegb_peer_connect.argtypes = [vp, vp]
    L.tegb_peer_allreduce_sum.argtypes = [vp, vp, i32, i32, vp]
    L.tegb_peer_error.argtypes = [vp]
    L.tegb_peer_destroy.argtypes = [vp]
    L.tegb_peer_destroy.restype = None
    L.tegb_program_set_peer.argtypes = [vp, vp]
    L.tegb_program_set_peer_groups.argtypes = [vp, vp, i32, vp]
    L.tegb_program_enable_timing.argtypes = [vp, i32]
    L.tegb_program_main_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.tegb_launch_count.restype = i64
    L.tegb_reduce_workspace_bytes.argtypes = [i32, i64, i32]
    L.tegb_reduce_workspace_bytes.restype = ctypes.c_size_t
    L.tegb_reduce_sum.argtypes = [ctypes.POINTER(vp), i32, i64, i32, vp, vp, ctypes.c_size_t, vp]
    L.tegb_nccl_unique_id.argtypes = [vp]
    L.tegb_nccl_comm_create.argtypes = [vp, i32, i32, i32, ctypes.POINTER(vp)]
    L.tegb_nccl_comm_destroy.argtypes = [vp]
    L.tegb_allreduce_sum.argtypes = [vp, i64, i32, vp, vp]
    L.tegb_fma_peak.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise TegbError(lib().tegb_last_error().decode(errors='replace'))


def last_error() -> str:
    return lib().tegb_last_error().decode(errors='replace')


def device_count() -> int:
    return lib().tegb_device_count()


def launch_count() -> int:
    return int(lib().tegb_launch_count())


# ------------------------------------------------------------------------------- JIT + cubin cache
def cache_dir() -> str:
    d = os.environ.get('TEGB200_CACHE_DIR') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '_jit_cache')
    os.makedirs(d, exist_ok=True)
    return d


class Module:
    """A compiled translation unit.  Content-hash keyed on-disk cubin cache (the only persistent state)."""

    def __init__(self, source: str, arch: str = 'sm_100a', opts: str = '', use_cache: bool = True):
        L = lib()
        self._L = L
        self.handle = ctypes.c_void_p()
        # the unit includes the runtime header by name only: its text (embedded in the library) is part of the key
        key = hashlib.sha256(f'{L.tegb_version()}|{L.tegb_runtime_header_hash():016x}|{L.tegb_nvrtc_version()}|{arch}|{opts}|'
                             f'{source}'.encode()).hexdigest()[:32]
        path = os.path.join(cache_dir(), f'{key}.cubin')
        self.from_cache = False
        if use_cache and os.path.exists(path):
            with open(path, 'rb') as f:
                blob = f.read()
            buf = ctypes.create_string_buffer(blob, len(blob))
            check(L.tegb_module_from_cubin(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(self.handle)))
            self.from_cache = True
            self.cubin_path = path
            return
        check(L.tegb_compile(source.encode(), arch.encode(), opts.encode(), ctypes.byref(self.handle)))
        self.cubin_path = None
        if use_cache:
            blob = self.cubin()
            tmp = path + f'.tmp{os.getpid()}'
            with open(tmp, 'wb') as f:
                f.write(blob)
            os.replace(tmp, path)
            self.cubin_path = path

    @classmethod
    def from_cubin(cls, blob: bytes) -> "Module":
        """A module from a ready cubin (e.g. one compiled ahead of time by nvcc)."""
        self = cls.__new__(cls)
        self._L = lib()
        self.handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        check(self._L.tegb_module_from_cubin(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(self.handle)))
        self.from_cache, self.cubin_path = False, None
        return self

    def cubin(self) -> bytes:
        ptr = ctypes.c_void_p()
        size = ctypes.c_size_t()
        check(self._L.tegb_module_cubin(self.handle, ctypes.byref(ptr), ctypes.byref(size)))
        return ctypes.string_at(ptr.value, size.value)

    def log(self) -> str:
        return self._L.tegb_module_log(self.handle).decode(errors='replace')

    def __del__(self):
        try:
            if getattr(self, 'handle', None) and self.handle.value:
                self._L.tegb_module_destroy(self.handle)
                self.handle.value = None
        except Exception:       # interpreter shutdown: module globals may already be gone
            pass


def _ptr_array(ptrs: Sequence[int]):
    arr = (ctypes.c_void_p * max(1, len(ptrs)))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def _dbl_array(vals: Sequence[float]):
    arr = (ctypes.c_double * max(1, len(vals)))()
    for i, v in enumerate(vals):
        arr[i] = float(v)
    return arr


class Program:
    """A loaded program: uniform prologue kernel + batched main kernel (see codegen.KernelSource)."""

    def __init__(self, module: Module, ks, device: int = 0):
        L = lib()
        self._L = L
        self.module = module
        self.ks = ks
        self.device = device
        self.elem_size = 4 if ks.ctype == 'float' else 8
        self._names = (ks.uniform_kernel.encode(), ks.main_kernel.encode(), b'tegU',
                       (getattr(ks, 'tile_kernel', '') or '').encode() or None)
        desc = ProgramDesc(self._names[0], self._names[1], self._names[2], self.elem_size, ks.n_uniform,
                           len(ks.scalar_args), len(ks.array_args), len(getattr(ks, 'grid_args', ())),
                           ks.n_outputs, ks.block_size, int(getattr(ks, 'reduce_outputs', False)),
                           int(getattr(ks, 'team', 1)), int(getattr(ks, 'vec', 1)), int(getattr(ks, 'grouped', False)),
                           len(getattr(ks, 'pixel_args', None) or ()) + len(getattr(ks, 'pixel_diff', None) or ()),
                           int(getattr(ks, 'accumulate', False)),
                           int(getattr(ks, 'tile_conds', 0)), int(getattr(ks, 'tile_rows', 0)),
                           int(getattr(ks, 'tile_cols', 0) or 32), int(ks.num_samples), self._names[3])
        self.handle = ctypes.c_void_p()
        check(L.tegb_program_create(module.handle, ctypes.byref(desc), device, ctypes.byref(self.handle)))

    def tile_table(self, count: int, stream: int = 0):
        """Tri-states of the tile conditions written by the last render (diagnostics): ndarray [count]."""
        import numpy as np
        out = np.empty(count, dtype=np.float32 if self.elem_size == 4 else np.float64)
        check(self._L.tegb_program_tile_table(self.handle, ctypes.c_void_p(out.ctypes.data),
                                              ctypes.c_size_t(out.nbytes), ctypes.c_void_p(stream)))
        return out

    def launch(self, scalars: Sequence[float], in_ptrs: Sequence[int], out_ptrs: Sequence[int], n: int,
               stream: int = 0) -> None:
        check(self._L.tegb_program_launch(self.handle, _dbl_array(scalars), _ptr_array(in_ptrs),
                                          _ptr_array(out_ptrs), n, ctypes.c_void_p(stream)))

    def eval_host(self, scalars: Sequence[float], host_in: Sequence[int], host_out: Sequence[int], n: int) -> None:
        check(self._L.tegb_program_eval_host(self.handle, _dbl_array(scalars), _ptr_array(host_in),
                                             _ptr_array(host_out), n))

    def render(self, scalars: Sequence[float], grid: Grid, out_ptrs: Sequence[int], stream: int = 0,
               interleave=(1, 0)) -> None:
        check(self._L.tegb_render_launch_strided(self.handle, _dbl_array(scalars), ctypes.byref(grid),
                                                 int(interleave[0]), int(interleave[1]),
                                                 _ptr_array(out_ptrs), ctypes.c_void_p(stream)))

    def render_groups(self, group_ptrs: Sequence[int], group_begin: int, group_count: int, grid: Grid,
                      windows_ptr: int, max_rows: int, max_cols: int, pixel_ptrs: Sequence[int],
                      out_ptrs: Sequence[int], stream: int = 0, prepared: bool = False) -> None:
        fn = self._L.tegb_render_groups_launch_prepared if prepared else self._L.tegb_render_groups_launch
        check(fn(self.handle, _ptr_array(group_ptrs), group_begin, group_count, ctypes.byref(grid),
                 ctypes.c_void_p(windows_ptr), max_rows, max_cols, _ptr_array(pixel_ptrs), _ptr_array(out_ptrs),
                 ctypes.c_void_p(stream)))

    def render_groups_gather(self, group_ptrs: Sequence[int], group_count: int, grid: Grid, windows_ptr: int,
                             max_rows: int, max_cols: int, bin_begin_ptr: int, bin_groups_ptr: int,
                             pixel_ptrs: Sequence[int], out_ptrs: Sequence[int], stream: int = 0,
                             clear: bool = False) -> None:
        check(self._L.tegb_render_groups_gather(self.handle, _ptr_array(group_ptrs), group_count, ctypes.byref(grid),
                                                ctypes.c_void_p(windows_ptr), max_rows, max_cols,
                                                ctypes.c_void_p(bin_begin_ptr), ctypes.c_void_p(bin_groups_ptr),
                                                int(bool(clear)),
                                                _ptr_array(pixel_ptrs), _ptr_array(out_ptrs), ctypes.c_void_p(stream)))

    def prepare_groups(self, group_ptrs: Sequence[int], group_begin: int, group_count: int, grid: Grid,
                       windows_ptr: int, max_rows: int, max_cols: int, stream: int = 0) -> None:
        check(self._L.tegb_render_groups_prepare(self.handle, _ptr_array(group_ptrs), group_begin, group_count,
                                                 ctypes.byref(grid), ctypes.c_void_p(windows_ptr), max_rows, max_cols,
                                                 ctypes.c_void_p(stream)))

    def enable_timing(self, on: bool = True) -> None:
        """CUDA events around the main kernel of every launch of this handle (measurement)."""
        check(self._L.tegb_program_enable_timing(self.handle, int(bool(on))))

    def main_kernel_ms(self) -> float:
        """Duration of the main kernel of the last launch (waits for it)."""
        ms = ctypes.c_double()
        check(self._L.tegb_program_main_kernel_ms(self.handle, ctypes.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if getattr(self, 'handle', None) and self.handle.value:
                self._L.tegb_program_destroy(self.handle)
                self.handle.value = None
        except Exception:       # interpreter shutdown
            pass


def reduce_workspace_bytes(n_cols: int, n: int, elem_size: int) -> int:
    return int(lib().tegb_reduce_workspace_bytes(n_cols, n, elem_size))


def reduce_sum(in_ptrs: Sequence[int], n: int, elem_size: int, out_ptr: int, ws_ptr: int, ws_bytes: int,
               stream: int = 0) -> None:
    check(lib().tegb_reduce_sum(_ptr_array(in_ptrs), len(in_ptrs), n, elem_size, ctypes.c_void_p(out_ptr),
                                ctypes.c_void_p(ws_ptr), ws_bytes, ctypes.c_void_p(stream)))


def fma_peak(elem_size: int = 4, device: int = 0):
    t = ctypes.c_double()
    mhz = ctypes.c_double()
    check(lib().tegb_fma_peak(elem_size, device, ctypes.byref(t), ctypes.byref(mhz)))
    return t.value, mhz.value


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    check(lib().tegb_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return buf.raw


class NcclComm:
    def __init__(self, unique_id: bytes, world_size: int, rank: int, device: int):
        assert len(unique_id) == 128
        self._L = lib()
        self.handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(unique_id, 128)
        check(self._L.tegb_nccl_comm_create(ctypes.cast(buf, ctypes.c_void_p), world_size, rank, device,
                                            ctypes.byref(self.handle)))

    def allreduce_sum(self, ptr: int, count: int, elem_size: int, stream: int = 0) -> None:
        check(self._L.tegb_allreduce_sum(ctypes.c_void_p(ptr), count, elem_size, self.handle,
                                         ctypes.c_void_p(stream)))

    def destroy(self):
        if self.handle.value:
            self._L.tegb_nccl_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()
