"""SURVEY 8(b) threading row: handles are independent; ONE handle may also be shared by threads and streams.

The reference's executor is "NOT thread-safe" (teg/eval/c_eval.py:55-56).  Here a handle owns its published uniform
table and scratch buffers: its launch calls are serialised by a mutex and ordered across streams by an event
(csrc/tegb200.cu begin_launch / end_launch); prepared calls own private handles and overlap freely."""
import threading

import numpy as np
import pytest

from conftest import load_program

from teg_b200.workloads import TRIANGLE

pytestmark = pytest.mark.gpu

RES = (512, 384)


def _scalars(prog, k):
    """k-th variation of the triangle (moves every vertex a little: different uniform tables)"""
    out = {}
    for lab in prog.args[4:]:
        out[lab] = TRIANGLE[lab.rsplit('_', 1)[0]] + 0.01 * k * (1 if 'x' in lab else -1)
    return out


@pytest.mark.parametrize('scene,reduce', [('tri_primal', False), ('tri_grad', True)])
def test_two_threads_two_streams_one_mode(cuda_device, scene, reduce):
    """Two Python threads, each on its own stream, hammer ONE CUDA_EvalMode (one shared handle) with different
    scalars; every result equals the serial one bit for bit."""
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program(scene)
    mode = CUDA_EvalMode(prog, num_samples=16)
    box = prog.args[:4]
    iters = 24
    serial = []
    for k in range(2 * iters):
        out = mode.render(box, RES, ((0, 1), (0, 1)), _scalars(prog, k), reduce=reduce)
        torch.cuda.synchronize()
        serial.append(out.cpu().numpy().copy())
    assert not np.array_equal(serial[0], serial[1])
    results = {}
    errors = []

    def worker(tid):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            outs = []
            for j in range(iters):
                k = 2 * j + tid
                outs.append((k, mode.render(box, RES, ((0, 1), (0, 1)), _scalars(prog, k), stream=st.cuda_stream,
                                            reduce=reduce)))
            st.synchronize()
            for k, o in outs:
                results[k] = o.cpu().numpy()
        except Exception as e:          # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(2 * iters):
        assert np.array_equal(results[k], serial[k]), f'launch {k} differs from the serial result'


def test_prepared_renders_own_their_handles(cuda_device):
    """Two PreparedRenders of the same program with different scalars on two streams: private handles (no shared
    uniform table), so they may overlap on the device and still give the serial bits."""
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program('tri_primal')
    mode = CUDA_EvalMode(prog, num_samples=16)
    box = prog.args[:4]
    a = mode.prepare_render(box, RES, ((0, 1), (0, 1)), _scalars(prog, 0))
    b = mode.prepare_render(box, RES, ((0, 1), (0, 1)), _scalars(prog, 5))
    assert a.prog.handle.value != b.prog.handle.value
    want_a = mode.render(box, RES, ((0, 1), (0, 1)), _scalars(prog, 0)).cpu().numpy()
    want_b = mode.render(box, RES, ((0, 1), (0, 1)), _scalars(prog, 5)).cpu().numpy()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):
        a.run(s1.cuda_stream)
        b.run(s2.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(a.out.cpu().numpy(), want_a) and np.array_equal(b.out.cpu().numpy(), want_b)


def test_array_api_from_two_threads(cuda_device):
    """eval() with device arrays from two threads on one mode (shared handle, torch's per-thread current stream)."""
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program('readme_deriv')
    mode = CUDA_EvalMode(prog, num_samples=50)
    t = torch.linspace(-0.25, 1.25, 1 << 16, device='cuda')
    want = {dt: mode.eval({prog.args[0]: t, prog.args[1]: dt}).cpu().numpy() for dt in (1.0, 2.0)}
    got, errors = {}, []

    def worker(dt):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()):
                outs = [mode.eval({prog.args[0]: t, prog.args[1]: dt}) for _ in range(50)]
                torch.cuda.current_stream().synchronize()
            got[dt] = [o.cpu().numpy() for o in outs]
        except Exception as e:          # pragma: no cover
            errors.append(e)

    th = [threading.Thread(target=worker, args=(dt,)) for dt in (1.0, 2.0)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors
    for dt in (1.0, 2.0):
        assert all(np.array_equal(o, want[dt]) for o in got[dt])
    assert not np.array_equal(want[1.0], want[2.0])
