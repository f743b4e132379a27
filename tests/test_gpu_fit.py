"""Optimisation loop on the device (teg_b200/optim.py): value + reverse derivative with an
upstream per-pixel dL drive Adam on the six vertex coordinates.  Loss, moments and parameters stay on the device: the
loop makes no host synchronisation (checked with torch's sync debug mode); the loss must fall and the vertices converge."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_triangle_fit_converges_without_host_syncs(cuda_device):
    """200 iterations at 1024^2 as replays of ONE captured CUDA graph; torch raises on any synchronising call made
    inside the loop (``.item()``, ``.cpu()``, blocking copies)."""
    import torch
    from teg_b200.optim import fit_triangle
    # set-up, capture and the first (eager) iteration may synchronise; the loop may not
    losses, v, vt, fit = fit_triangle(res=1024, iters=0, lr=4e-3, graph=True)
    assert fit.graph is not None
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode('error')
    try:
        fit.run(200)
    finally:
        torch.cuda.set_sync_debug_mode('default')
    torch.cuda.synchronize()
    losses = fit.loss_history()
    verts, _ = fit.params()
    assert len(losses) == 200 and np.all(np.isfinite(losses))
    assert losses[-1] < 0.02 * losses[0], (losses[0], losses[-1])
    assert np.abs(verts[0] - vt).max() < 0.02


def test_eager_loop_matches_the_captured_one(cuda_device):
    """launch-by-launch iterations (no graph) give the same parameters as graph replays, bit for bit, and make no
    host synchronisation either"""
    import torch
    from teg_b200.optim import fit_triangle
    _, _, _, a = fit_triangle(res=128, iters=0, graph=True)
    _, _, _, b = fit_triangle(res=128, iters=0, graph=False)
    b.iteration()                    # allocations of the eager path happen once
    b.p.copy_(a.p); b.m.zero_(); b.v.zero_(); b.t.zero_(); b.it.zero_(); b.losses.zero_(); b._push_params()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode('error')
    try:
        a.run(40)
        b.run(40)
    finally:
        torch.cuda.set_sync_debug_mode('default')
    torch.cuda.synchronize()
    assert torch.equal(a.p, b.p) and torch.equal(a.losses[:40], b.losses[:40])
    la = a.loss_history()
    assert la[-1] < 0.2 * la[0]


def test_legacy_tool_entry_point(cuda_device):
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from fit_triangle import fit
    losses, v, vt = fit(res=128, iters=120, lr=5e-3)
    assert losses[-1] < 0.02 * losses[0], (losses[0], losses[-1])
    assert np.abs(v - vt).max() < 0.02
