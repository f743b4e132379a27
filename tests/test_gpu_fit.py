"""Optimisation loop through the CUDA backend: value + reverse derivative with an
upstream per-pixel dL drive Adam on the six vertex coordinates; the loss must fall and the vertices converge."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_triangle_fit_converges(cuda_device):
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from fit_triangle import fit
    losses, v, vt = fit(res=128, iters=120, lr=5e-3)
    assert losses[-1] < 0.02 * losses[0], (losses[0], losses[-1])
    assert np.abs(v - vt).max() < 0.02
