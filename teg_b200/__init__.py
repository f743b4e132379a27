"""teg_b200 -- a B200-native (sm_100a) evaluation backend for Teg.

Takes a *reduced* Teg program (the base-language tree produced by Teg's own frontend: integrand,
FwdDeriv / reverse-derivative forms and the Dirac-delta boundary terms of ``reduce_to_base``) and
evaluates it by midpoint quadrature over a batch of independent instances as generated CUDA kernels,
behind Teg's ``evaluate()`` / ``EvalMode`` plugin interface.

    from teg_b200 import evaluate, CUDA_EvalMode, render_image
    evaluate(expr, {t: np.linspace(0, 1, 1 << 20)}, backend='CUDA', num_samples=50)

``teg_b200.raster`` renders scenes of many triangles (grouped launches), ``teg_b200.optim`` runs the optimisation loop
(loss, upstream gradient, Adam) on the device, ``teg_b200.parallel`` holds the multi-GPU exchange.

Importing the package registers the ``'CUDA'`` backend with ``teg.eval.backend.BACKENDS`` when the
reference ``teg`` package is importable.
"""
from .tegp import Program, from_teg                      # noqa: F401
from .eval import CUDA_EvalMode, EvalMode, BACKENDS, evaluate, render_image, register   # noqa: F401

register()

__all__ = ['Program', 'from_teg', 'CUDA_EvalMode', 'EvalMode', 'BACKENDS', 'evaluate', 'render_image', 'register']
