"""Synthetic instance batches for the BASELINE.json configs.

An *instance* is one evaluation of a program: a pixel box, a parameter set, ...  Every generator
returns ``{argument base name: scalar | float64 array[n]}``; ``bind(prog_args, values)`` orders them
like the program's TEGP ``args`` line.  Arrays are produced in float64 and cast by the consumer to the
working precision, the way the reference receives Python floats (c_eval.py:135-142).

Pixel boxes follow the reference renderer bit for bit: ``np.linspace(lo, hi, res + 1)`` in float64,
instance index ``nx * res_y + ny`` (tests/plot.py:11-16).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple, Union

import numpy as np

ArrayOrScalar = Union[float, np.ndarray]

TRIANGLE = dict(px0=0.15, py0=0.2, px1=0.85, py1=0.3, px2=0.4, py2=0.9)       # CCW, area 0.2325
SHADER = dict(c0=0.2, c1=0.5, c2=0.3, bg=0.05)


def base_name(label: str) -> str:
    """'x_lb_1002' -> 'x_lb' (labels are f'{name}_{uid}', compile.py:63)."""
    return label.rsplit('_', 1)[0]


def pixel_boxes(res: Tuple[int, int], bounds=((0.0, 1.0), (0.0, 1.0)),
                rows: Tuple[int, int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """x_lb, x_ub, y_lb, y_ub for pixels nx in rows[0]:rows[1] (all rows if None), ny in 0:res[1]."""
    xs = np.linspace(bounds[0][0], bounds[0][1], res[0] + 1)
    ys = np.linspace(bounds[1][0], bounds[1][1], res[1] + 1)
    r0, r1 = (0, res[0]) if rows is None else rows
    nx = np.repeat(np.arange(r0, r1), res[1])
    ny = np.tile(np.arange(res[1]), r1 - r0)
    return xs[nx], xs[nx + 1], ys[ny], ys[ny + 1]


def _boxes(names: Sequence[str], res, bounds, rows=None) -> Dict[str, np.ndarray]:
    return dict(zip(names, pixel_boxes(res, bounds, rows)))


def scene_values(scene: str, res: Tuple[int, int] = (256, 256), n: int = None, rows=None) -> Dict[str, ArrayOrScalar]:
    """Inputs for one of the oracle/scenes.py programs at the BASELINE sizes (or ``res`` / ``n``)."""
    if scene in ('readme_primal', 'readme_deriv'):
        n = (1 << 20) if n is None else n
        t = np.random.default_rng(0).uniform(-0.25, 1.25, size=n).astype(np.float32).astype(np.float64)
        return {'t': t, 'dt': 1.0}
    if scene in ('tri_primal', 'tri_grad'):
        return {**_boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((0, 1), (0, 1)), rows), **TRIANGLE}
    if scene in ('shaded_primal', 'shaded_grad'):
        return {**_boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((0, 1), (0, 1)), rows), **TRIANGLE, **SHADER,
                'dL': 1.0}
    if scene in ('tri_color_primal', 'tri_color_grad'):
        boxes = _boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((0, 1), (0, 1)), rows)
        m = boxes['x_lb'].size
        dl = np.random.default_rng(3).uniform(-1, 1, size=m).astype(np.float32).astype(np.float64)
        return {**boxes, **TRIANGLE, 'col': 0.7, 'dL': dl}
    if scene in ('raster_theta', 'raster_theta_primal'):
        return {**_boxes(('xl', 'xu', 'yl', 'yu'), res, ((0, 1), (0, 1)), rows), 'theta': 0.0, 'dtheta': 1.0}
    if scene.startswith('nested2d') or scene.startswith('nested3d'):
        n = 4096 if n is None else n
        d = 2 if scene.startswith('nested2d') else 3
        rng = np.random.default_rng(1)
        names = 'abcefg' if d == 2 else 'abcdefg'
        vals: Dict[str, ArrayOrScalar] = {nm: rng.uniform(-1, 1, size=n) for nm in names}
        for k in range(d):
            vals[f'l{k}'] = 0.0
            vals[f'h{k}'] = 1.0
        return vals
    if scene in ('simple_line', 'simple_line_grad'):
        return {**_boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((0, 1), (0, 1)), rows), 't1': 0.0, 't2': 0.0}
    if scene in ('circle', 'circle_dt'):
        return {**_boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((-1, 1), (-1, 1)), rows), 't': 0.5}
    if scene in ('rect_hyperbola', 'rect_hyperbola_grad'):
        return {**_boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((-1, 1), (-1, 1)), rows),
                't': 0.25, 't1': 1.0, 't2': 1.0, 't3': 0.0, 't4': 0.0}
    if scene in ('bilinear_lerp', 'bilinear_lerp_grad'):
        return {**_boxes(('x_lb', 'x_ub', 'y_lb', 'y_ub'), res, ((10e-3, 1 - 10e-3), (10e-3, 1 - 10e-3)), rows),
                'c00': 0.1, 'c11': 0.1, 'c01': 0.5, 'c10': 0.9, 't': 0.55}
    raise KeyError(scene)


def bind(prog_args: Sequence[str], values: Dict[str, ArrayOrScalar]) -> List[ArrayOrScalar]:
    """Order ``values`` (keyed by base name) like the program's argument labels."""
    out = []
    for lab in prog_args:
        nm = base_name(lab)
        if nm not in values:
            raise KeyError(f'no value for program argument {lab}')
        out.append(values[nm])
    return out


# default number of samples per scene = what oracle/build_ref.py compiled the reference C with
SCENE_SAMPLES = {
    'readme_primal': 50, 'readme_deriv': 50, 'tri_primal': 16, 'tri_grad': 16,
    'raster_theta_primal': 10, 'raster_theta': 10, 'shaded_primal': 16, 'shaded_grad': 16, 'tri_color_primal': 16, 'tri_color_grad': 16,
    'nested2d_value': 16, 'nested2d_grad': 16, 'nested3d_value': 16, 'nested3d_grad': 16,
    'simple_line': 20, 'simple_line_grad': 20, 'circle': 20, 'circle_dt': 20,
    'rect_hyperbola': 20, 'rect_hyperbola_grad': 20, 'bilinear_lerp': 20, 'bilinear_lerp_grad': 20,
}


# ----------------------------------------------------------------------------------------- config 5
def multi_triangle_scene(bands: int = 1, per_band_grid: int = 16, seed: int = 2, res: int = 4096):
    """K = bands * per_band_grid^2 flat-coloured triangles.

    Domain: x in [0, bands] (one unit square per band / GPU), y in [0, 1]; pixel grid (res * bands) x res.
    Band b's triangles: centres on a jittered per_band_grid^2 grid of its unit square, circumradius
    U(0.03, 0.06) * 16 / per_band_grid, random rotation, counter-clockwise, colour U(0, 1),
    ``default_rng(seed + b)``.  Triangles are ordered by (class, band) with class = ((global grid row) % 3,
    j % 3): every class is a contiguous range whose pixel windows are pairwise disjoint -- the forward pass adds one
    class per launch into the image without atomics, in class order; a rank keeps only the triangles whose windows
    reach into its rows (raster.MultiTriangleRasterizer.set_scene).

    Returns dict(verts [K, 6] float64 (px0, py0, px1, py1, px2, py2), col [K], windows [K, 4] int32
    (row_begin, row_end, col_begin, col_end), classes [(begin, end)] group ranges, max_rows, max_cols)."""
    g = per_band_grid
    scale = 16.0 / g
    verts, cols, keys = [], [], []
    for b in range(bands):
        rng = np.random.default_rng(seed + b)
        jit = rng.uniform(-0.002, 0.002, size=(g, g, 2)) * scale
        rad = rng.uniform(0.03, 0.06, size=(g, g)) * scale
        rot = rng.uniform(0, 2 * np.pi, size=(g, g))
        col = rng.uniform(0, 1, size=(g, g))
        for i in range(g):
            for j in range(g):
                cx = b + (i + 0.5) / g + jit[i, j, 0]
                cy = (j + 0.5) / g + jit[i, j, 1]
                ang = rot[i, j] + np.array([0.0, 2 * np.pi / 3, 4 * np.pi / 3])
                px = cx + rad[i, j] * np.cos(ang)
                py = cy + rad[i, j] * np.sin(ang)
                verts.append([px[0], py[0], px[1], py[1], px[2], py[2]])
                cols.append(col[i, j])
                keys.append((((b * g + i) % 3) * 3 + (j % 3), b, i, j))
    order = sorted(range(len(keys)), key=lambda k: keys[k])
    verts = np.array(verts)[order]
    cols = np.array(cols)[order]
    cls = np.array([keys[k][0] for k in order])
    band = np.array([keys[k][1] for k in order])
    # values as the kernels see them (fp32-representable), so host-side windows are consistent
    verts = verts.astype(np.float32).astype(np.float64)
    cols = cols.astype(np.float32).astype(np.float64)
    xs, ys = verts[:, 0::2], verts[:, 1::2]
    rows_total = res * bands
    r0 = np.clip(np.floor(xs.min(axis=1) * res).astype(np.int64) - 1, 0, rows_total)
    r1 = np.clip(np.ceil(xs.max(axis=1) * res).astype(np.int64) + 1, 0, rows_total)
    c0 = np.clip(np.floor(ys.min(axis=1) * res).astype(np.int64) - 1, 0, res)
    c1 = np.clip(np.ceil(ys.max(axis=1) * res).astype(np.int64) + 1, 0, res)
    windows = np.stack([r0, r1, c0, c1], axis=1).astype(np.int32)
    # groups are ordered by (class, band): every class is ONE contiguous, window-disjoint range over all bands (the class
    # of a triangle comes from its GLOBAL grid row, so the 3-row spacing also holds across band boundaries): nine
    # accumulating launches whatever the number of bands
    classes = []
    for c in range(9):
        idx = np.nonzero(cls == c)[0]
        if idx.size:
            classes.append((int(idx[0]), int(idx[-1]) + 1))
    band_ranges = [(0, len(order))] * bands           # (kept for callers that pass an `active` range: the windows decide)
    return {'verts': verts, 'col': cols, 'windows': windows, 'classes': classes, 'band_ranges': band_ranges,
            'max_rows': int((r1 - r0).max()), 'max_cols': int((c1 - c0).max())}


def windows_disjoint(windows: np.ndarray) -> bool:
    """True if no two pixel windows [row_begin, row_end) x [col_begin, col_end) overlap."""
    w = np.asarray(windows)
    for a in range(len(w)):
        for b in range(a + 1, len(w)):
            if (w[a, 0] < w[b, 1] and w[b, 0] < w[a, 1] and w[a, 2] < w[b, 3] and w[b, 2] < w[a, 3]):
                return False
    return True
