"""Program definitions for the BASELINE.json configs, built with the REFERENCE frontend.

TEST INFRASTRUCTURE: this module needs the ``teg`` package (only importable where
/root/reference exists).  It is *input data* for the oracle and for the golden fixtures:
frontend calls only (language nodes, reverse_deriv / FwdDeriv, reduce_to_base), no backend
code.  Sources of each scene:

  cfg1  README example                      /root/reference/README.md:21-28
  cfg2  triangle per-pixel area + grads     SURVEY.md §8(d) / Appendix D (tests/rasterize.py is unrunnable);
        raster_theta: tests/rasterize.py:22-40 with TegVars + pixel-box Vars
  cfg3  shaded triangle (k=10 grads, dL)    SURVEY.md §8(d); render_test.py scenes
        /root/reference/tests/render_test.py:30-243
  cfg4  nested 2-D / 3-D affine discontinuities, value + grads   SURVEY.md §8(d)
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple


def _teg():
    import teg  # noqa: F401  (reference frontend)
    from teg import Const, Var, TegVar, IfElse, Teg, Tup
    from teg.derivs import FwdDeriv
    from teg.derivs.reverse_deriv import reverse_deriv
    from teg.passes.reduce import reduce_to_base
    return dict(Const=Const, Var=Var, TegVar=TegVar, IfElse=IfElse, Teg=Teg, Tup=Tup,
                FwdDeriv=FwdDeriv, reverse_deriv=reverse_deriv, reduce_to_base=reduce_to_base)


def _grad(expr, params, out=None):
    t = _teg()
    out = t['Tup'](t['Const'](1)) if out is None else out
    return t['reduce_to_base'](t['reverse_deriv'](expr, out, output_list=params)[1])


# Each builder returns (program, arglist).  Builders are independent: the caller resets
# Var.global_uid before each one so labels are reproducible.

def readme_primal():
    t = _teg()
    x, tv = t['TegVar']('x'), t['Var']('t', 0.5)
    return t['Teg'](0, 1, t['IfElse'](x < tv, 1, 0), x), [tv]


def readme_deriv():
    t = _teg()
    x, tv = t['TegVar']('x'), t['Var']('t', 0.5)
    expr = t['Teg'](0, 1, t['IfElse'](x < tv, 1, 0), x)
    return t['FwdDeriv'](expr, [(tv, 1)]), [tv]


def _triangle(shaded: bool):
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    vs = [(Var(f'px{i}'), Var(f'py{i}')) for i in range(3)]

    def edge(a, b):
        return (b[0] - a[0]) * (y - a[1]) - (b[1] - a[1]) * (x - a[0])

    inside = (edge(vs[0], vs[1]) > 0) & (edge(vs[1], vs[2]) > 0) & (edge(vs[2], vs[0]) > 0)
    verts = [v for p in vs for v in p]
    if shaded:
        c0, c1, c2, bg = Var('c0'), Var('c1'), Var('c2'), Var('bg')
        body = IfElse(inside, c0 + c1 * x + c2 * y, bg)
        shader = [c0, c1, c2, bg]
    else:
        body = IfElse(inside, 1, 0)
        shader = []
    primal = Teg(box[0], box[1], Teg(box[2], box[3], body, y), x)
    return primal, box, verts, shader


def tri_primal():
    primal, box, verts, _ = _triangle(False)
    return primal, box + verts


def tri_grad():
    primal, box, verts, _ = _triangle(False)
    return _grad(primal, verts), box + verts


def shaded_primal():
    primal, box, verts, shader = _triangle(True)
    return primal, box + verts + shader


def shaded_grad():
    t = _teg()
    primal, box, verts, shader = _triangle(True)
    dL = t['Var']('dL')
    return _grad(primal, verts + shader, out=t['Tup'](dL)), [dL] + box + verts + shader


def _colored_triangle():
    """config 5 building block: one flat-coloured triangle, `col` inside, 0 outside (SURVEY §8(d) cfg 5)."""
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    vs = [(Var(f'px{i}'), Var(f'py{i}')) for i in range(3)]
    col = Var('col')

    def edge(a, b):
        return (b[0] - a[0]) * (y - a[1]) - (b[1] - a[1]) * (x - a[0])

    inside = (edge(vs[0], vs[1]) > 0) & (edge(vs[1], vs[2]) > 0) & (edge(vs[2], vs[0]) > 0)
    primal = Teg(box[0], box[1], Teg(box[2], box[3], IfElse(inside, col, 0), y), x)
    return primal, box, [v for p in vs for v in p], col


def tri_color_primal():
    primal, box, verts, col = _colored_triangle()
    return primal, box + verts + [col]


def tri_color_grad():
    """reverse derivative w.r.t. the 6 vertex coordinates and the colour, with an upstream per-pixel dL"""
    t = _teg()
    primal, box, verts, col = _colored_triangle()
    dL = t['Var']('dL')
    return _grad(primal, verts + [col], out=t['Tup'](dL)), [dL] + box + verts + [col]


def raster_theta():
    """tests/rasterize.py:22-40 (corner (0.7,0.7)) with TegVars and pixel-box Vars, FwdDeriv wrt theta."""
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    theta = Var('theta', 0)
    xl, xu, yl, yu = Var('xl'), Var('xu'), Var('yl'), Var('yu')
    x0, y0 = 0.7, 0.7
    cond = (y < y0) & (x < x0 + theta) & (x - x0 + y - y0 + 0.75 + theta > 0)
    integral = Teg(yl, yu, Teg(xl, xu, IfElse(cond, 1, 0), x), y)
    return t['FwdDeriv'](integral, [(theta, 1)]), [xl, xu, yl, yu, theta]


def raster_theta_primal():
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    theta = Var('theta', 0)
    xl, xu, yl, yu = Var('xl'), Var('xu'), Var('yl'), Var('yu')
    x0, y0 = 0.7, 0.7
    cond = (y < y0) & (x < x0 + theta) & (x - x0 + y - y0 + 0.75 + theta > 0)
    return Teg(yl, yu, Teg(xl, xu, IfElse(cond, 1, 0), x), y), [xl, xu, yl, yu, theta]


def nested2d_value():
    expr, box, coef = _nested2d()
    return expr, box + coef


def nested2d_grad():
    expr, box, coef = _nested2d()
    return _grad(expr, coef), box + coef


def _nested2d():
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    box = [Var('l0'), Var('h0'), Var('l1'), Var('h1')]
    a, b, c, e, f, g = [Var(n) for n in 'abcefg']
    body = IfElse(a * x + b * y + c < 0, 1, 0) * IfElse(e * x + f * y + g < 0, 2, 1)
    return Teg(box[0], box[1], Teg(box[2], box[3], body, y), x), box, [a, b, c, e, f, g]


def _nested3d():
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y, z = TegVar('x'), TegVar('y'), TegVar('z')
    box = [Var('l0'), Var('h0'), Var('l1'), Var('h1'), Var('l2'), Var('h2')]
    a, b, c, d, e, f, g = [Var(n) for n in 'abcdefg']
    body = IfElse(a * x + b * y + c * z + d < 0, 1, 0) * IfElse(e * x + f * y + g < 0, 2, 1)
    expr = Teg(box[0], box[1], Teg(box[2], box[3], Teg(box[4], box[5], body, z), y), x)
    return expr, box, [a, b, c, d, e, f, g]


def nested3d_value():
    expr, box, coef = _nested3d()
    return expr, box + coef


def nested3d_grad():
    expr, box, coef = _nested3d()
    return _grad(expr, coef), box + coef


# ---- render_test.py scenes (functional coverage of translate / polar / hyperbolic handlers) ----

def simple_line():
    """render_test.py:30-53"""
    t = _teg()
    from teg.maps.transform import translate
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    t1, t2 = Var('t1'), Var('t2')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    translate_map, (x_, y_) = translate([x, y], [t1, t2])
    integral = Teg(box[0], box[1], Teg(box[2], box[3], translate_map(IfElse(x_ + y_ > 0.75, 2, 1)), y), x)
    return t['reduce_to_base'](integral), box + [t1, t2]


def simple_line_grad():
    t = _teg()
    from teg.maps.transform import translate
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    t1, t2 = Var('t1'), Var('t2')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    translate_map, (x_, y_) = translate([x, y], [t1, t2])
    integral = Teg(box[0], box[1], Teg(box[2], box[3], translate_map(IfElse(x_ + y_ > 0.75, 2, 1)), y), x)
    return _grad(integral, [t1, t2]), box + [t1, t2]


def _circle():
    t = _teg()
    from teg.maps.polar import polar_2d_map
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y, r = TegVar('x'), TegVar('y'), TegVar('r')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    tv = Var('t')
    integral = Teg(box[2], box[3],
                   Teg(box[0], box[1], polar_2d_map(IfElse(r < tv, 1, 0.5), x=x, y=y, r=r), x), y)
    return integral, box, tv


def circle():
    """render_test.py:55-92 (primal)"""
    t = _teg()
    integral, box, tv = _circle()
    return t['reduce_to_base'](integral), box + [tv]


def circle_dt():
    """render_test.py:72 (reverse derivative wrt the radius threshold)"""
    integral, box, tv = _circle()
    return _grad(integral, [tv], out=None), box + [tv]


def _rect_hyperbola():
    t = _teg()
    from teg.maps.transform import scale, translate
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    tv = Var('t')
    t1, t2, t3, t4 = Var('t1'), Var('t2'), Var('t3'), Var('t4')
    scale_map, (x_s, y_s) = scale([x, y], [t1, t2])
    translate_map, (x_st, y_st) = translate([x_s, y_s], [t3, t4])
    integral = Teg(box[0], box[1],
                   Teg(box[2], box[3], scale_map(translate_map(IfElse(x_st * y_st > tv, 1, 0))), y), x)
    return integral, box, [tv, t1, t2, t3, t4]


def rect_hyperbola():
    """render_test.py:94-138 (primal)"""
    t = _teg()
    integral, box, ps = _rect_hyperbola()
    return t['reduce_to_base'](integral), box + ps


def rect_hyperbola_grad():
    integral, box, ps = _rect_hyperbola()
    return _grad(integral, ps), box + ps


def _bilinear():
    t = _teg()
    Var, TegVar, Teg, IfElse = t['Var'], t['TegVar'], t['Teg'], t['IfElse']
    x, y = TegVar('x'), TegVar('y')
    tv = Var('t')
    c00, c01, c10, c11 = Var('c00'), Var('c01'), Var('c10'), Var('c11')
    box = [Var('x_lb'), Var('x_ub'), Var('y_lb'), Var('y_ub')]
    bilinear_lerp = (c00 * (1 - x) + c01 * (x)) * (1 - y) + (c10 * (1 - x) + c11 * (x)) * (y)
    integral = Teg(box[0], box[1], Teg(box[2], box[3], IfElse(bilinear_lerp > tv, 1, 0), y), x)
    return integral, box, [tv, c00, c01, c10, c11]


def bilinear_lerp():
    """render_test.py:187-243 (primal)"""
    integral, box, ps = _bilinear()
    return integral, box + ps


def bilinear_lerp_grad():
    integral, box, ps = _bilinear()
    return _grad(integral, ps), box + ps


SCENES: Dict[str, Callable[[], Tuple[object, List]]] = {
    'readme_primal': readme_primal,
    'readme_deriv': readme_deriv,
    'tri_primal': tri_primal,
    'tri_grad': tri_grad,
    'raster_theta_primal': raster_theta_primal,
    'raster_theta': raster_theta,
    'shaded_primal': shaded_primal,
    'shaded_grad': shaded_grad,
    'tri_color_primal': tri_color_primal,
    'tri_color_grad': tri_color_grad,
    'nested2d_value': nested2d_value,
    'nested2d_grad': nested2d_grad,
    'nested3d_value': nested3d_value,
    'nested3d_grad': nested3d_grad,
    'simple_line': simple_line,
    'simple_line_grad': simple_line_grad,
    'circle': circle,
    'circle_dt': circle_dt,
    'rect_hyperbola': rect_hyperbola,
    'rect_hyperbola_grad': rect_hyperbola_grad,
    'bilinear_lerp': bilinear_lerp,
    'bilinear_lerp_grad': bilinear_lerp_grad,
}
