"""Generate tests/golden/reference_geometry.npz by EXECUTING the reference's own
src/dLux/utils/{units,coordinates,geometry,zernikes}.py on the NumPy-backed jax stand-in of
make_golden.py (build container only; see that file for the approach)."""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def _vmap(fn, in_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(len(a) for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[a if ax is None else a[i] for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(onp.stack(o) for o in zip(*outs))
        return onp.stack(outs)
    return mapped


def main():
    mods = MG.load_reference_utils()
    import jax
    jnp = jax.numpy
    for name in ("clip where maximum minimum sign roll max min cos sin".split()):
        setattr(jnp, name, getattr(onp, name))
    jax.vmap = _vmap
    base_linspace = jnp.linspace

    def linspace(start, stop, num=50, endpoint=True, **kw):
        if endpoint:
            return base_linspace(start, stop, num, **kw)
        # jax/_src/numpy/array_creation.py: same lerp with div = num and no appended stop
        F = onp.float32
        step = onp.arange(num, dtype=onp.float32) / F(num)
        return (F(start) * (F(1) - step) + F(stop) * step).astype(onp.float32)
    jnp.linspace = linspace
    jax.lax.switch = lambda idx, fns: fns[int(idx)]()
    jax.lax.reduce = None
    jax.lax.bitwise_or = None
    # zernikes.py: float32 pow / exp / lgamma / cond, equinox.filter_jit as the identity
    import types
    from scipy.special import gammaln
    for name in ("ceil floor".split()):
        setattr(jnp, name, getattr(onp, name))
    jax.lax.pow = lambda a, b: onp.power(onp.asarray(a, onp.float32), onp.asarray(b, onp.float32)).astype(onp.float32)
    jax.lax.exp = lambda x: onp.exp(onp.asarray(x, onp.float32)).astype(onp.float32)
    jax.lax.lgamma = lambda x: onp.asarray(gammaln(onp.asarray(x, onp.float32)), onp.float32)
    jax.lax.cond = lambda pred, t, f, x: t(x) if bool(pred) else f(x)
    eqx = types.ModuleType("equinox")
    eqx.filter_jit = lambda fn: fn
    sys.modules["equinox"] = eqx
    utils = sys.modules["dLux.utils"]
    for name in ("units", "geometry", "zernikes"):
        spec = importlib.util.spec_from_file_location(f"dLux.utils.{name}",
                                                      os.path.join(MG.REF, "utils", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"dLux.utils.{name}"] = mod
        if name == "geometry":
            utils.math = mods["math"]
        spec.loader.exec_module(mod)
        setattr(utils, name, mod)
        for sym in getattr(mod, "__all__", []):
            setattr(utils, sym, getattr(mod, sym))
    geo, co = sys.modules["dLux.utils.geometry"], mods["coordinates"]
    f32 = MG.f32
    out = {}
    n, diam = 48, 2.0
    coords = co.pixel_coords(n, diameter=f32(diam)).astype(onp.float32)
    out["coords"] = coords
    tr = co.translate_coords(coords, f32([0.11, -0.07]))
    sh = co.shear_coords(tr, f32([0.05, -0.02]))
    cm = co.compress_coords(sh, f32([1.1, 0.9]))
    ro = co.rotate_coords(cm, f32(0.3))
    out["translated"], out["sheared"], out["compressed"], out["rotated"] = tr, sh, cm, ro
    out["polar"] = co.cart2polar(coords)
    clip = f32(diam / n * 1.5 / 2)
    cases = {
        "soft_circle": lambda c, inv: geo.soft_circle(c, f32(0.7), clip, inv),
        "soft_square": lambda c, inv: geo.soft_square(c, f32(1.1), clip, inv),
        "soft_rectangle": lambda c, inv: geo.soft_rectangle(c, f32(1.3), f32(0.6), clip, inv),
        "soft_hexagon": lambda c, inv: geo.soft_reg_polygon(c, f32(0.8), 6, clip, inv),
        "soft_pentagon": lambda c, inv: geo.soft_reg_polygon(c, f32(0.75), 5, clip, inv),
        "soft_spider": lambda c, inv: geo.soft_spider(c, f32(0.08), f32([0.0, 120.0, 240.0]), clip, inv),
        "circle": lambda c, inv: geo.circle(c, f32(0.7), inv),
        "square": lambda c, inv: geo.square(c, f32(1.1), inv),
        "rectangle": lambda c, inv: geo.rectangle(c, f32(1.3), f32(0.6), inv),
        "hexagon": lambda c, inv: geo.reg_polygon(c, f32(0.8), 6, inv),
    }
    for name, fn in cases.items():
        for inv in (False, True):
            for cname, c in (("plain", coords), ("xf", ro.astype(onp.float32))):
                out[f"{name}_{int(inv)}_{cname}"] = onp.asarray(fn(c.copy(), inv), dtype=onp.float32)
    # a soften() on a constant array (the `cond` branch)
    out["soften_constant"] = onp.asarray(geo.soften(onp.full((4, 4), 2.0, onp.float32), f32(0.5)), onp.float32)
    # Zernike basis, Noll 1..15, on the unit-radius pupil of the same grid
    zer = sys.modules["dLux.utils.zernikes"]
    out["zernikes_1_15"] = onp.asarray(zer.zernike_basis(list(range(1, 16)), coords, f32(diam)), onp.float32)
    out["noll_nm_1_21"] = onp.array([zer.noll_indices(j) for j in range(1, 22)], onp.int64)
    # the same polynomials on n-sided regular polygons ("polikes", utils/zernikes.py:318-416), plain and on the
    # transformed coordinates
    for ns in (4, 5, 6):
        out[f"polike_{ns}_1_10"] = onp.asarray(zer.polike_basis(ns, list(range(1, 11)), coords, f32(diam)), onp.float32)
    out["polike_6_1_10_xf"] = onp.asarray(zer.polike_basis(6, list(range(1, 11)), ro.astype(onp.float32), f32(1.6)),
                                          onp.float32)
    path = os.path.join(HERE, "reference_geometry.npz")
    onp.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
