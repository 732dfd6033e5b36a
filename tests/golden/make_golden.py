"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (needs /root/reference; the GPU box does not have
it):  ``python tests/golden/make_golden.py``

JAX is not installable in this image, so the reference package cannot be imported
as-is.  Its hot path, however, touches a tiny part of ``jax.numpy``
(arange/linspace/outer/exp/log/meshgrid/pad/fft/...), all of which NumPy 2 offers
with the same weak-scalar promotion rules (NEP 50 == JAX weak types: a Python float
times a float32 array stays float32; ``-2j*pi*f32`` is complex64).  This script
installs a NumPy-backed stand-in for ``jax`` into ``sys.modules`` (only
``linspace`` needs re-stating -- JAX uses the lerp form), loads the reference files

    src/dLux/utils/helpers.py, math.py, coordinates.py, propagation.py

unchanged from /root/reference, and records what the reference's own ``MFT``,
``FFT``, ``transfer_matrix``, ``calc_nfringes`` and ``nd_coords`` return on seeded
float32 inputs (the dLux default dtype) -- including the fixture of the
reference's own test (tests/utils/test_propagation.py:12-49,101-139).

What this pins: every line of the reference's arithmetic on this path, its
operation order and dtype promotion.  What it cannot pin: XLA's own exp/dot
kernels (ulp-level differences).  See DESIGN.md "Oracle".
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as onp

REF = "/root/reference/src/dLux"
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- jax stand-in
class _ArrayMeta(type):
    def __instancecheck__(cls, obj):
        return isinstance(obj, (onp.ndarray, onp.generic))


class Array(metaclass=_ArrayMeta):
    def __class_getitem__(cls, item):
        return cls


def _linspace(start, stop, num=50, endpoint=True, dtype=None, axis=0):
    """jax.numpy.linspace (jax/_src/numpy/array_creation.py): lerp form, float32
    unless inputs are float64 arrays (x64 is off by default in JAX)."""
    start = onp.asarray(start)
    stop = onp.asarray(stop)
    cdt = onp.result_type(start.dtype, stop.dtype, onp.float32)
    if cdt == onp.float64 and not (start.dtype == onp.float64 and start.ndim):
        cdt = onp.dtype(onp.float32)            # python floats are weak -> f32
    F = cdt.type
    start, stop = start.astype(cdt), stop.astype(cdt)
    assert endpoint and start.ndim == 0
    if num == 1:
        return onp.array([start], dtype=cdt)
    div = num - 1
    step = onp.arange(div, dtype=cdt) / F(div)
    out = start * (F(1) - step) + stop * step
    return onp.concatenate([out, stop[None]]).astype(cdt)


def _asarray(x, dtype=None):
    if dtype is float:
        dtype = onp.float32
    if dtype is complex:
        dtype = onp.complex64
    a = onp.asarray(x, dtype=dtype)
    if dtype is None and a.dtype == onp.float64:
        a = a.astype(onp.float32)               # JAX default: no x64
    if dtype is None and a.dtype == onp.int64:
        a = a.astype(onp.int32)
    return a


def _zeros(shape, dtype=None):
    return onp.zeros(shape, dtype=onp.float32 if dtype in (None, float) else dtype)


def _arange(*a, dtype=None):
    out = onp.arange(*a, dtype=dtype)
    return out.astype(onp.int32) if out.dtype == onp.int64 else out


def _log(x):
    # jnp.log(python int) -> float32
    return onp.log(onp.asarray(x, dtype=onp.float32)) if isinstance(x, (int, float)) else onp.log(x)


def _vmap(fn):
    def mapped(x):
        outs = [fn(xi) for xi in x]
        if isinstance(outs[0], tuple):
            return tuple(onp.stack(o) for o in zip(*outs))
        return onp.stack(outs)
    return mapped


def install_jax_standin():
    jnp = types.ModuleType("jax.numpy")
    for name in ("outer exp cos sin meshgrid squeeze pad fft pi arctan2 hypot flip "
                 "prod repeat tensordot transpose stack sqrt abs sum ones inf nan").split():
        setattr(jnp, name, getattr(onp, name))
    jnp.linspace = _linspace
    jnp.asarray = _asarray
    jnp.array = _asarray
    jnp.zeros = _zeros
    jnp.arange = _arange
    jnp.log = _log
    jnp.ndarray = onp.ndarray
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.Array = Array
    jax.vmap = _vmap
    jax.lax = types.ModuleType("jax.lax")
    jax.scipy = types.ModuleType("jax.scipy")
    tree = types.ModuleType("jax.tree")
    tree.map = lambda f, *trees: tuple(f(*xs) for xs in zip(*trees))
    jax.tree = tree
    for k, v in {"jax": jax, "jax.numpy": jnp, "jax.lax": jax.lax, "jax.scipy": jax.scipy,
                 "jax.tree": tree}.items():
        sys.modules[k] = v
    return jnp


def load_reference_utils():
    """Load the four reference source files, unmodified, as dLux.utils.*"""
    install_jax_standin()
    pkg = types.ModuleType("dLux")
    pkg.__path__ = []
    utils = types.ModuleType("dLux.utils")
    utils.__path__ = []
    pkg.utils = utils
    sys.modules["dLux"] = pkg
    sys.modules["dLux.utils"] = utils
    mods = {}
    for name in ("helpers", "math", "coordinates", "propagation"):
        spec = importlib.util.spec_from_file_location(
            f"dLux.utils.{name}", os.path.join(REF, "utils", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"dLux.utils.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
        setattr(utils, name, mod)
        for sym in getattr(mod, "__all__", []):      # what dLux/_exports.py:reexport does
            setattr(utils, sym, getattr(mod, sym))
    return mods


def f32(x):
    return onp.asarray(x, dtype=onp.float32)


def main():
    mods = load_reference_utils()
    prop, coords = mods["propagation"], mods["coordinates"]
    out = {}
    cases = []

    # (1) the reference's own test fixture: tests/utils/test_propagation.py:12-49
    ones32 = onp.ones((32, 32), dtype=onp.complex64)
    k = 0
    for focal_length in (None, f32(2.0)):
        for inverse in (False, True):
            for pixel in (True, False):
                res = prop.MFT(ones32, f32(1.0), f32(0.1), 16, f32(0.05),
                               focal_length=focal_length, shift=f32([1.0, -2.0]),
                               pixel=pixel, inverse=inverse)
                out[f"reftest_{k}"] = res.astype(onp.complex64)
                cases.append(("reftest", k, focal_length is not None, inverse, pixel))
                k += 1

    # (2) seeded random phasors at several shapes / geometries
    rng = onp.random.default_rng(20261017)
    geoms = [
        # n_in, n_out, wavelength, ps_in, ps_out, focal_length, shift, pixel, inverse
        (64, 32, 1.0e-6, 1.0 / 64, 1.2e-7, None, (0.0, 0.0), True, False),
        (64, 48, 1.3e-6, 2.4 / 64, 9.0e-8, None, (0.5, -1.25), True, False),
        (96, 40, 5.5e-7, 0.125 / 96, 6.0e-6, 1.5, (2.0, 3.0), True, True),
        (128, 128, 1.0e-6, 1.0 / 128, 1.0e-6, None, (0.0, 0.0), True, False),
        (100, 36, 2.1e-6, 6.5 / 100, 3.1e-8, None, (1.0e-7, -2.0e-7), False, False),
        (256, 128, 1.0e-6, 1.0 / 256, 2.42406841e-7, None, (0.0, 0.0), True, False),
    ]
    for g, (n_in, n_out, wl, psi, pso, fl, shift, pixel, inverse) in enumerate(geoms):
        ph = (rng.standard_normal((n_in, n_in)) + 1j * rng.standard_normal((n_in, n_in)))
        ph = (ph / n_in).astype(onp.complex64)
        res = prop.MFT(ph, f32(wl), f32(psi), n_out, f32(pso),
                       focal_length=None if fl is None else f32(fl),
                       shift=f32(shift), pixel=pixel, inverse=inverse)
        out[f"mft_{g}_in"] = ph
        out[f"mft_{g}_out"] = res.astype(onp.complex64)
        out[f"mft_{g}_geom"] = onp.array([n_in, n_out, wl, psi, pso,
                                          -1.0 if fl is None else fl, shift[0], shift[1],
                                          float(pixel), float(inverse)], dtype=onp.float64)
        tm = prop.transfer_matrix(f32(wl), n_in, f32(psi), n_out, f32(pso), f32(shift[0] if pixel else 0.0),
                                  None if fl is None else f32(fl), 0.0, inverse)
        out[f"mft_{g}_tmx"] = tm.astype(onp.complex64)
        out[f"mft_{g}_nfringes"] = onp.float32(prop.calc_nfringes(
            f32(wl), n_in, f32(psi), n_out, f32(pso), None if fl is None else f32(fl)))
    out["n_geoms"] = onp.int64(len(geoms))

    # (3) nd_coords (1-D) as the transfer matrix uses it
    for j, (n, sc, off) in enumerate([(32, 1.0 / 32, 0.03125), (16, 0.7, -1.4), (513, 0.1217, 0.06),
                                      (1024, 1.0 / 1024, 0.0), (512, 0.12207, 0.0)]):
        out[f"coords_{j}"] = coords.nd_coords(n, f32(sc), f32(off)).astype(onp.float32)
        out[f"coords_{j}_args"] = onp.array([n, sc, off], dtype=onp.float64)
    out["n_coords"] = onp.int64(5)

    # (4) FFT: tests/utils/test_propagation.py:55-92 fixture + a random case
    for j, (pad, inverse) in enumerate([(1, False), (2, False), (2, True), (3, False)]):
        ph = (rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32))).astype(onp.complex64)
        res, ps = prop.FFT(ph, f32(1.0e-6), f32(0.01), None, pad, inverse)
        out[f"fft_{j}_in"], out[f"fft_{j}_out"] = ph, res.astype(onp.complex64)
        out[f"fft_{j}_meta"] = onp.array([pad, float(inverse), float(ps)], dtype=onp.float64)
    out["n_fft"] = onp.int64(4)

    path = os.path.join(HERE, "reference_propagation.npz")
    onp.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays;", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
