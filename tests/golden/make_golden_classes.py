"""Generate tests/golden/reference_classes.npz by EXECUTING THE REFERENCE'S OWN CLASS-LEVEL
SOURCE FILES, unmodified, from /root/reference (build container only):

    src/dLux/wavefronts.py, psfs.py, coordinates.py, parametric.py, spectra.py, sources.py,
    optical_systems.py, layers/optical_layers.py, layers/optics.py, layers/propagators.py
    and utils/{helpers,math,coordinates,propagation,units,source,array_ops}.py

``python tests/golden/make_golden_classes.py``

JAX / equinox / zodiax are not installable here (no network, no wheels).  As in
make_golden.py the reference runs on a NumPy-backed ``jax`` stand-in; this script adds the
two structural dependencies of the class layer, neither of which does arithmetic:

* ``zodiax.Base``  -> a plain Python object with the functional update methods the reference
  calls (``set / add / multiply / divide / get``), each returning a modified shallow copy;
* ``equinox``      -> ``filter_vmap`` as a Python loop over the leading axis that stacks the
  leaves of the returned objects (a "vectorised" Wavefront, as vmap returns), ``field`` and
  ``filter_jit`` as no-ops.

JAX's dtype rules that NumPy 2 does not share are restated in the stand-in: default float32
(``float`` -> float32, ``complex`` -> complex64; float64 when ``X64`` is set, which is JAX's
``jax_enable_x64``), ``int32 (op) float`` -> that float type (NumPy would give float64), and the
lerp form of ``linspace``.

What the golden file pins: tilt sign and axis convention, layer order, normalisation, spectrum
and flux weighting, the unit conversions of ``to_focus`` (including the Cartesian system's
quirk of not passing its focal length, optical_systems.py:771-775), ``PointSources`` /
``BinarySource`` / ``ResolvedSource`` / ``PointResolvedSource`` / ``Scene`` sums -- i.e.
``OpticalSystem.propagate`` and ``*Source.model`` end to end -- plus, in x64 mode,
central-difference gradients of a loss through the executed reference (what ``jax.grad``
differentiates), which pin the autograd twin.  What it cannot pin: XLA's own exp / dot /
reduction kernels (ulp-level, summation order).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from collections import OrderedDict

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402

REF = MG.REF
X64 = False          # jax_enable_x64


def F():
    return onp.float64 if X64 else onp.float32


def C():
    return onp.complex128 if X64 else onp.complex64


# ------------------------------------------------------------------ jax.numpy stand-in
class WeakInt(onp.ndarray):
    """An integer array with JAX's promotion: int (op) float -> the float's type (NumPy
    promotes int32 with float32 or a Python float to float64)."""

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        ftype = None
        for x in inputs:
            if isinstance(x, WeakInt):
                continue
            if isinstance(x, float):
                ftype = ftype or F()
            elif isinstance(x, (onp.ndarray, onp.generic)) and x.dtype.kind in "fc":
                ftype = x.real.dtype.type
        if ufunc is onp.true_divide and ftype is None:
            ftype = F()
        conv = []
        for x in inputs:
            if isinstance(x, WeakInt):
                x = x.view(onp.ndarray)
                if ftype is not None:
                    x = x.astype(ftype)
            conv.append(x)
        return getattr(ufunc, method)(*conv, **kw)


def _dt(dtype):
    if dtype is float:
        return F()
    if dtype is complex:
        return C()
    return dtype


def _asarray(x, dtype=None, **kw):
    dtype = _dt(dtype)
    if isinstance(x, WeakInt) and dtype is None:
        return x
    a = onp.asarray(x, dtype=dtype)
    if dtype is None and not X64:
        if a.dtype == onp.float64:
            a = a.astype(onp.float32)
        elif a.dtype == onp.complex128:
            a = a.astype(onp.complex64)
    if dtype is None and a.dtype == onp.int64:
        a = a.astype(onp.int32)
    return a


def _arange(*a, dtype=None):
    out = onp.arange(*a, dtype=_dt(dtype))
    if out.dtype.kind == "i":
        return out.astype(onp.int32).view(WeakInt)
    return out


def _creator(fn):
    def make(shape, dtype=None):
        return fn(shape, dtype=F() if dtype in (None, float) else _dt(dtype))
    return make


def _like(fn):
    def make(a, dtype=None):
        return fn(onp.asarray(a), dtype=_dt(dtype))
    return make


def _linspace(start, stop, num=50, endpoint=True, dtype=None, axis=0):
    if not X64:
        return MG._linspace(start, stop, num, endpoint, dtype, axis)
    start, stop = onp.float64(start), onp.float64(stop)
    if num == 1:
        return onp.array([start])
    step = onp.arange(num - 1, dtype=onp.float64) / onp.float64(num - 1)
    return onp.concatenate([start * (1.0 - step) + stop * step, stop[None]])


def _log(x):
    return onp.log(onp.asarray(x, dtype=F())) if isinstance(x, (int, float)) else onp.log(x)


def _vmap(fn, in_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(len(a) for a, ax in zip(args, axes) if ax is not None)
        return _stack([fn(*[a if ax is None else a[i] for a, ax in zip(args, axes)]) for i in range(n)])
    return mapped


def _tree_map(f, tree, *rest, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    if isinstance(tree, dict):
        return type(tree)((k, _tree_map(f, v, *[r[k] for r in rest], is_leaf=is_leaf)) for k, v in tree.items())
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(f, v, *[r[i] for r in rest], is_leaf=is_leaf) for i, v in enumerate(tree))
    return f(tree, *rest)


def install_jax():
    from scipy import signal as sps
    jnp = types.ModuleType("jax.numpy")
    jnp.__getattr__ = lambda name: getattr(onp, name)      # anything dtype-neutral comes from NumPy
    jnp.asarray = jnp.array = _asarray
    jnp.arange = _arange
    jnp.zeros, jnp.ones = _creator(onp.zeros), _creator(onp.ones)
    jnp.zeros_like, jnp.ones_like = _like(onp.zeros_like), _like(onp.ones_like)
    jnp.linspace = _linspace
    jnp.log = _log
    jnp.ndarray = onp.ndarray
    jax = types.ModuleType("jax")
    jax.numpy, jax.Array, jax.vmap = jnp, MG.Array, _vmap
    jax.lax = types.ModuleType("jax.lax")
    jsp = types.ModuleType("jax.scipy")
    jsp.signal = types.ModuleType("jax.scipy.signal")

    def convolve(a, b, mode="full", method="auto"):
        # jax.scipy.signal.convolve: direct sum in the input precision
        return sps.convolve(onp.asarray(a), onp.asarray(b), mode=mode, method="direct").astype(onp.asarray(a).dtype)
    jsp.signal.convolve = convolve
    jax.scipy = jsp
    tree = types.ModuleType("jax.tree")
    tree.map = _tree_map

    def flatten(t):
        leaves = []
        _tree_map(leaves.append, t)
        return leaves, None
    tree.flatten = flatten
    jax.tree = tree
    for k, v in {"jax": jax, "jax.numpy": jnp, "jax.lax": jax.lax, "jax.scipy": jsp,
                 "jax.scipy.signal": jsp.signal, "jax.tree": tree}.items():
        sys.modules[k] = v


# ------------------------------------------------------------------ zodiax / equinox stand-ins
class Base:
    """zodiax.Base: attribute container with functional updates (no arithmetic of its own)."""

    def _clone(self):
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)
        return new

    def _update(self, parameters, values, kwargs, fn):
        if parameters is None:
            parameters, values = list(kwargs.keys()), list(kwargs.values())
        elif isinstance(parameters, str):
            parameters, values = [parameters], [values]
        new = self._clone()
        for p, v in zip(parameters, values):
            if "." in p:
                raise NotImplementedError("nested paths are not needed by the executed sources")
            new.__dict__[p] = fn(self.__dict__.get(p), v)
        return new

    def get(self, parameter):
        return getattr(self, parameter)

    def set(self, parameters=None, values=None, **kwargs):
        return self._update(parameters, values, kwargs, lambda old, v: v)

    def add(self, parameters=None, values=None, **kwargs):
        return self._update(parameters, values, kwargs, lambda old, v: old + v)

    def multiply(self, parameters=None, values=None, **kwargs):
        return self._update(parameters, values, kwargs, lambda old, v: old * v)

    def divide(self, parameters=None, values=None, **kwargs):
        return self._update(parameters, values, kwargs, lambda old, v: old / v)


def _stack(outs):
    o0 = outs[0]
    if isinstance(o0, Base):
        new = o0._clone()
        for k in o0.__dict__:
            new.__dict__[k] = _stack([o.__dict__[k] for o in outs])
        return new
    if isinstance(o0, (onp.ndarray, onp.generic)):
        return onp.stack([onp.asarray(o) for o in outs])
    if isinstance(o0, tuple):
        return tuple(_stack(list(o)) for o in zip(*outs))
    if isinstance(o0, dict):
        return type(o0)((k, _stack([o[k] for o in outs])) for k in o0)
    return o0       # static leaves (None, str, int, bool)


def install_zodiax_equinox():
    zdx = types.ModuleType("zodiax")
    zdx.Base = Base
    eqx = types.ModuleType("equinox")
    eqx.field = lambda **kw: None
    eqx.filter_jit = lambda fn: fn
    eqx.Module = Base

    def filter_vmap(fn):
        def mapped(*args):
            n = len(args[0])
            return _stack([fn(*[a[i] for a in args]) for i in range(n)])
        return mapped
    eqx.filter_vmap = filter_vmap
    sys.modules["zodiax"], sys.modules["equinox"] = zdx, eqx


# ------------------------------------------------------------------ load the reference files
def _load(modname, relpath, pkg=None):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    if pkg is not None:
        setattr(pkg, modname.rsplit(".", 1)[1], mod)
        if pkg.__name__ in ("dLux.utils", "dLux.layers", "dLux"):
            for sym in getattr(mod, "__all__", []):       # dLux/_exports.py:reexport
                setattr(pkg, sym, getattr(mod, sym))
    return mod


def load_reference(x64=False):
    """Fresh import of the reference's class layer on the stand-ins; returns the dLux package."""
    global X64
    X64 = bool(x64)
    for k in [k for k in sys.modules if k == "dLux" or k.startswith("dLux.") or k == "jax" or k.startswith("jax.")
              or k in ("zodiax", "equinox")]:
        del sys.modules[k]
    install_jax()
    install_zodiax_equinox()
    dl = types.ModuleType("dLux")
    dl.__path__ = []
    utils = types.ModuleType("dLux.utils")
    utils.__path__ = []
    layers = types.ModuleType("dLux.layers")
    layers.__path__ = []
    dl.utils, dl.layers = utils, layers
    sys.modules.update({"dLux": dl, "dLux.utils": utils, "dLux.layers": layers})
    for name in ("helpers", "math", "coordinates", "propagation", "units", "source", "array_ops"):
        _load(f"dLux.utils.{name}", f"utils/{name}.py", utils)
    # only what the executed paths touch is given a stub: fourier/interpolation need abcdLux / interpax
    for name in ("coordinates", "psfs", "parametric", "wavefronts"):
        _load(f"dLux.{name}", f"{name}.py", dl)
    for name in ("optical_layers", "optics", "propagators"):
        _load(f"dLux.layers.{name}", f"layers/{name}.py", layers)
    for name in ("spectra", "sources", "optical_systems"):
        _load(f"dLux.{name}", f"{name}.py", dl)
    return dl


# ------------------------------------------------------------------ cases
def as_f(x):
    return onp.asarray(x, dtype=F())


def build_angular(dl, cfg, coefficients=None, layer="basis"):
    co = cfg["coefficients"] if coefficients is None else coefficients
    if layer == "basis":
        lay = dl.layers.BasisOptic(as_f(cfg["basis"]), as_f(cfg["transmission"]), as_f(co), normalise=True)
    else:
        lay = layer
    return dl.AngularOpticalSystem(cfg["wf_npixels"], cfg["diameter"], [("pupil", lay)], cfg["psf_npixels"],
                                   cfg["psf_pixel_scale"], cfg["oversample"])


def small_config():
    """64 -> 32 px, 3 wavelengths, 5 smooth OPD modes: the case differentiated by finite differences."""
    rng = onp.random.default_rng(77)
    n, d = 64, 1.2
    x = (onp.arange(n) - (n - 1) / 2) * (d / n)
    X, Y = onp.meshgrid(x, x)
    T = (onp.hypot(X, Y) <= d / 2).astype(onp.float32)
    R2 = (X ** 2 + Y ** 2) / (d / 2) ** 2
    basis = onp.stack([2 * R2 - 1, X / (d / 2), Y / (d / 2), (X ** 2 - Y ** 2) / (d / 2) ** 2,
                       2 * X * Y / (d / 2) ** 2]).astype(onp.float32) * T * onp.float32(1e-9)
    return dict(wf_npixels=n, diameter=d, psf_npixels=32, psf_pixel_scale=0.08, oversample=1,
                transmission=T, basis=basis, coefficients=(30 * rng.standard_normal(5)).astype(onp.float32),
                wavelengths=onp.array([0.9e-6, 1.0e-6, 1.15e-6], onp.float32),
                weights=onp.array([0.5, 1.0, 0.75], onp.float32),
                position=onp.array([2.0e-7, -1.2e-7], onp.float32), flux=onp.float32(3.0),
                G=rng.standard_normal((32, 32)).astype(onp.float32),
                phase=(0.3 * rng.standard_normal((n, n))).astype(onp.float32),
                opd=(2e-8 * rng.standard_normal((n, n))).astype(onp.float32))


def main():
    from dlux_b200 import workloads
    out = {}

    dl = load_reference(x64=False)
    arcsec = sys.modules["dLux.utils.units"].arcsec2rad

    # ---- C1 exactly as SURVEY 8(d): 256 px circular aperture, Zernike 4-13 (seed 0), 1 wavelength -> 128
    c1 = workloads.config("c1")
    optics = build_angular(dl, c1)
    src = dl.PointSource(as_f(c1["wavelengths"]), as_f([0.0, 0.0]), 1.0)
    out["c1_psf"] = onp.asarray(optics.model(src), onp.float32)
    wf = optics.propagate_mono(as_f(c1["wavelengths"])[0], return_wf=True)
    out["c1_field"] = onp.asarray(wf.phasor, onp.complex64)
    out["c1_pixel_scale"] = onp.float64(wf.pixel_scale)

    # ---- C2-shaped: 512 -> 256, 32 wavelengths, offset source (0.3, -0.2) px
    c2 = workloads.config("c2")
    optics = build_angular(dl, c2)
    pos = as_f([0.3, -0.2]) * arcsec(as_f(c2["psf_pixel_scale"]))
    out["c2_position"] = onp.asarray(pos, onp.float32)
    src = dl.PointSource(as_f(c2["wavelengths"]), pos, 1.0, weights=as_f(c2["weights"]))
    out["c2_psf"] = onp.asarray(optics.model(src), onp.float32)

    # ---- small config: everything else
    sm = small_config()
    for k in ("transmission", "basis", "coefficients", "wavelengths", "weights", "position", "G", "phase", "opd"):
        out[f"sm_{k}"] = sm[k]
    out["sm_scalars"] = onp.array([sm["wf_npixels"], sm["diameter"], sm["psf_npixels"], sm["psf_pixel_scale"],
                                   sm["oversample"], float(sm["flux"])], onp.float64)
    optics = build_angular(dl, sm)
    src = dl.PointSource(as_f(sm["wavelengths"]), as_f(sm["position"]), sm["flux"], weights=as_f(sm["weights"]))
    out["sm_point_psf"] = onp.asarray(optics.model(src), onp.float32)
    wfs = optics.model(src, return_wf=True)
    out["sm_point_fields"] = onp.asarray(wfs.phasor, onp.complex64)            # [L, M, M], sqrt(w) applied
    psf_obj = optics.model(src, return_psf=True)
    out["sm_point_psf_pixel_scale"] = onp.float64(psf_obj.pixel_scale)
    # default weights / no offset through OpticalSystem.propagate
    out["sm_propagate_default"] = onp.asarray(optics.propagate(as_f(sm["wavelengths"])), onp.float32)

    # three stars
    positions = as_f([[2.0e-7, -1.2e-7], [-3.1e-7, 0.4e-7], [0.0, 2.5e-7]])
    fluxes = as_f([1.0, 2.5, 0.3])
    out["sm_positions"], out["sm_fluxes"] = positions, fluxes
    stars = dl.PointSources(as_f(sm["wavelengths"]), positions, fluxes, weights=as_f(sm["weights"]))
    out["sm_stars_psf"] = onp.asarray(optics.model(stars), onp.float32)
    out["sm_stars_fields"] = onp.asarray(optics.model(stars, return_wf=True).phasor, onp.complex64)   # [S, L, M, M]

    # Optic with transmission + opd + phase, explicit Tilt / Normalise layers
    optic = dl.layers.Optic(as_f(sm["transmission"]), as_f(sm["opd"]), as_f(sm["phase"]), normalise=True)
    optics2 = build_angular(dl, sm, layer=optic)
    out["sm_optic_psf"] = onp.asarray(optics2.model(src), onp.float32)
    tilt = as_f([1.5e-7, 0.8e-7])
    out["sm_tilt_angles"] = tilt
    lay = dl.LayeredOpticalSystem(sm["wf_npixels"], sm["diameter"], [
        ("t", dl.layers.TransmissiveLayer(as_f(sm["transmission"]))),
        ("a", dl.layers.AberratedLayer(as_f(sm["opd"]), as_f(sm["phase"]))),
        ("tilt", dl.layers.Tilt(tilt)),
        ("n", dl.layers.Normalise()),
        ("mft", dl.layers.MFT(40, arcsec(as_f(0.07)))),
    ])
    out["sm_layered_psf"] = onp.asarray(lay.propagate(as_f(sm["wavelengths"]), as_f(sm["position"]),
                                                      as_f(sm["weights"])), onp.float32)
    # amplitude-effect basis layer
    amp = dl.layers.BasisOptic(as_f(sm["basis"]) * as_f(1e7), as_f(sm["transmission"]), as_f(sm["coefficients"]),
                               normalise=True, effect="amplitude")
    out["sm_amplitude_psf"] = onp.asarray(build_angular(dl, sm, layer=amp).model(src), onp.float32)

    # Cartesian system (the focal length is stored but NOT passed to propagate: optical_systems.py:771-775)
    cart = dl.CartesianOpticalSystem(sm["wf_npixels"], sm["diameter"],
                                     [("pupil", dl.layers.BasisOptic(as_f(sm["basis"]), as_f(sm["transmission"]),
                                                                     as_f(sm["coefficients"]), normalise=True))],
                                     2.5, sm["psf_npixels"], 0.4, 2)
    out["sm_cartesian_psf"] = onp.asarray(cart.model(src), onp.float32)

    # BinarySource, ResolvedSource, PointResolvedSource, Scene
    w2 = as_f(onp.stack([sm["weights"], sm["weights"][::-1]]))
    out["sm_w2"] = w2
    binary = dl.BinarySource(as_f(sm["wavelengths"]), as_f(sm["position"]), 2.0, 4.0e-7, 0.7, 3.0, weights=w2)
    out["sm_binary_psf"] = onp.asarray(optics.model(binary), onp.float32)
    rng = onp.random.default_rng(5)
    dist = as_f(rng.uniform(0, 1, (5, 5)))
    out["sm_distribution"] = dist
    resolved = dl.ResolvedSource(as_f(sm["wavelengths"]), as_f(sm["position"]), 1.7, dist, weights=as_f(sm["weights"]))
    out["sm_resolved_psf"] = onp.asarray(optics.model(resolved), onp.float32)
    pres = dl.PointResolvedSource(as_f(sm["wavelengths"]), as_f(sm["position"]), 1.7, dist, 5.0, weights=w2)
    out["sm_point_resolved_psf"] = onp.asarray(optics.model(pres), onp.float32)
    scene = dl.Scene([("a", src), ("b", stars)])
    out["sm_scene_psf"] = onp.asarray(optics.model(scene), onp.float32)

    # ---- x64: central differences of loss = sum(G * psf) through the executed reference
    dl = load_reference(x64=True)
    G = sm["G"].astype(onp.float64)

    def loss(coefficients, position, flux, weights):
        optics = build_angular(dl, sm, coefficients=coefficients)
        src = dl.PointSource(as_f(sm["wavelengths"]), as_f(position), flux, weights=as_f(weights))
        return float((optics.model(src) * G).sum())

    c0 = sm["coefficients"].astype(onp.float64)
    p0 = sm["position"].astype(onp.float64)
    f0 = float(sm["flux"])
    w0 = sm["weights"].astype(onp.float64)
    out["sm_x64_psf"] = onp.asarray(build_angular(dl, sm).model(
        dl.PointSource(as_f(sm["wavelengths"]), as_f(p0), f0, weights=as_f(w0))), onp.float64)
    out["sm_x64_loss"] = onp.float64(loss(c0, p0, f0, w0))

    def central(f, x0, h):
        g = onp.zeros_like(x0)
        for i in range(x0.size):
            e = onp.zeros_like(x0)
            e.flat[i] = h
            g.flat[i] = (f(x0 + e) - f(x0 - e)) / (2 * h)
        return g

    out["sm_fd_grad_coefficients"] = central(lambda c: loss(c, p0, f0, w0), c0, 1e-3)
    out["sm_fd_grad_position"] = central(lambda p: loss(c0, p, f0, w0), p0, 1e-11)
    out["sm_fd_grad_flux"] = onp.float64((loss(c0, p0, f0 + 1e-4, w0) - loss(c0, p0, f0 - 1e-4, w0)) / 2e-4)
    # NB: Spectrum normalises the weights (spectra.py:89-92, 113-117): this is the gradient through that
    out["sm_fd_grad_weights"] = central(lambda w: loss(c0, p0, f0, w), w0, 1e-5)

    path = os.path.join(HERE, "reference_classes.npz")
    onp.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays;", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
