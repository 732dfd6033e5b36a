"""JAX binding of the C ABI: ``jax.ffi.ffi_call`` + ``jax.custom_vjp`` wiring the adjoint
kernel, and ``install()`` which swaps ``dLux.utils.propagation.MFT`` and the re-exported
``dLux.utils.MFT`` (src/dLux/_exports.py:12) for the CUDA version.

UNTESTED in this repository's environment: JAX / jaxlib / dLux's dependencies are not
installable there (no network).  Importing this module without JAX raises ImportError; the
tested binder is ``dlux_b200.ops`` (ctypes + torch).  See INTEGRATION.md."""
from __future__ import annotations

import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
    import numpy as np
except ImportError as e:  # pragma: no cover
    raise ImportError("integration.jax_ffi needs JAX (>= 0.4.31 for jax.ffi); it is not available "
                      "in this environment") from e

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "..", "..", "dlux_b200", "lib")
_LIB = os.path.join(_LIBDIR, "libdlux_b200_ffi.so")
_core = ctypes.CDLL(os.path.join(_LIBDIR, "libdlux_b200.so"))
_core.dlux_mft_scratch_bytes.restype = ctypes.c_size_t


def _register():
    lib = ctypes.CDLL(_LIB)
    jax.ffi.register_ffi_target("dlux_mft", jax.ffi.pycapsule(lib.dlux_mft_ffi), platform="CUDA")


class _Desc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_in", "n_out", "batch", "inverse", "adjoint", "precision")]


def _mft_call(x, scale_out, shift_xy, norm, n_in, n_out, inverse, adjoint, precision=0):
    batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    n_dst = n_in if adjoint else n_out
    d = _Desc(n_in, n_out, batch, int(inverse), int(adjoint), precision)
    nbytes = _core.dlux_mft_scratch_bytes(ctypes.byref(d))
    out_t = jax.ShapeDtypeStruct(x.shape[:-2] + (n_dst, n_dst), jnp.complex64)
    scr_t = jax.ShapeDtypeStruct((nbytes,), jnp.uint8)
    bc = lambda v, w: jnp.broadcast_to(jnp.asarray(v, jnp.float32).reshape((-1, w))[:batch] if jnp.ndim(v) else
                                       jnp.full((batch, w), v, jnp.float32), (batch, w))
    out, _ = jax.ffi.ffi_call("dlux_mft", (out_t, scr_t), vmap_method="broadcast_all")(
        x, bc(scale_out, 1)[:, 0], bc(shift_xy, 2), bc(norm, 1)[:, 0],
        n_in=np.int32(n_in), n_out=np.int32(n_out), inverse=np.int32(inverse),
        adjoint=np.int32(adjoint), precision=np.int32(precision))
    return out


def make_mft():
    """Returns a drop-in for ``dlu.MFT`` (same signature as propagation.py:178-188)."""
    _register()

    from functools import partial

    @partial(jax.custom_vjp, nondiff_argnums=(4, 5, 6))
    def _mft(x, scale_out, shift_xy, norm, n_in, n_out, inverse):
        return _mft_call(x, scale_out, shift_xy, norm, n_in, n_out, inverse, False)

    def _fwd(x, scale_out, shift_xy, norm, n_in, n_out, inverse):
        return _mft(x, scale_out, shift_xy, norm, n_in, n_out, inverse), (scale_out, shift_xy, norm)

    def _bwd(n_in, n_out, inverse, res, g):
        scale_out, shift_xy, norm = res
        # JAX cotangent convention for C -> C linear maps: conj(A^H conj(g)) = A^T g ... the adjoint
        # kernel applies A^H, so conjugate around it.
        gx = jnp.conj(_mft_call(jnp.conj(g), scale_out, shift_xy, norm, n_in, n_out, inverse, True))
        return gx, jnp.zeros_like(scale_out), jnp.zeros_like(shift_xy), jnp.zeros_like(norm)

    _mft.defvjp(_fwd, _bwd)

    def MFT(phasor, wavelength, pixel_scale_in, npixels_out, pixel_scale_out, focal_length=None,
            shift=jnp.zeros(2), pixel=True, inverse=False):
        n_in = phasor.shape[-1]
        if not pixel:
            shift = shift / pixel_scale_out
        fringe_size = wavelength / (pixel_scale_in * n_in)
        scale_out = pixel_scale_out / fringe_size
        output_size = npixels_out * pixel_scale_out
        if focal_length is not None:
            scale_out = scale_out / (focal_length + 0.0)
            output_size = output_size / (focal_length + 0.0)
        nfringes = output_size / (wavelength / (n_in * pixel_scale_in))
        norm = jnp.exp(jnp.log(nfringes) - (jnp.log(n_in) + jnp.log(npixels_out)))
        return _mft(phasor.astype(jnp.complex64), scale_out, shift, norm, n_in, int(npixels_out), bool(inverse))

    return MFT


class _PolyDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_pupil", "n_psf", "n_wavels", "n_sources", "normalise",
                                              "precision", "save_field", "reserved")]


def make_polypsf(n_pupil, n_psf, normalise=True, precision=0):
    """Fused polychromatic PSF with a custom VJP over ``dlux_polypsf_fwd/bwd``:

        psf = polypsf(transmission, opd, wavenumber, scale_out, norm, weights, delta_xy)

    transmission, opd [N, N]; wavenumber, scale_out, norm [L]; weights [S, L]; delta_xy [S, L, 2]
    (fringes).  Differentiable w.r.t. every operand (what ``jax.grad`` through
    ``OpticalSystem.propagate`` returns).  An ``OpticalSystem.propagate`` override forms the
    operands with the reference's own float32 expressions (optical_systems.py:147-223)."""
    lib = ctypes.CDLL(_LIB)
    jax.ffi.register_ffi_target("dlux_polypsf_fwd", jax.ffi.pycapsule(lib.dlux_polypsf_fwd_ffi), platform="CUDA")
    jax.ffi.register_ffi_target("dlux_polypsf_bwd", jax.ffi.pycapsule(lib.dlux_polypsf_bwd_ffi), platform="CUDA")
    _core.dlux_polypsf_scratch_bytes.restype = ctypes.c_size_t
    attrs = dict(n_pupil=np.int32(n_pupil), n_psf=np.int32(n_psf), normalise=np.int32(bool(normalise)),
                 precision=np.int32(precision))
    empty = jnp.zeros((0,), jnp.float32)

    def _scratch(L, S):
        d = _PolyDesc(n_pupil, n_psf, L, S, int(bool(normalise)), precision, 1, 0)
        return jax.ShapeDtypeStruct((_core.dlux_polypsf_scratch_bytes(ctypes.byref(d)),), jnp.uint8)

    @jax.custom_vjp
    def polypsf(transmission, opd, wavenumber, scale_out, norm, weights, delta_xy):
        return _fwd(transmission, opd, wavenumber, scale_out, norm, weights, delta_xy)[0]

    def _fwd(transmission, opd, wavenumber, scale_out, norm, weights, delta_xy):
        S, L = weights.shape
        outs = (jax.ShapeDtypeStruct((n_psf, n_psf), jnp.float32),
                jax.ShapeDtypeStruct((S * L, n_psf, n_psf), jnp.complex64), _scratch(L, S))
        psf, field, _ = jax.ffi.ffi_call("dlux_polypsf_fwd", outs)(
            transmission, opd, empty, wavenumber, scale_out, norm, weights, delta_xy, **attrs)
        return psf, (transmission, opd, wavenumber, scale_out, norm, weights, delta_xy, field)

    def _bwd(res, psf_bar):
        transmission, opd, wavenumber, scale_out, norm, weights, delta_xy, field = res
        S, L = weights.shape
        f32 = lambda *shape: jax.ShapeDtypeStruct(shape, jnp.float32)
        outs = (f32(n_pupil, n_pupil), f32(0), f32(S, L), f32(S, L, 2), f32(n_pupil, n_pupil), f32(S, L),
                f32(S, L), _scratch(L, S))
        opd_b, _, w_b, d_b, t_b, s_b, k_b, _ = jax.ffi.ffi_call("dlux_polypsf_bwd", outs)(
            transmission, opd, empty, wavenumber, scale_out, norm, weights, delta_xy, field,
            psf_bar.astype(jnp.float32), **attrs)
        norm_b = 2.0 * (weights * w_b).sum(0) / norm        # the PSF is quadratic in norm
        return t_b, opd_b, k_b.sum(0), s_b.sum(0), norm_b, w_b, d_b

    polypsf.defvjp(_fwd, _bwd)
    return polypsf


def install():
    """Monkey-patch dLux (both the defining module and the re-exported copy)."""
    import dLux.utils as dlu
    import dLux.utils.propagation as prop
    fn = make_mft()
    prop._reference_MFT = prop.MFT
    prop.MFT = fn
    dlu.MFT = fn
    return fn


def uninstall():
    import dLux.utils as dlu
    import dLux.utils.propagation as prop
    if hasattr(prop, "_reference_MFT"):
        prop.MFT = dlu.MFT = prop._reference_MFT
