"""CPU oracle for the dLux MFT diffraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dlux_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the timed CPU arm.

It is a NumPy restatement, operation by operation and in the same floating
point type (float32 / complex64, the JAX default), of these reference lines
(paths relative to /root/reference/):

* ``jnp.linspace``                        -> :func:`jnp_linspace` (JAX lerp form)
* ``src/dLux/utils/coordinates.py:329-332``   -> :func:`nd_coords_1d`
* ``src/dLux/utils/propagation.py:109-127``   -> :func:`transfer_matrix`
* ``src/dLux/utils/propagation.py:165-175``   -> :func:`calc_nfringes`
* ``src/dLux/utils/propagation.py:223-256``   -> :func:`MFT`
* ``src/dLux/utils/propagation.py:44-64``     -> :func:`FFT`
* ``src/dLux/coordinates.py:129``, ``src/dLux/wavefronts.py:111-113, 279, 303,
  330, 349, 368, 392-395, 418-424, 605-606, 762-772``  -> :class:`OracleWavefront`
* ``src/dLux/layers/optics.py:91-96,168-175``, ``src/dLux/utils/math.py:188-196``
  -> :func:`apply_optic`, :func:`eval_basis`
* ``src/dLux/optical_systems.py:185-223, 383-389, 415-425, 676-680`` ->
  :func:`propagate_mono`, :func:`propagate`
* ``src/dLux/sources.py:322-327, 398-411``, ``src/dLux/spectra.py:84-92`` ->
  :func:`point_source_model`, :func:`point_sources_model`

Parity pin: the reference cannot be imported here or on the GPU box (no JAX in
the image).  ``tests/golden/make_golden.py`` executes the reference's *own*
``utils/propagation.py`` / ``utils/coordinates.py`` source files on top of a
NumPy-backed stand-in for the tiny part of ``jax.numpy`` they use and commits the
outputs as fixtures; this oracle is checked against those fixtures and against
analytic known answers (tests/test_oracle.py).  XLA's own ``exp``/``dot`` kernels
are not in the loop, so the pin is "reference source on a NumPy substrate", not
"reference on XLA" -- see DESIGN.md.

Every function takes ``dtype`` (np.float32 default = the reference's arithmetic;
np.float64 = an error-budget twin).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "jnp_linspace", "nd_coords_1d", "transfer_matrix", "calc_nfringes", "mft_scalars",
    "MFT", "FFT", "OracleWavefront", "eval_basis", "apply_optic", "propagate_mono",
    "propagate", "point_source_model", "point_sources_model", "arcsec2rad",
    "mft_flops",
]


def _ctype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


def jnp_linspace(start, stop, n: int, dtype=np.float32) -> np.ndarray:
    """``jax.numpy.linspace(start, stop, n)``: ``start*(1-t) + stop*t`` with
    ``t = iota/(n-1)`` and the endpoint appended verbatim."""
    F = np.dtype(dtype).type
    start, stop = F(start), F(stop)
    if n == 1:
        return np.array([start], dtype=dtype)
    div = n - 1
    t = np.arange(div, dtype=dtype) / F(div)
    out = start * (F(1) - t) + stop * t
    return np.concatenate([out, np.array([stop], dtype=dtype)]).astype(dtype)


def nd_coords_1d(n: int, scale, offset, dtype=np.float32) -> np.ndarray:
    """1-D case of ``dlu.nd_coords`` (utils/coordinates.py:329-332)."""
    F = np.dtype(dtype).type
    offset = F(offset)
    h = (n - 1) / 2
    if isinstance(scale, float):
        # weak (Python) scalar, as transfer_matrix's scale_in = 1.0 / npixels_in:
        # -(n-1)/2 * scale is evaluated in Python float64 and rounded once.
        start = F(-h * scale) - offset
        end = F(h * scale) - offset
    else:
        scale = F(scale)
        start = F(F(-h) * scale) - offset
        end = F(F(h) * scale) - offset
    return jnp_linspace(start, end, n, dtype)


def mft_scalars(wavelength, npixels_in, pixel_scale_in, pixel_scale_out,
                focal_length=None, focal_shift=0.0, dtype=np.float32):
    """``scale_out`` exactly as utils/propagation.py:110,117-120 forms it."""
    F = np.dtype(dtype).type
    fringe_size = F(wavelength) / F(F(pixel_scale_in) * F(npixels_in))
    scale_out = F(pixel_scale_out) / fringe_size
    if focal_length is not None:
        scale_out = scale_out / F(F(focal_length) + F(focal_shift))
    return F(scale_out)


def transfer_matrix(wavelength, npixels_in, pixel_scale_in, npixels_out, pixel_scale_out,
                    shift=0.0, focal_length=None, focal_shift=0.0, inverse=False,
                    dtype=np.float32) -> np.ndarray:
    """utils/propagation.py:67-127.  Returns the (npixels_in, npixels_out) matrix."""
    F = np.dtype(dtype).type
    shift = F(shift)
    scale_in = 1.0 / npixels_in                      # Python float (weak), :113
    in_vec = nd_coords_1d(npixels_in, scale_in, shift * F(scale_in), dtype)
    scale_out = mft_scalars(wavelength, npixels_in, pixel_scale_in, pixel_scale_out,
                            focal_length, focal_shift, dtype)
    out_vec = nd_coords_1d(npixels_out, scale_out, shift * scale_out, dtype)
    two_pi = F(-2.0 * np.pi)
    arg = two_pi * np.outer(in_vec, out_vec).astype(dtype)
    if inverse:
        arg = -arg
    return _exp_i(arg, dtype)


def _exp_i(arg, dtype):
    """``np.exp(1j * arg)`` as a complex exp of (0, arg), like the reference."""
    z = np.zeros(arg.shape, dtype=_ctype(dtype))
    z.imag = arg
    return np.exp(z)


def calc_nfringes(wavelength, npixels_in, pixel_scale_in, npixels_out, pixel_scale_out,
                  focal_length=None, focal_shift=0.0, dtype=np.float32):
    """utils/propagation.py:130-175."""
    F = np.dtype(dtype).type
    diameter = F(npixels_in) * F(pixel_scale_in)
    fringe_size = F(wavelength) / diameter
    output_size = F(npixels_out) * F(pixel_scale_out)
    if focal_length is not None:
        output_size = output_size / F(F(focal_length) + F(focal_shift))
    return F(output_size / fringe_size)


def mft_norm(nfringes, npixels_in, npixels_out, dtype=np.float32):
    """utils/propagation.py:254: exp(log nf - (log N + log M))."""
    F = np.dtype(dtype).type
    return F(np.exp(F(np.log(F(nfringes))) - F(F(np.log(F(npixels_in))) + F(np.log(F(npixels_out))))))


def MFT(phasor, wavelength, pixel_scale_in, npixels_out, pixel_scale_out,
        focal_length=None, shift=(0.0, 0.0), pixel=True, inverse=False,
        dtype=np.float32) -> np.ndarray:
    """utils/propagation.py:178-256: ``(y_mat.T @ phasor) @ x_mat`` times the
    nfringes normalisation.  ``shift = (x, y)``; x drives the column matrix."""
    F = np.dtype(dtype).type
    phasor = np.asarray(phasor).astype(_ctype(dtype))
    npixels_in = phasor.shape[-1]
    shift = np.asarray(shift, dtype=dtype)
    if not pixel:
        shift = shift / F(pixel_scale_out)
    mats = [transfer_matrix(wavelength, npixels_in, pixel_scale_in, npixels_out,
                            pixel_scale_out, s, focal_length, 0.0, inverse, dtype)
            for s in shift]
    x_mat, y_mat = mats
    out = (y_mat.T @ phasor) @ x_mat
    nf = calc_nfringes(wavelength, npixels_in, pixel_scale_in, npixels_out,
                       pixel_scale_out, focal_length, 0.0, dtype)
    return (out * mft_norm(nf, npixels_in, npixels_out, dtype)).astype(_ctype(dtype))


def FFT(phasor, wavelength, pixel_scale, focal_length=None, pad=2, inverse=False,
        dtype=np.float32):
    """utils/propagation.py:8-64.  Returns ``(phasor, new_pixel_scale)``."""
    F = np.dtype(dtype).type
    phasor = np.asarray(phasor).astype(_ctype(dtype))
    npixels = phasor.shape[-1]
    fringe_size = F(wavelength) / F(F(pixel_scale) * F(npixels))
    new_pixel_scale = fringe_size / F(pad)
    if focal_length is not None:
        new_pixel_scale = new_pixel_scale * F(focal_length)
    npad = (npixels * (pad - 1)) // 2
    phasor = np.pad(phasor, npad)
    if inverse:
        phasor = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(phasor)))
        phasor = phasor * F(phasor.shape[-1])
    else:
        phasor = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(phasor)))
        phasor = phasor / F(phasor.shape[-1])
    return phasor.astype(_ctype(dtype)), F(new_pixel_scale)


# ----------------------------------------------------------------------------
# Callers of the hot path: wavefront state, pupil layers, optical system, sources
# ----------------------------------------------------------------------------

ARCSEC = np.pi / (180.0 * 3600.0)  # utils/units.py: _BASE_TO_RAD["arcsec"]


def arcsec2rad(v, dtype=np.float32):
    F = np.dtype(dtype).type
    return F(F(v) * F(ARCSEC))


class OracleWavefront:
    """Minimal restatement of ``dLux.wavefronts.Wavefront`` (phasor, wavelength,
    pixel_scale) with the methods on the hot path."""

    def __init__(self, wavelength, npixels, diameter, dtype=np.float32):
        F = np.dtype(dtype).type
        self.dtype = dtype
        self.wavelength = F(wavelength)
        self.pixel_scale = F(F(diameter) / F(npixels))          # wavefronts.py:107
        amplitude = np.ones((npixels, npixels), dtype=dtype) / F(npixels ** 2)
        self.phasor = amplitude.astype(_ctype(dtype))            # wavefronts.py:111-113

    @property
    def npixels(self):
        return self.phasor.shape[-1]

    @property
    def wavenumber(self):                                        # wavefronts.py:303
        F = np.dtype(self.dtype).type
        return F(F(2 * np.pi) / self.wavelength)

    @property
    def xs(self):                                                # coordinates.py:129
        F = np.dtype(self.dtype).type
        n = self.npixels
        return (np.arange(n).astype(self.dtype) - F((n - 1) / 2)) * self.pixel_scale

    @property
    def psf(self):                                               # wavefronts.py:279
        return np.abs(self.phasor) ** 2

    def _mul_exp_i(self, phase):
        self.phasor = self.phasor * _exp_i(phase, self.dtype)

    def add_phase(self, phase):                                  # wavefronts.py:349
        if phase is not None:
            self._mul_exp_i(np.asarray(phase, dtype=self.dtype))
        return self

    def add_opd(self, opd):                                      # wavefronts.py:368
        if opd is not None:
            self.add_phase(self.wavenumber * np.asarray(opd, dtype=self.dtype))
        return self

    def tilt(self, angles):                                      # wavefronts.py:388-395
        angles = np.asarray(angles, dtype=self.dtype)
        xs = self.xs
        X, Y = np.meshgrid(xs, xs)                               # wavefronts.py:605-606
        return self.add_opd(angles[0] * X + angles[1] * Y)

    def multiply(self, t):
        self.phasor = self.phasor * np.asarray(t)
        return self

    def normalise(self):                                         # wavefronts.py:418-424
        F = np.dtype(self.dtype).type
        power = np.sum(np.abs(self.phasor) ** 2, dtype=self.dtype)
        self.phasor = self.phasor * F(np.sqrt(F(1.0) / power))
        return self

    def propagate(self, npixels, pixel_scale, focal_length=None, inverse=False):
        self.phasor = MFT(self.phasor, self.wavelength, self.pixel_scale, npixels,
                          pixel_scale, focal_length, inverse=inverse, dtype=self.dtype)
        self.pixel_scale = np.dtype(self.dtype).type(pixel_scale)
        return self


def eval_basis(basis, coefficients, dtype=np.float32):
    """utils/math.py:177-196."""
    basis = np.asarray(basis, dtype=dtype)
    coefficients = np.asarray(coefficients, dtype=dtype)
    axes = tuple(range(coefficients.ndim))
    return np.tensordot(basis, coefficients, axes=(axes, axes)).astype(dtype)


def apply_optic(wf: OracleWavefront, transmission=None, opd=None, phase=None,
                basis=None, coefficients=None, normalise=True):
    """``Optic.__call__`` (layers/optics.py:91-96) / ``BasisOptic.__call__``
    (:168-175): ``*= T``; ``add_opd``; ``add_phase``; ``normalise``."""
    if transmission is not None:
        wf.multiply(np.asarray(transmission, dtype=wf.dtype))
    if basis is not None:
        wf.add_opd(eval_basis(basis, coefficients, wf.dtype))
    wf.add_opd(opd)
    wf.add_phase(phase)
    if normalise:
        wf.normalise()
    return wf


def propagate_mono(optics: dict, wavelength, offset=None, return_field=False,
                   dtype=np.float32):
    """``LayeredOpticalSystem.propagate_mono`` + ``AngularOpticalSystem.to_focus``
    (optical_systems.py:391-425, 659-680).  ``optics`` is a plain dict:
    wf_npixels, diameter, psf_npixels, psf_pixel_scale (arcsec), oversample,
    transmission, opd, phase, basis, coefficients, normalise."""
    F = np.dtype(dtype).type
    wf = OracleWavefront(wavelength, optics["wf_npixels"], optics["diameter"], dtype)
    wf.tilt(np.zeros(2) if offset is None else offset)
    apply_optic(wf, optics.get("transmission"), optics.get("opd"), optics.get("phase"),
                optics.get("basis"), optics.get("coefficients"),
                optics.get("normalise", True))
    true_pixel_scale = F(optics["psf_pixel_scale"]) / F(optics.get("oversample", 1))
    pixel_scale = arcsec2rad(true_pixel_scale, dtype)
    npix = optics["psf_npixels"] * optics.get("oversample", 1)
    wf.propagate(npix, pixel_scale)
    return wf.phasor if return_field else wf.psf


def propagate(optics: dict, wavelengths, offset=None, weights=None, return_field=False,
              dtype=np.float32):
    """``OpticalSystem.propagate`` (optical_systems.py:147-223): per-wavelength
    field times sqrt(weight), PSF = sum over wavelengths of |E|^2."""
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=dtype))
    if weights is None:
        weights = np.ones_like(wavelengths) / np.dtype(dtype).type(len(wavelengths))
    weights = np.atleast_1d(np.asarray(weights, dtype=dtype))
    if weights.shape != wavelengths.shape:
        raise ValueError("Wavelength and weight shape mismatch")
    fields = [propagate_mono(optics, wl, offset, True, dtype) * np.sqrt(w)
              for wl, w in zip(wavelengths, weights)]
    fields = np.stack(fields).astype(_ctype(dtype))
    if return_field:
        return fields
    return (np.abs(fields) ** 2).sum(0).astype(dtype)


def point_source_model(optics, wavelengths, position, flux=1.0, weights=None,
                       dtype=np.float32):
    """``PointSource.model`` (sources.py:316-327)."""
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=dtype))
    if weights is None:
        weights = np.ones(wavelengths.shape, dtype=dtype) / np.dtype(dtype).type(len(wavelengths))
    weights = np.asarray(weights, dtype=dtype)
    weights = weights / weights.sum()                            # spectra.py:113-117
    return propagate(optics, wavelengths, position, weights * np.dtype(dtype).type(flux),
                     dtype=dtype)


def point_sources_model(optics, wavelengths, positions, fluxes, weights=None,
                        dtype=np.float32):
    """``PointSources.model`` (sources.py:392-411): sum over stars and wavelengths."""
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=dtype))
    if weights is None:
        weights = np.ones(wavelengths.shape, dtype=dtype) / np.dtype(dtype).type(len(wavelengths))
    weights = np.asarray(weights, dtype=dtype)
    weights = weights / weights.sum()
    fluxes = np.asarray(fluxes, dtype=dtype)
    out = None
    for pos, fl in zip(np.asarray(positions, dtype=dtype), fluxes):
        psf = propagate(optics, wavelengths, pos, weights * fl, dtype=dtype)
        out = psf if out is None else out + psf
    return out.astype(dtype)


def mft_flops(n_in: int, n_out: int) -> float:
    """Algorithmic real FLOPs of one MFT, one direction: 8*M*N*(N+M)."""
    return 8.0 * n_out * n_in * (n_in + n_out)
