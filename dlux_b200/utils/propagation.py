"""Host-side mirror of ``dLux.utils.propagation`` for the MFT
(/root/reference/src/dLux/utils/propagation.py:67-256): same names, argument
meaning and error behaviour; the arithmetic runs in libdlux_b200.so.

Geometry scalars are formed in float32 in the reference's own operation order
(:110, :117-120, :165-175, :254) -- on the host with NumPy when they arrive as Python
/ NumPy numbers, on the device with torch when they arrive as CUDA tensors -- and are
handed to the kernels as small device arrays, never read back.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import ops

__all__ = ["MFT", "FFT", "calc_nfringes", "mft_geometry", "arcsec2rad", "eval_basis", "fft_spec",
           "fft_phase_ramp"]

_ARCSEC = math.pi / (180.0 * 3600.0)   # dLux/utils/units.py _BASE_TO_RAD["arcsec"]


def _is_dev(x) -> bool:
    return torch.is_tensor(x) and x.is_cuda


def _f(x, like=None):
    """float32 view of a scalar/array: torch on device if any input lives there."""
    if _is_dev(x):
        return x.to(torch.float32)
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


def _coerce(*xs):
    """All-NumPy float32, or all-torch float32 on the device of the first CUDA input."""
    dev = next((x.device for x in xs if _is_dev(x)), None)
    if dev is None:
        return tuple(None if x is None else _f(x) for x in xs)
    return tuple(None if x is None else (x.to(torch.float32) if _is_dev(x) else
                 torch.as_tensor(_f(x), device=dev)) for x in xs)


def arcsec2rad(values):
    """dLux.utils.units.arcsec2rad: ``values * (pi / 648000)`` in float32."""
    v = _f(values)
    if _is_dev(v):
        return v * np.float32(_ARCSEC)
    return (v * np.float32(_ARCSEC)).astype(np.float32)


def calc_nfringes(wavelength, npixels_in, pixel_scale_in, npixels_out, pixel_scale_out,
                  focal_length=None, focal_shift=0.0):
    """propagation.py:130-175."""
    wl, psi, pso, fl = _coerce(wavelength, pixel_scale_in, pixel_scale_out, focal_length)
    diameter = np.float32(npixels_in) * psi
    fringe_size = wl / diameter
    output_size = np.float32(npixels_out) * pso
    if fl is not None:
        output_size = output_size / (fl + np.float32(focal_shift))
    return output_size / fringe_size


def mft_geometry(wavelength, npixels_in, pixel_scale_in, npixels_out, pixel_scale_out,
                 focal_length=None):
    """(scale_out, norm): propagation.py:110,117-120 and :246-254, float32."""
    wl, psi, pso, fl = _coerce(wavelength, pixel_scale_in, pixel_scale_out, focal_length)
    fringe_size = wl / (psi * np.float32(npixels_in))
    scale_out = pso / fringe_size
    if fl is not None:
        scale_out = scale_out / (fl + np.float32(0.0))
    nf = calc_nfringes(wavelength, npixels_in, pixel_scale_in, npixels_out, pixel_scale_out,
                       focal_length)
    log_nm = np.float32(np.log(np.float32(npixels_in))) + np.float32(np.log(np.float32(npixels_out)))
    if _is_dev(nf):
        norm = torch.exp(torch.log(nf) - log_nm)
    else:
        norm = np.exp(np.log(nf) - log_nm).astype(np.float32)
    return scale_out, norm


def _to_dev(x, device):
    if _is_dev(x):
        return x.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=device)


def MFT(phasor, wavelength, pixel_scale_in, npixels_out, pixel_scale_out, focal_length=None,
        shift=None, pixel: bool = True, inverse: bool = False, precision=None):
    """Drop-in for ``dlu.MFT`` (propagation.py:178-256).

    ``phasor``: complex64 CUDA tensor [..., N, N]; leading dimensions are the vmapped
    batch, in which case the scalar arguments may be scalars or arrays of the batch
    shape.  ``shift`` is (x, y) in output pixels (or in ``pixel_scale_out`` units with
    ``pixel=False``), shape (2,) or [..., 2].  Differentiable w.r.t. ``phasor``.
    """
    if not torch.is_tensor(phasor):
        raise TypeError("phasor must be a torch CUDA tensor")
    if phasor.dim() < 2 or phasor.shape[-1] != phasor.shape[-2]:
        raise ValueError("phasor must be [..., N, N]")
    npixels_out = int(npixels_out)
    npixels_in = phasor.shape[-1]
    dev = phasor.device
    scale_out, norm = mft_geometry(wavelength, npixels_in, pixel_scale_in, npixels_out,
                                   pixel_scale_out, focal_length)
    if shift is not None:
        shift = _f(shift)
        if not pixel:                                # propagation.py:225-226
            pso = _f(pixel_scale_out)
            if _is_dev(shift) or _is_dev(pso):
                shift = _to_dev(shift, dev) / _to_dev(pso, dev)[..., None]
            else:
                shift = (shift / np.asarray(pso)[..., None]).astype(np.float32)
        shift = _to_dev(shift, dev)
    scale_out = _to_dev(scale_out, dev)
    norm = _to_dev(norm, dev)
    return ops.MFTFunction.apply(phasor, scale_out, npixels_out, shift, None, norm, bool(inverse),
                                 precision, False)


def FFT(phasor, wavelength, pixel_scale, focal_length=None, pad: int = 2, inverse: bool = False,
        precision=None):
    """Drop-in for ``dlu.FFT`` (propagation.py:8-64): zero-pad by ``(N*(pad-1))//2`` per side,
    ``fftshift(fft2(ifftshift(.)))/N_pad`` (inverse: ``ifft2 * N_pad``).  Returns
    ``(phasor, new_pixel_scale)``.

    The padded, centred DFT is evaluated by the same phasor-GEMM kernels as the MFT (SURVEY 8f
    NEXT-2) in their exact-DFT mode: integer index offsets from the origins ``N_pad//2 - npad``
    (input) and ``N_pad//2`` (output; numpy's fftshift convention), phases
    ``2 pi ((j - j0)(b - b0) mod N_pad) / N_pad`` reduced exactly before the sincos, normalisation
    ``1 / N_pad``.  This costs O(N^2 N_pad) instead of O(N_pad^2 log N_pad) but runs on the tensor
    cores and shares the adjoint; ``tools/fft_probe.py`` times it against cuFFT."""
    if not torch.is_tensor(phasor):
        raise TypeError("phasor must be a torch CUDA tensor")
    n = phasor.shape[-1]
    pad = int(pad)
    wl, ps, fl = _coerce(wavelength, pixel_scale, focal_length)
    fringe_size = wl / (ps * np.float32(n))
    new_pixel_scale = fringe_size / np.float32(pad)
    if fl is not None:
        new_pixel_scale = new_pixel_scale * fl
    npad = (n * (pad - 1)) // 2
    n_out = n + 2 * npad
    # exact-DFT mode of the kernels: integer index offsets from the fftshift origins (input index N_pad//2 - npad,
    # output index N_pad//2) and phases 2 pi ((j - j0)(b - b0) mod N_pad) / N_pad -- no float32 phase rounding
    j0 = float(n_out // 2 - npad)
    b0 = float(n_out // 2)
    dev = phasor.device
    batch = int(np.prod(phasor.shape[:-2])) if phasor.dim() > 2 else 1
    f = lambda v, w: torch.full((batch, w), float(v), dtype=torch.float32, device=dev)
    # dlu.FFT scales the forward transform by 1/N_pad and the inverse (ifft2 = 1/N_pad^2 inside) by N_pad:
    # both are 1/N_pad on the unnormalised DFT sum
    out = ops.MFTFunction.apply(phasor, f(1.0, 1).reshape(batch), n_out, f(j0, 2), f(b0, 2),
                                f(np.float32(1.0) / np.float32(n_out), 1), bool(inverse), (precision, n_out), False)
    return out, new_pixel_scale


def fft_spec(npixels_in, pixel_scale_in, wavelength, focal_length=None):
    """dLux.utils.fourier.fft_spec (utils/fourier.py:14-45): FFT output pixel scale and centre."""
    wl, ps, fl = _coerce(wavelength, pixel_scale_in, focal_length)
    d_out = wl / (np.float32(npixels_in) * ps)
    if fl is not None:
        d_out = d_out * fl
    if npixels_in % 2 != 0:
        return d_out, np.float32(0.0)
    return d_out, np.float32(-0.5) * d_out


def fft_phase_ramp(xs, wavelength, shift, focal_length=None, inverse=False):
    """dLux.utils.fourier.fft_phase_ramp (utils/fourier.py:48-78), torch tensors."""
    sign = -1.0 if inverse else 1.0
    to_t = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float32), device=xs.device)
    alpha = to_t(wavelength) if focal_length is None else to_t(wavelength) * to_t(focal_length)
    ang = (np.float32(sign * 2 * math.pi) * xs * to_t(shift) / alpha).to(torch.float32)
    ramp = torch.polar(torch.ones_like(ang), ang)
    return ramp[None, :] * ramp[:, None]


def eval_basis(basis, coefficients):
    """dlu.eval_basis (utils/math.py:177-196)."""
    if tuple(basis.shape[:coefficients.dim()]) != tuple(coefficients.shape):
        raise ValueError(
            "The leading basis dimensions must match the coefficient shape, "
            f"received {tuple(basis.shape)} and {tuple(coefficients.shape)}.")
    return ops.BasisEvalFunction.apply(coefficients, basis, None)
