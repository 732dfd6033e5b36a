"""Autograd twin of the oracle (TEST INFRASTRUCTURE ONLY -- see mft_oracle.py).

Gradients of a scalar loss through the polychromatic PSF, on torch-CPU complex64
(or complex128), standing in for ``jax.grad`` through
``OpticalSystem.propagate`` (/root/reference/docs/phase_retrieval.md:269-287).
The DFT matrices and every geometry scalar come from the NumPy oracle, so the
forward value is the oracle's; autograd differentiates w.r.t. the pupil-plane
leaves (coefficients / opd / transmission) and the spectral weights.
"""
from __future__ import annotations

import numpy as np
import torch

from . import mft_oracle as O


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def poly_psf(transmission, opd, wavelengths, weights, *, diameter, psf_npixels,
             pixel_scale_rad, offset=(0.0, 0.0), basis=None, coefficients=None,
             normalise=True, dtype=np.float32):
    """PSF = sum_l w_l |MFT_l(T * exp(i k_l (opd + basis.c + tilt)))|^2.
    ``transmission, opd, coefficients, weights`` may be torch tensors requiring grad."""
    rdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    F = np.dtype(dtype).type
    T = transmission if torch.is_tensor(transmission) else _t(transmission, rdt)
    N = T.shape[-1]
    total_opd = torch.zeros((N, N), dtype=rdt)
    if opd is not None:
        total_opd = total_opd + (opd if torch.is_tensor(opd) else _t(opd, rdt))
    if basis is not None:
        c = coefficients if torch.is_tensor(coefficients) else _t(coefficients, rdt)
        total_opd = total_opd + torch.tensordot(c, _t(basis, rdt), dims=1)
    wf0 = O.OracleWavefront(1.0, N, diameter, dtype)
    xs = wf0.xs
    X, Y = np.meshgrid(xs, xs)
    if torch.is_tensor(offset):      # differentiable source position
        tilt = offset[0] * _t(X, rdt) + offset[1] * _t(Y, rdt)
    else:
        tilt = _t(F(offset[0]) * X + F(offset[1]) * Y, rdt)
    w = weights if torch.is_tensor(weights) else _t(weights, rdt)
    psf = 0.0
    for l, wl in enumerate(np.asarray(wavelengths, dtype=dtype)):
        k = float(F(F(2 * np.pi) / F(wl)))
        ph = torch.polar(torch.ones_like(total_opd), k * tilt) * (1.0 / N ** 2)
        ph = ph * T
        ph = ph * torch.polar(torch.ones_like(total_opd), k * total_opd)
        if normalise:
            ph = ph * torch.rsqrt((ph.abs() ** 2).sum())
        ps_in = wf0.pixel_scale
        ax = _t(O.transfer_matrix(wl, N, ps_in, psf_npixels, pixel_scale_rad, 0.0, dtype=dtype), cdt)
        nf = O.calc_nfringes(wl, N, ps_in, psf_npixels, pixel_scale_rad, dtype=dtype)
        nrm = float(O.mft_norm(nf, N, psf_npixels, dtype))
        E = (ax.T @ ph.to(cdt)) @ ax * nrm
        psf = psf + w[l] * (E.real ** 2 + E.imag ** 2)
    return psf
