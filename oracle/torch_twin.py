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


def poly_psf_full(transmission, opd, wavelengths, weights, *, diameter, psf_npixels, pixel_scale_rad,
                  offset=(0.0, 0.0), basis=None, coefficients=None, phase=None, normalise=True,
                  focal_length=None):
    """Float64 twin in which EVERY operand is a torch expression, geometry included
    (propagation.py:110-127, 165-175, 246-254 restated in torch): wavelengths, pixel_scale_rad,
    diameter, focal_length, offset, weights, transmission, opd, phase, coefficients may all
    require grad.  Equal to ``poly_psf(..., dtype=float64)`` to rounding; used where the tests
    need exact float64 derivatives w.r.t. the pixel scale or the wavelengths."""
    rdt, cdt = torch.float64, torch.complex128
    tt = lambda v: v.to(rdt) if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float64))
    T = tt(transmission)
    N = T.shape[-1]
    M = int(psf_npixels)
    total_opd = torch.zeros((N, N), dtype=rdt)
    if opd is not None:
        total_opd = total_opd + tt(opd)
    if basis is not None:
        total_opd = total_opd + torch.tensordot(tt(coefficients), tt(basis), dims=1)
    diameter, ps_out, off, w, wls = tt(diameter), tt(pixel_scale_rad), tt(offset), tt(weights), tt(wavelengths)
    ps_in = diameter / N
    idx = torch.arange(N, dtype=rdt) - (N - 1) / 2
    xs = idx * ps_in                                               # coordinates.py:129
    X, Y = torch.meshgrid(xs, xs, indexing="xy")                   # wavefronts.py:605-606
    tilt = off[0] * X + off[1] * Y
    x_in = idx / N                                                 # propagation.py:113-114
    a_out = torch.arange(M, dtype=rdt) - (M - 1) / 2
    psf = 0.0
    for l in range(wls.shape[0]):
        wl = wls[l]
        k = 2 * np.pi / wl
        ph = torch.polar(torch.full_like(total_opd, 1.0 / N ** 2), k * tilt) * T
        ph = ph * torch.polar(torch.ones_like(total_opd), k * total_opd)
        if phase is not None:
            ph = ph * torch.polar(torch.ones_like(total_opd), tt(phase))
        if normalise:
            ph = ph * torch.rsqrt((ph.abs() ** 2).sum())
        fringe = wl / (ps_in * N)
        s = ps_out / fringe
        nf = M * ps_out / fringe
        if focal_length is not None:
            s, nf = s / tt(focal_length), nf / tt(focal_length)
        ax = torch.polar(torch.ones((N, M), dtype=rdt), -2 * np.pi * torch.outer(x_in, a_out * s))
        E = (ax.T @ ph.to(cdt)) @ ax * (nf / (N * M))
        psf = psf + w[l] * (E.real ** 2 + E.imag ** 2)
    return psf
