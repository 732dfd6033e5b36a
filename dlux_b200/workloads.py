"""Synthetic inputs for the BASELINE.json configurations: arrays are
built directly with NumPy (seeded), no dLux import.  Used by bench.py, the tests and
``__graft_entry__.smoke()``."""
from __future__ import annotations

import numpy as np

__all__ = ["hex_nrm_pupil", "config", "CONFIGS"]


def _coords(n: int, diameter: float):
    x = (np.arange(n) - (n - 1) / 2) * (diameter / n)
    return np.meshgrid(x, x)


def hex_nrm_pupil(n: int, diameter: float, hole_flat: float, centres: np.ndarray):
    """Transmission of a hexagonal-hole non-redundant mask and a per-hole
    piston/tip/tilt OPD basis (3 modes per hole), shaped like what
    ``dlu.sparse_aperture(shape='hex')`` (/root/reference/src/dLux/utils/apertures.py:
    376-469) feeds a ``BasisOptic``."""
    X, Y = _coords(n, diameter)
    T = np.zeros((n, n), np.float32)
    basis = []
    r_in = hole_flat / 2
    for cx, cy in centres:
        x, y = X - cx, Y - cy
        inside = np.ones((n, n), bool)
        for k in range(3):                       # three pairs of flat sides
            a = np.pi / 3 * k
            inside &= np.abs(x * np.cos(a) + y * np.sin(a)) <= r_in
        T += inside
        m = inside.astype(np.float32)
        basis += [m, m * (x / r_in).astype(np.float32), m * (y / r_in).astype(np.float32)]
    return np.clip(T, 0, 1).astype(np.float32), np.stack(basis).astype(np.float32)


# 7-hole pattern (metres) inside a 6.6 m aperture; fixed table
_NRM7 = np.array([[0.0, 2.64], [-2.29, 1.32], [2.29, 0.0], [-1.14, -0.66], [1.14, -1.98],
                  [-2.29, -1.32], [0.0, -2.64]])


def _circ(n, diameter):
    X, Y = _coords(n, diameter)
    return (np.hypot(X, Y) <= diameter / 2).astype(np.float32)


def _zernike_opd_basis(n, diameter, nz, T):
    """nz Zernike modes from Noll 4 (defocus) upwards on the aperture, unit rms, as in the
    reference tutorials (``dlu.zernike_basis(np.arange(4, 4 + nz), coords, diameter)``)."""
    from .utils.zernikes import zernike_basis
    X, Y = _coords(n, diameter)
    return zernike_basis(range(4, 4 + nz), np.stack([X, Y]), diameter) * T


def config(name: str):
    """Returns a dict: wf_npixels, diameter, psf_npixels, psf_pixel_scale (arcsec),
    oversample, transmission, basis (metres per unit coefficient), coefficients,
    wavelengths, weights, positions [S,2] rad, fluxes [S], G (psf cotangent)."""
    if name == "c1":      # 256 px circular aperture + 10-term Zernike OPD, 1 source, 1 wavelength, -> 128
        rng = np.random.default_rng(0)
        n, d = 256, 1.0
        T = _circ(n, d)
        basis = _zernike_opd_basis(n, d, 10, T) * np.float32(1e-9)
        out = dict(wf_npixels=n, diameter=d, psf_npixels=128, psf_pixel_scale=0.05, oversample=1,
                   wavelengths=np.array([1.0e-6], np.float32),
                   coefficients=(20 * rng.standard_normal(10)).astype(np.float32))
    elif name == "c2":    # 512 px pupil, 32 wavelengths, -> 256, grad w.r.t. coefficients
        rng = np.random.default_rng(1)
        n, d = 512, 1.0
        T = _circ(n, d)
        basis = _zernike_opd_basis(n, d, 10, T) * np.float32(1e-9)
        out = dict(wf_npixels=n, diameter=d, psf_npixels=256, psf_pixel_scale=0.025, oversample=1,
                   wavelengths=np.linspace(0.9e-6, 1.1e-6, 32).astype(np.float32),
                   coefficients=(20 * rng.standard_normal(10)).astype(np.float32))
    elif name == "c3":    # JWST-AMI-like hex NRM: 1024 px, 64 wavelengths, oversampled -> 512
        rng = np.random.default_rng(2)
        n, d = 1024, 6.6
        T, basis = hex_nrm_pupil(n, d, 0.8, _NRM7)
        basis = basis * np.float32(30e-9)
        out = dict(wf_npixels=n, diameter=d, psf_npixels=128, psf_pixel_scale=0.0656, oversample=4,
                   wavelengths=np.linspace(4.1e-6, 4.5e-6, 64).astype(np.float32),
                   coefficients=rng.standard_normal(21).astype(np.float32))
    elif name == "c4":    # Toliman-like diffractive pupil: 2048 px, binary 0/pi phase mask, 1000 stars x 64 wavelengths
        rng = np.random.default_rng(3)
        n, d = 2048, 0.125
        T = _circ(n, d)
        f = np.fft.fft2(rng.standard_normal((n, n)))
        f[40:-40, :] = 0
        f[:, 40:-40] = 0
        phase = (np.pi * (np.fft.ifft2(f).real > 0)).astype(np.float32)
        M, ps = 256, 0.7                                   # arcsec / pixel: ~ Nyquist / 1.5 at 585 nm
        fov = M * ps * np.pi / 648000.0
        out = dict(wf_npixels=n, diameter=d, psf_npixels=M, psf_pixel_scale=ps, oversample=1,
                   wavelengths=np.linspace(530e-9, 640e-9, 64).astype(np.float32), phase=phase,
                   coefficients=np.zeros(1, np.float32), n_stars=1000)
        basis = np.zeros((1, n, n), np.float32)
        stars_pos = (rng.uniform(-0.35, 0.35, (1000, 2)) * fov).astype(np.float32)
        stars_flux = (10 ** rng.uniform(0, 3, 1000)).astype(np.float32)
    elif name == "c5":    # Fisher / mask-design sweep: 1024 -> 256, 32 wavelengths, 4096 perturbed coefficient vectors
        rng = np.random.default_rng(4)
        n, d = 1024, 1.0
        T = _circ(n, d)
        basis = _zernike_opd_basis(n, d, 10, T) * np.float32(1e-9)
        fid = (20 * rng.standard_normal(10)).astype(np.float32)
        out = dict(wf_npixels=n, diameter=d, psf_npixels=256, psf_pixel_scale=0.05, oversample=1,
                   wavelengths=np.linspace(0.9e-6, 1.1e-6, 32).astype(np.float32), coefficients=fid,
                   perturbations=(fid[None, :] + 5.0 * rng.standard_normal((4096, 10))).astype(np.float32))
    elif name == "tiny":  # smoke-test size
        rng = np.random.default_rng(9)
        n, d = 128, 1.0
        T = _circ(n, d)
        basis = _zernike_opd_basis(n, d, 4, T) * np.float32(1e-9)
        out = dict(wf_npixels=n, diameter=d, psf_npixels=64, psf_pixel_scale=0.05, oversample=1,
                   wavelengths=np.linspace(0.95e-6, 1.05e-6, 3).astype(np.float32),
                   coefficients=(20 * rng.standard_normal(4)).astype(np.float32))
    else:
        raise KeyError(name)
    L = len(out["wavelengths"])
    M = out["psf_npixels"] * out["oversample"]
    out.update(transmission=T, basis=basis.astype(np.float32), normalise=True,
               weights=np.full(L, 1.0 / L, np.float32),
               positions=np.zeros((1, 2), np.float32), fluxes=np.ones(1, np.float32),
               G=rng.standard_normal((M, M)).astype(np.float32), name=name)
    if name == "c4":
        out.update(positions=stars_pos, fluxes=stars_flux)
    return out


CONFIGS = ("c1", "c2", "c3", "c4", "c5", "tiny")
