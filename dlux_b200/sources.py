"""Mirror of ``PointSource`` / ``PointSources`` / ``BinarySource`` (/root/reference/src/dLux/
sources.py:316-327, 392-411, 524-635), the image-plane sources built on them (``ResolvedSource``
:414-521, ``PointResolvedSource`` :638-748, ``Scene`` :751-850) and the spectrum normalisation they
rely on (spectra.py:84-117)."""
from __future__ import annotations

import numpy as np
import torch

from .psfs import PSF, convolve_same

__all__ = ["PointSource", "PointSources", "BinarySource", "ResolvedSource", "PointResolvedSource",
           "Scene", "convolve_same"]


def _np32(x):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


class _Source:
    def __init__(self, wavelengths, weights=None):
        self.wavelengths = np.atleast_1d(_np32(wavelengths))
        if weights is None:
            weights = np.ones(self.wavelengths.shape, np.float32) / np.float32(self.wavelengths.shape[-1])
        weights = _np32(weights)
        if weights.ndim == 2:                                            # spectra.py:88-92: per-source rows
            self.weights = (weights / weights.sum(-1)[:, None]).astype(np.float32)
            if self.weights.shape[-1:] != self.wavelengths.shape:
                raise ValueError("wavelengths and weights must have the same trailing shape.")
        else:
            self.weights = (weights / weights.sum()).astype(np.float32)
            if self.weights.shape != self.wavelengths.shape:
                raise ValueError("wavelengths and weights must have the same shape.")

    def normalised_weights(self):                                         # spectra.py:113-117
        if self.weights.ndim == 2:
            return (self.weights / self.weights.sum(-1)[:, None]).astype(np.float32)
        return (self.weights / self.weights.sum()).astype(np.float32)


def _validate_return_mode(return_wf, return_psf):
    if return_wf and return_psf:
        raise ValueError("Cannot return both Wavefront and PSF objects.")


class PointSource(_Source):
    def __init__(self, wavelengths=None, position=None, flux=1.0, weights=None):
        position = np.zeros(2, np.float32) if position is None else position
        self.position = position if torch.is_tensor(position) else _np32(position)
        if tuple(self.position.shape) != (2,):
            raise ValueError("position must be a 1d array of shape (2,).")
        self.flux = flux if torch.is_tensor(flux) else np.float32(flux)
        super().__init__(wavelengths, weights)

    def model(self, optics, return_wf=False, return_psf=False):           # sources.py:316-327
        _validate_return_mode(return_wf, return_psf)
        w = self.normalised_weights()
        if torch.is_tensor(self.flux):
            weights = torch.as_tensor(w, device=self.flux.device) * self.flux
        else:
            weights = w * self.flux
        return optics.propagate(self.wavelengths, self.position, weights, return_wf, return_psf)


class PointSources(_Source):
    def __init__(self, wavelengths=None, position=None, flux=None, weights=None):
        self.position = position if torch.is_tensor(position) else _np32(position)
        if self.position.ndim != 2 or self.position.shape[-1] != 2:
            raise ValueError("position must be a 2d array of shape (nstars, 2).")
        if flux is None:
            flux = np.ones(len(self.position), np.float32)
        self.flux = flux if torch.is_tensor(flux) else _np32(flux)
        if self.flux.ndim != 1:
            raise ValueError("flux must be a 1d array.")
        if len(self.flux) != len(self.position):
            raise ValueError("Length of flux must be equal to length of positions.")
        super().__init__(wavelengths, weights)

    def model(self, optics, return_wf=False, return_psf=False):           # sources.py:392-411
        _validate_return_mode(return_wf, return_psf)
        w = self.normalised_weights()
        if torch.is_tensor(self.flux):
            weights = torch.as_tensor(w, device=self.flux.device)[None, :] * self.flux[:, None]
        else:
            weights = w[None, :] * self.flux[:, None]
        if return_wf:                                                    # :401-407: vectorised Wavefront [S, L, ...]
            wfs = [optics.propagate(self.wavelengths, self.position[s], weights[s], return_wf=True)
                   for s in range(len(self.position))]
            stack = lambda name: torch.stack([getattr(w, name) for w in wfs])
            return wfs[0].set(phasor=stack("phasor"), wavelength=stack("wavelength"),
                              pixel_scale=stack("pixel_scale"), center=stack("center"))
        if getattr(optics, "fused", False) and hasattr(optics, "fused_propagate") and optics._can_fuse():
            out = optics.fused_propagate(self.wavelengths, self.position, weights)
        else:
            out = None
            for s in range(len(self.position)):
                psf = optics.propagate(self.wavelengths, self.position[s], weights[s])
                out = psf if out is None else out + psf
        if return_psf:                                                   # :408-409
            return PSF(out, optics._psf_pixel_scale_out(out.device))
        return out


class BinarySource(_Source):
    """sources.py:524-635: two point sources parametrised by mean position, separation,
    position angle, mean flux and contrast (utils/source.py:10-52); ``weights`` may be [2, L]
    (one spectrum per component).  Any parameter given as a CUDA tensor with requires_grad is
    differentiable: it reaches the fused kernels as source offsets / spectral weights."""

    def __init__(self, wavelengths=None, position=None, mean_flux=1.0, separation=0.0,
                 position_angle=np.pi / 2, contrast=1.0, weights=None):
        wl = np.atleast_1d(_np32(wavelengths))
        if weights is None:
            weights = np.ones((2, len(wl)), np.float32)
        position = np.zeros(2, np.float32) if position is None else position
        keep = lambda v: v if torch.is_tensor(v) else _np32(v)
        self.position = keep(position)
        if tuple(self.position.shape) != (2,):
            raise ValueError("position must be a 1d array of shape (2,).")
        self.mean_flux, self.separation = keep(mean_flux), keep(separation)
        self.position_angle, self.contrast = keep(position_angle), keep(contrast)
        super().__init__(wl, weights)

    def _device(self, optics):
        return getattr(optics, "device", torch.device("cuda"))

    def model(self, optics, return_wf=False, return_psf=False):
        _validate_return_mode(return_wf, return_psf)
        dev = self._device(optics)
        t = lambda v: v.to(dev, torch.float32) if torch.is_tensor(v) else torch.as_tensor(v, device=dev)
        pos, sep, pa = t(self.position), t(self.separation), t(self.position_angle)
        mean_flux, contrast = t(self.mean_flux), t(self.contrast)
        sep_vec = torch.stack([sep / 2 * torch.sin(pa), sep / 2 * torch.cos(pa)])     # utils/source.py:50-52
        positions = torch.stack([pos + sep_vec, pos - sep_vec])
        flux = 2 * torch.stack([contrast * mean_flux, mean_flux]) / (1 + contrast)    # utils/source.py:26
        w = t(self.normalised_weights())
        if w.dim() == 1:
            w = w[None, :].expand(2, -1)
        weights = w * flux[:, None]
        if return_wf:                                                    # :620-628: vectorised Wavefront [2, L, ...]
            wfs = [optics.propagate(self.wavelengths, positions[s], weights[s], return_wf=True) for s in range(2)]
            stack = lambda name: torch.stack([getattr(w, name) for w in wfs])
            return wfs[0].set(phasor=stack("phasor"), wavelength=stack("wavelength"),
                              pixel_scale=stack("pixel_scale"), center=stack("center"))
        if getattr(optics, "fused", False) and hasattr(optics, "fused_propagate") and optics._can_fuse():
            out = optics.fused_propagate(self.wavelengths, positions, weights)
        else:
            out = None
            for s in range(2):
                psf = optics.propagate(self.wavelengths, positions[s], weights[s])
                out = psf if out is None else out + psf
        if return_psf:                                                   # :631-632
            return PSF(out, optics._psf_pixel_scale_out(out.device))
        return out


def _wavefront_not_supported():
    raise NotImplementedError(
        "Wavefront information cannot be preserved through convolution. "
        "Convolution can only operate on PSFs (incoherent light). "
        "Please use return_wf=False to get the PSF array or return_psf=True "
        "to get a PSF object.")


class ResolvedSource(PointSource):
    """sources.py:414-521: a point-source PSF convolved with a (normalised, floored) intensity
    distribution -- the convolution is image-plane work after the fused PSF."""

    def __init__(self, wavelengths=None, position=None, flux=1.0, distribution=None, weights=None):
        d = np.ones((3, 3), np.float32) if distribution is None else distribution
        d = d if torch.is_tensor(d) else _np32(d)
        if d.ndim != 2:
            raise ValueError("distribution must be a 2d array.")
        self.distribution = d / d.sum()
        super().__init__(wavelengths, position, flux, weights)

    def _distribution(self, device):
        d = self.distribution if torch.is_tensor(self.distribution) else torch.as_tensor(self.distribution, device=device)
        d = torch.clamp(d.to(device, torch.float32), min=0.0)                       # sources.py:470-474
        return d / d.sum()

    def model(self, optics, return_wf=False, return_psf=False):
        _validate_return_mode(return_wf, return_psf)
        if return_wf:
            _wavefront_not_supported()
        psf = PointSource.model(self, optics)
        conv = convolve_same(psf, self._distribution(psf.device))
        return PSF(conv, optics._psf_pixel_scale_out(conv.device)) if return_psf else conv


class PointResolvedSource(ResolvedSource):
    """sources.py:638-748: an unresolved star plus a resolved component sharing its spectrum
    shape, fluxes set by the mean flux and the contrast; ``weights`` may be [2, L]."""

    def __init__(self, wavelengths=None, position=None, flux=1.0, distribution=None, contrast=1.0,
                 weights=None):
        wl = np.atleast_1d(_np32(wavelengths))
        if weights is None:
            weights = np.ones((2, len(wl)), np.float32)
        self.contrast = contrast if torch.is_tensor(contrast) else np.float32(contrast)
        super().__init__(wl, position, flux, distribution, weights)

    def model(self, optics, return_wf=False, return_psf=False):
        _validate_return_mode(return_wf, return_psf)
        if return_wf:
            _wavefront_not_supported()
        dev = getattr(optics, "device", torch.device("cuda"))
        t = lambda v: v.to(dev, torch.float32) if torch.is_tensor(v) else torch.as_tensor(v, device=dev)
        flux, contrast = t(self.flux), t(self.contrast)
        fluxes = 2 * torch.stack([contrast * flux, flux]) / (1 + contrast)          # utils/source.py:26
        w = t(self.normalised_weights())
        if w.dim() == 1:
            w = w[None, :].expand(2, -1)
        weights = w * fluxes[:, None]
        # sources.py:728-743: the optics are propagated with their DEFAULT spectral weights 1/L
        # (the per-component weights cannot ride along), and each wavelength's PSF is then scaled by
        # weights[k]: psf_k = sum_l weights[k, l] * (1/L) |E_l|^2.  The 1/L factor is the reference's
        # behaviour and is kept (pinned by tests/golden/reference_classes.npz: sm_point_resolved_psf).
        weights = weights / np.float32(len(self.wavelengths))
        point = optics.propagate(self.wavelengths, self.position, weights[0])
        resolved = optics.propagate(self.wavelengths, self.position, weights[1])
        psf = point + convolve_same(resolved, self._distribution(point.device))
        return PSF(psf, optics._psf_pixel_scale_out(psf.device)) if return_psf else psf


class Scene:
    """sources.py:751-850: several sources modelled through the same optics and summed."""

    def __init__(self, sources):
        if isinstance(sources, dict):
            self.sources = dict(sources)
        else:
            self.sources = {}
            for i, s in enumerate(sources):
                key, src = s if isinstance(s, tuple) else (f"{type(s).__name__}_{i}", s)
                self.sources[key] = src

    def __getattr__(self, key):
        srcs = self.__dict__.get("sources", {})
        if key in srcs:
            return srcs[key]
        raise AttributeError(key)

    def model(self, optics, return_wf=False, return_psf=False):
        _validate_return_mode(return_wf, return_psf)
        if return_wf:                                                    # :827-835: one Wavefront per source
            return {k: src.model(optics, return_wf=True) for k, src in self.sources.items()}
        out = None
        for src in self.sources.values():
            psf = src.model(optics)
            out = psf if out is None else out + psf
        if return_psf:                                                   # :838-847
            return PSF(out, optics._psf_pixel_scale_out(out.device))
        return out
