"""Mirror of ``PointSource`` / ``PointSources`` (/root/reference/src/dLux/sources.py:
316-327, 392-411) and the spectrum normalisation they rely on (spectra.py:84-117)."""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["PointSource", "PointSources"]


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
        self.weights = (weights / weights.sum()).astype(np.float32)      # spectra.py:88-92
        if self.weights.shape != self.wavelengths.shape:
            raise ValueError("wavelengths and weights must have the same shape.")

    def normalised_weights(self):                                         # spectra.py:113-117
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
        if getattr(optics, "fused", False) and not return_wf and hasattr(optics, "fused_propagate") \
                and optics._fusable() is not None:
            return optics.fused_propagate(self.wavelengths, self.position, weights)
        out = None
        for s in range(len(self.position)):
            psf = optics.propagate(self.wavelengths, self.position[s], weights[s])
            out = psf if out is None else out + psf
        return out
