"""Image-plane tail after the MFT path (SURVEY 8f NEXT-4): the detector
layers that act on it and ``Telescope``, which chains optics -> source -> detector.  Mirrors
/root/reference/src/dLux/psfs.py:14-111, layers/detector_layers.py:100-296, detectors.py:47-128
and instruments.py:37-172.  Everything here is O(M^2) torch arithmetic on the oversampled PSF the
fused kernels return (differentiable, device-agnostic); it is kept out of the CUDA library on
purpose -- at C3 it touches 1 MB per step against the GB-scale operand traffic of the MFT."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .psfs import PSF
from .sources import Scene, _Source
from .utils.array_ops import downsample

__all__ = ["ApplyInterpolation", "PSF", "DetectorLayer", "ApplyPixelResponse", "ApplyJitter", "ApplySaturation", "AddConstant",
           "Downsample", "LayeredDetector", "Telescope", "gaussian_kernel"]


def gaussian_kernel(sigma, npixels: int, extent: float = 5.0, device=None):
    """``dlu.gaussian(mean=0, std=(sigma, sigma), npixels)`` (utils/math.py:20-68): separable normal
    pdf sampled on linspace(-extent, extent, npixels), normalised to unit sum."""
    s = sigma if torch.is_tensor(sigma) else torch.as_tensor(float(sigma), dtype=torch.float32, device=device)
    x = torch.linspace(-extent, extent, npixels, dtype=s.dtype, device=s.device)
    g = torch.exp(-0.5 * (x / s) ** 2) / (s * np.sqrt(2 * np.pi))
    k = g[None, :] * g[:, None]
    return k / k.sum()


class DetectorLayer:
    """layers/detector_layers.py:23-65: the user-extension API of the detector."""

    def apply(self, psf: PSF) -> PSF:
        return self(psf)

    def __call__(self, psf: PSF) -> PSF:  # pragma: no cover - abstract
        raise NotImplementedError


class ApplyInterpolation(DetectorLayer):
    """detector_layers.py:68-97: re-samples the PSF on transformed pixel coordinates (``PSF.interpolate``; bilinear
    only, see utils/interpolation.py)."""

    def __init__(self, transformation, method: str = "linear", fill: float = 0.0):
        from .apertures import CoordTransform
        if not isinstance(transformation, CoordTransform):
            raise TypeError("transformation must be a BaseCoordTransform.")
        self.transformation = transformation
        self.method = str(method)
        self.fill = float(fill)

    def __call__(self, psf):
        return psf.interpolate(self.transformation, method=self.method, fill=self.fill)


class ApplyPixelResponse(DetectorLayer):
    def __init__(self, pixel_response):
        self.pixel_response = pixel_response if torch.is_tensor(pixel_response) else np.asarray(pixel_response, np.float32)
        if self.pixel_response.ndim != 2:
            raise ValueError("pixel_response must be a 2d array.")

    def __call__(self, psf):                           # detector_layers.py:130-131
        return psf * self.pixel_response


class ApplyJitter(DetectorLayer):
    """Gaussian jitter: the kernel is a (kernel_size * oversample)^2 normal pdf summed down to
    kernel_size^2 (detector_layers.py:134-199); ``sigma`` in the units of the +-5 extent."""

    def __init__(self, sigma, kernel_size: int = 9, oversample: int = 3):
        self.kernel_size, self.oversample = int(kernel_size), int(oversample)
        self.sigma = sigma if torch.is_tensor(sigma) else np.float32(sigma)
        if self.kernel_size <= 0:
            raise ValueError("kernel_size must be greater than 0.")

    def kernel(self, device=None):
        k = gaussian_kernel(self.sigma, self.kernel_size * self.oversample, device=device)
        return downsample(k, self.oversample, mean=False)

    def __call__(self, psf):
        return psf.convolve(self.kernel(psf.data.device))


class ApplySaturation(DetectorLayer):
    def __init__(self, threshold):
        self.threshold = threshold if torch.is_tensor(threshold) else np.float32(threshold)

    def __call__(self, psf):                           # detector_layers.py:228-229
        return psf.set(data=torch.clamp(psf.data, max=float(self.threshold) if not torch.is_tensor(self.threshold)
                                        else self.threshold))


class AddConstant(DetectorLayer):
    def __init__(self, value):
        self.value = value if torch.is_tensor(value) else np.float32(value)

    def __call__(self, psf):                           # detector_layers.py:258-259
        return psf + self.value


class Downsample(DetectorLayer):
    def __init__(self, kernel_size: int):
        self.kernel_size = int(kernel_size)
        if self.kernel_size <= 0:
            raise ValueError("kernel_size must be greater than 0.")

    def __call__(self, psf):                           # detector_layers.py:292-293
        return psf.downsample(self.kernel_size)


class LayeredDetector:
    """detectors.py:47-128: detector layers applied in order."""

    def __init__(self, layers):
        if isinstance(layers, (list, tuple)):
            od = OrderedDict()
            for i, l in enumerate(layers):
                key, layer = l if isinstance(l, tuple) else (f"{type(l).__name__}_{i}", l)
                od[key] = layer
            layers = od
        for layer in layers.values():
            if not isinstance(layer, DetectorLayer):
                raise TypeError("layers must be DetectorLayer instances")
        self.layers = OrderedDict(layers)

    def __getattr__(self, key):
        layers = self.__dict__.get("layers", {})
        if key in layers:
            return layers[key]
        raise AttributeError(key)

    def __call__(self, psf: PSF, return_psf: bool = False):
        for layer in self.layers.values():
            psf = layer(psf)
        return psf if return_psf else psf.data

    model = __call__


class Telescope:
    """instruments.py:37-172: optics + source(s) + optional detector."""

    def __init__(self, optics, source, detector=None):
        if not hasattr(optics, "propagate"):
            raise TypeError(f"optics must be an OpticalSystem instance, got {type(optics).__name__}.")
        self.optics = optics
        if isinstance(source, (_Source, Scene)):
            self.source = source
        elif isinstance(source, tuple):
            if len(source) != 2 or not isinstance(source[1], (_Source, Scene)):
                raise TypeError("source tuple must be of the form (key: str, Source: Source)")
            self.source = source[1]
        else:
            self.source = Scene(source)
        if detector is not None and not isinstance(detector, LayeredDetector):
            raise TypeError(f"detector must be a Detector instance, got {type(detector).__name__}.")
        self.detector = detector

    def model(self, return_psf: bool = False):
        data = self.source.model(self.optics)
        npix, ps, _ = self.optics._focal_args() if hasattr(self.optics, "_focal_args") else (None, 1.0, None)
        psf = PSF(data, ps if torch.is_tensor(ps) else torch.as_tensor(np.float32(ps), device=data.device))
        if self.detector is not None:
            return self.detector.model(psf, return_psf=return_psf)
        return psf if return_psf else psf.data
