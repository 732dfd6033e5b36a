"""Mirror of ``dLux.psfs.PSF`` (/root/reference/src/dLux/psfs.py:14-266): the PSF array + pixel
scale container that ``propagate(..., return_psf=True)`` / ``*Source.model(..., return_psf=True)``
return and the detector layers act on.  Plain O(M^2) torch arithmetic after the fused kernels."""
from __future__ import annotations

import copy

import numpy as np
import torch

from .utils.array_ops import downsample

__all__ = ["PSF", "convolve_same"]


def convolve_same(image: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """``jax.scipy.signal.convolve(image, kernel, mode="same")`` for 2-d arrays: the full linear
    convolution cropped to the shape of ``image`` around its centre."""
    kh, kw = kernel.shape
    full = torch.nn.functional.conv2d(image[None, None], torch.flip(kernel, (0, 1))[None, None].to(image.dtype),
                                      padding=(kh - 1, kw - 1))[0, 0]
    y0, x0 = (kh - 1) // 2, (kw - 1) // 2
    return full[y0:y0 + image.shape[0], x0:x0 + image.shape[1]]


class PSF:
    """psfs.py:14-111: PSF array + pixel scale."""

    def __init__(self, data, pixel_scale):
        self.data = data if torch.is_tensor(data) else torch.as_tensor(np.asarray(data, dtype=np.float32))
        self.pixel_scale = pixel_scale if torch.is_tensor(pixel_scale) else torch.as_tensor(
            np.asarray(pixel_scale, dtype=np.float32), device=self.data.device)

    def set(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    @property
    def npixels(self):
        return self.data.shape[-1]

    @property
    def ndim(self):
        return self.pixel_scale.dim()

    def downsample(self, n: int):                      # psfs.py:74-91: sum over n x n blocks
        return self.set(data=downsample(self.data, n, mean=False), pixel_scale=self.pixel_scale * n)

    def convolve(self, other):                         # psfs.py:93-110
        other = other if torch.is_tensor(other) else torch.as_tensor(np.asarray(other, np.float32))
        return self.set(data=convolve_same(self.data, other.to(self.data.device, self.data.dtype)))

    def rotate(self, angle, method: str = "linear"):   # psfs.py:112-128
        from .utils import interpolation as _interp
        return self.set(data=_interp.rotate(self.data, angle, method))

    def interpolate(self, transformation, method: str = "linear", fill: float = 0.0):   # psfs.py:130-157
        from .apertures import CoordTransform
        from .utils import geometry as _G, interpolation as _interp
        if not isinstance(transformation, CoordTransform):
            raise TypeError("transformation must be a BaseCoordTransform.")
        knots = _G.pixel_coords(self.npixels, self.npixels * self.pixel_scale.to(self.data.dtype),
                                device=self.data.device, dtype=self.data.dtype)
        return self.set(data=_interp.interp(self.data, knots, transformation(knots), method, fill))

    def resize(self, npixels: int):                    # psfs.py:159-173 -> dlu.resize: centred crop / zero pad
        n_in = self.npixels
        if npixels == n_in:
            return self
        if n_in % 2 != npixels % 2:
            raise ValueError("Center-preserving resizing requires parity consistency, i.e. even -> even "
                             f"or odd -> odd: {n_in} -> {npixels}.")
        if npixels < n_in:
            a, b = (n_in - npixels) // 2, (n_in + npixels) // 2
            return self.set(data=self.data[..., a:b, a:b])
        p = (npixels - n_in) // 2
        return self.set(data=torch.nn.functional.pad(self.data, (p, p, p, p)))

    def flip(self, axis):                              # psfs.py:175-190 (np.flip on the given axes)
        axes = (axis,) if isinstance(axis, int) else tuple(axis)
        return self.set(data=torch.flip(self.data, list(axes)))

    def _op(self, other, fn):
        if other is None:
            return self
        if isinstance(other, PSF):
            other = other.data
        if isinstance(other, np.ndarray):
            other = torch.as_tensor(other, device=self.data.device)
        return self.set(data=fn(self.data, other))

    def __add__(self, other):
        return self._op(other, lambda a, b: a + b)

    def __sub__(self, other):
        return self._op(other, lambda a, b: a - b)

    def __mul__(self, other):
        return self._op(other, lambda a, b: a * b)

    def __truediv__(self, other):
        return self._op(other, lambda a, b: a / b)
