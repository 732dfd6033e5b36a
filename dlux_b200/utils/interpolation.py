"""Interpolating re-sampling of images and fields: ``interp`` / ``scale`` / ``rotate`` of
/root/reference/src/dLux/utils/interpolation.py:13-107, the producers behind ``Wavefront.scale_to / rotate /
interpolate`` (wavefronts.py:442-566), ``PSF.rotate / interpolate`` (psfs.py:112-157) and the ``Rotate`` layer.

The reference delegates the arithmetic to ``interpax.interp2d`` (a third-party package that is neither vendored under
/root/reference nor installable here), so this file restates the PUBLISHED algorithm of its default method: bilinear
interpolation on the rectilinear knot grid, a constant ``fill`` outside it.  It is pinned to SciPy's
``RegularGridInterpolator(method="linear", bounds_error=False, fill_value=fill)`` (tests/test_host.py), not to the
reference dependency itself -- parity with interpax is therefore UNPINNED, and the cubic / spline methods, whose
definitions are interpax's own, are refused rather than guessed.  Plain differentiable torch (values and sampling
coordinates), off the hot path: O(n^2) per image."""
from __future__ import annotations

import numpy as np
import torch

from . import geometry as G

__all__ = ["interp", "scale", "rotate"]


def _check_method(method: str):
    if method != "linear":
        raise NotImplementedError(f"dlux_b200: interpolation method '{method}' is defined by interpax, which is not "
                                  "available to pin against; only 'linear' (bilinear) is implemented")


def interp(image, knot_coords, sample_coords, method: str = "linear", fill: float = 0.0):
    """``image`` [ny, nx] known at ``knot_coords`` [2, ny, nx] (x first, x varying along the last axis), evaluated at
    ``sample_coords`` [2, ...]; points outside the knot grid get ``fill`` (interpolation.py:13-46)."""
    _check_method(method)
    xs, ys = knot_coords[0][0, :].contiguous(), knot_coords[1][:, 0].contiguous()
    xq, yq = sample_coords[0], sample_coords[1]
    shape = xq.shape
    xq, yq = xq.reshape(-1), yq.reshape(-1)

    def cell(knots, q):
        i = torch.searchsorted(knots, q.detach().contiguous(), right=True) - 1
        i = i.clamp(0, knots.numel() - 2)
        t = (q - knots[i]) / (knots[i + 1] - knots[i])
        return i, t

    ix, tx = cell(xs, xq)
    iy, ty = cell(ys, yq)
    v = (image[iy, ix] * (1 - tx) * (1 - ty) + image[iy, ix + 1] * tx * (1 - ty) +
         image[iy + 1, ix] * (1 - tx) * ty + image[iy + 1, ix + 1] * tx * ty)
    outside = (xq < xs[0]) | (xq > xs[-1]) | (yq < ys[0]) | (yq > ys[-1])
    v = torch.where(outside, torch.as_tensor(fill, dtype=v.dtype, device=v.device), v)
    return v.reshape(shape)


def scale(array, npixels: int, ratio, method: str = "linear"):
    """Paraxial re-sampling of a square array onto ``npixels`` pixels ``ratio`` times the input pixel size
    (interpolation.py:49-79): both grids are centred, unit-diameter pixel coordinates."""
    n_in = array.shape[-1]
    coords_in = G.pixel_coords(n_in, 1.0, device=array.device, dtype=array.dtype)
    r = ratio if torch.is_tensor(ratio) else torch.as_tensor(np.asarray(ratio, np.float64), dtype=array.dtype,
                                                             device=array.device)
    coords_out = G.pixel_coords(int(npixels), 1.0, device=array.device, dtype=array.dtype) * (r * npixels / n_in)
    return interp(array, coords_in, coords_out, method)


def rotate(array, angle, method: str = "linear"):
    """Rotation of a square array about its centre by ``angle`` radians (interpolation.py:82-107)."""
    n = array.shape[0]
    coords_in = G.pixel_coords(n, float(n), device=array.device, dtype=array.dtype)   # unit pixels, centred
    return interp(array, coords_in, G.rotate_coords(coords_in, angle), method)
