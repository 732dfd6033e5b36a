"""Soft-edged aperture shapes and coordinate transforms in torch (SURVEY 8f NEXT-3): the
producers of the transmission array the fused path consumes.  Everything is differentiable
torch arithmetic, so with the transmission cotangent of ``dlux_polypsf_bwd`` the shape
parameters (radius, width, translation, rotation, softening ...) are fitted parameters of
the fused route.  Mirrors /root/reference/src/dLux/utils/geometry.py and the coordinate
transforms of utils/coordinates.py:23-101 (behaviour pinned by tests/golden/
reference_geometry.npz, produced by executing those files)."""
from __future__ import annotations

import math

import numpy as np
import torch

__all__ = ["pixel_coords", "translate_coords", "compress_coords", "shear_coords", "rotate_coords",
           "cart2polar", "soften", "combine", "circle", "square", "rectangle", "reg_polygon", "spider",
           "soft_circle", "soft_square", "soft_rectangle", "soft_reg_polygon", "soft_spider"]


def _t(x, like):
    return x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=like.dtype,
                                                        device=like.device)


# ------------------------------------------------------------------ coordinates
def pixel_coords(npixels: int, diameter, device=None, dtype=torch.float32):
    """utils/coordinates.py:208-266 (default, symmetric pixel-centre grid): [2, n, n], x along
    the last axis."""
    d = diameter if torch.is_tensor(diameter) else torch.as_tensor(float(diameter), dtype=dtype, device=device)
    idx = torch.arange(npixels, dtype=d.dtype, device=d.device) - (npixels - 1) / 2
    xs = idx * (d / npixels)
    X, Y = torch.meshgrid(xs, xs, indexing="xy")
    return torch.stack([X, Y])


def translate_coords(coords, translation):             # coordinates.py:23-39
    return coords - _t(translation, coords)[:, None, None]


def compress_coords(coords, compress):                 # coordinates.py:42-58
    return coords * _t(compress, coords)[:, None, None]


def shear_coords(coords, shear):                       # coordinates.py:61-78
    return coords + coords.transpose(-1, -2) * _t(shear, coords)[:, None, None]


def rotate_coords(coords, rotation):                   # coordinates.py:81-101
    r = _t(rotation, coords)
    x, y = coords[0], coords[1]
    c, s = torch.cos(-r), torch.sin(-r)
    return torch.stack([c * x + s * y, -s * x + c * y])


def cart2polar(coords):                                # coordinates.py:166-184
    return torch.stack([torch.hypot(coords[0], coords[1]), torch.atan2(coords[1], coords[0])])


# ------------------------------------------------------------------ signed distances (> 0 outside)
def _circ_distance(coords, radius):
    return torch.hypot(coords[0], coords[1]) - _t(radius, coords)


def _square_distance(coords, width):
    return coords.abs().amax(0) - _t(width, coords) / 2


def _rectangle_distance(coords, width, height):
    return torch.maximum(coords[0].abs() - _t(width, coords) / 2, coords[1].abs() - _t(height, coords) / 2)


def _spider_distance(coords, width, angle_deg):
    rc = rotate_coords(coords, _t(angle_deg, coords) * (math.pi / 180.0))
    return torch.maximum(rc[0].abs() - _t(width, coords) / 2, rc[1])


def _reg_polygon_distance(coords, nsides: int, radius):
    """Largest signed distance to the n edge lines of the regular polygon whose vertices sit at
    angles -pi + 2 pi i / n on the circle of `radius` (geometry.py:529-585 builds the same lines in
    slope / intercept form); edge i has outward normal at -pi + (2 i + 1) pi / n and lies at the
    apothem radius * cos(pi / n)."""
    r = _t(radius, coords)
    i = torch.arange(nsides, dtype=coords.dtype, device=coords.device)
    ang = -math.pi + (2 * i + 1) * (math.pi / nsides)
    nx, ny = torch.cos(ang)[:, None, None], torch.sin(ang)[:, None, None]
    return (nx * coords[0] + ny * coords[1]).amax(0) - r * math.cos(math.pi / nsides)


# ------------------------------------------------------------------ edges
def soften(distances, clip_dist, invert: bool = False):
    """geometry.py:66-92: clip the (inside-positive) distances to +-clip_dist and rescale to
    [0, 1]; a constant array becomes its hard support."""
    if invert:
        distances = -distances
    c = _t(clip_dist, distances)
    d = torch.maximum(torch.minimum(distances, c), -c)
    lo, hi = d.min(), d.max()
    if bool(hi == lo):
        return (d > 0).to(d.dtype)
    return (d - lo) / (hi - lo)


def combine(arrays, oversample: int = 1, use_sum: bool = False):   # geometry.py:20-46
    from .array_ops import downsample
    a = torch.stack(list(arrays)) if not torch.is_tensor(arrays) else arrays
    out = a.sum(0) if use_sum else a.prod(0)
    return out if oversample == 1 else downsample(out, oversample)


def _hard(dist, invert):
    return (dist > 0).to(dist.dtype) if invert else (dist < 0).to(dist.dtype)


def circle(coords, radius, invert: bool = False):
    return _hard(_circ_distance(coords, radius), invert)


def square(coords, width, invert: bool = False):
    return _hard(_square_distance(coords, width), invert)


def rectangle(coords, width, height, invert: bool = False):
    return _hard(_rectangle_distance(coords, width, height), invert)


def reg_polygon(coords, rmax, nsides: int, invert: bool = False):
    return _hard(_reg_polygon_distance(coords, nsides, rmax), invert)


def spider(coords, width, angles):                     # geometry.py:195-222: 0 under any arm
    angles = _t(angles, coords).reshape(-1)
    under = torch.stack([_spider_distance(coords, width, a) < 0 for a in angles]).any(0)
    return (~under).to(coords.dtype)


def soft_circle(coords, radius, clip_dist=0.1, invert: bool = False):
    return soften(-_circ_distance(coords, radius), clip_dist, invert)


def soft_square(coords, width, clip_dist=0.1, invert: bool = False):
    return soften(-_square_distance(coords, width), clip_dist, invert)


def soft_rectangle(coords, width, height, clip_dist=0.1, invert: bool = False):
    return soften(-_rectangle_distance(coords, width, height), clip_dist, invert)


def soft_reg_polygon(coords, radius, nsides: int, clip_dist=0.1, invert: bool = False):
    return soften(-_reg_polygon_distance(coords, nsides, radius), clip_dist, invert)


def soft_spider(coords, width, angles, clip_dist=0.1, invert: bool = False):
    angles = _t(angles, coords).reshape(-1)
    d = torch.stack([_spider_distance(coords, width, a) for a in angles]).amin(0)
    return soften(-d, clip_dist, invert)
