"""Dynamic (parametrised, soft-edged) aperture layers -- SURVEY 8f NEXT-3.  Mirrors the shape
layers of /root/reference/src/dLux/layers/apertures.py (CircularAperture :240-311,
SquareAperture :314-387, RectangularAperture :390-472, RegPolyAperture :475-555, Spider
:558-640, CompoundAperture :1005-1069, MultiAperture :1072-1117) and ``CoordTransform``
(coordinates.py:230-316).  They PRODUCE the transmission array the MFT path consumes: the
arithmetic is differentiable torch (dlux_b200/utils/geometry.py), and in the fused route the
transmission cotangent of ``dlux_polypsf_bwd`` flows back through it, so radii, widths,
translations, rotations ... given as CUDA tensors with requires_grad are fitted parameters.
``AberratedAperture`` (:643-800) adds a Zernike OPD / phase / amplitude basis evaluated on the aperture's own
(transformed, normalised) coordinates; circular apertures only (the polygon "polike" bases are not mirrored)."""
from __future__ import annotations

from collections import OrderedDict

import math

import numpy as np
import torch

from .layers import OpticalLayer
from .utils import geometry as G

__all__ = ["CoordTransform", "CircularAperture", "SquareAperture", "RectangularAperture",
           "RegPolyAperture", "Spider", "CompoundAperture", "MultiAperture", "AberratedAperture"]


def _param(v, shape, name):
    if v is None:
        return None
    if not torch.is_tensor(v):
        v = np.asarray(v, dtype=np.float32)
    if tuple(v.shape) != shape:
        raise ValueError(f"{name} must have shape {shape}.")
    return v


class CoordTransform:
    """coordinates.py:230-316: translation, shear, compression, rotation, applied in that order."""

    def __init__(self, translation=None, rotation=None, compression=None, shear=None):
        self.translation = _param(translation, (2,), "translation")
        self.rotation = _param(rotation, (), "rotation")
        self.compression = _param(compression, (2,), "compression")
        self.shear = _param(shear, (2,), "shear")

    def __call__(self, coords):
        if self.translation is not None:
            coords = G.translate_coords(coords, self.translation)
        if self.shear is not None:
            coords = G.shear_coords(coords, self.shear)
        if self.compression is not None:
            coords = G.compress_coords(coords, self.compression)
        if self.rotation is not None:
            coords = G.rotate_coords(coords, self.rotation)
        return coords

    apply = __call__


class _DynamicAperture(OpticalLayer):
    def __init__(self, transformation=None, occulting: bool = False, softening=1.0, normalise: bool = False):
        if transformation is not None and not isinstance(transformation, CoordTransform):
            raise TypeError("transformation must be a BaseCoordTransform instance, "
                            f"got {type(transformation).__name__}.")
        self.transformation = transformation
        self.occulting = bool(occulting)
        self.softness = softening if torch.is_tensor(softening) else np.float32(softening)
        if float(self.softness) <= 0:
            raise ValueError("softening must be greater than 0.")
        self.normalise = bool(normalise)

    def _shape(self, coords, clip):  # pragma: no cover - abstract
        raise NotImplementedError

    def transmission(self, coords, pixel_scale):
        """apertures.py:299-303 and siblings: edges softened over `softening` pixels."""
        if self.transformation is not None:
            coords = self.transformation(coords)
        return self._shape(coords, pixel_scale * self.softness / 2)

    def __call__(self, wavefront):                     # apertures.py:134-152
        wavefront = wavefront * self.transmission(wavefront.coordinates(), wavefront.pixel_scale)
        return wavefront.normalise() if self.normalise else wavefront


class CircularAperture(_DynamicAperture):
    def __init__(self, radius, transformation=None, occulting=False, softening=1.0, normalise=False):
        super().__init__(transformation, occulting, softening, normalise)
        self.radius = _param(radius, (), "radius")

    def _shape(self, coords, clip):
        return G.soft_circle(coords, self.radius, clip, self.occulting)


    @property
    def extent(self):                                   # apertures.py:306-307
        return self.radius

    nsides = 0


class AberratedAperture(OpticalLayer):
    """apertures.py:643-800: a dynamic aperture carrying Zernike aberrations generated on its own coordinates
    (transformed, divided by the aperture extent), so that the basis follows the aperture when its position /
    size are fitted.  ``effect`` in {"opd", "phase", "amplitude"}; coefficients may be a CUDA tensor with
    requires_grad.  Runs on the fused route (transmission and OPD / phase producers) and layer by layer."""

    def __init__(self, aperture, noll_inds, coefficients=None, effect: str = "opd"):
        if isinstance(aperture, (_Composite, Spider)) or not isinstance(aperture, _DynamicAperture):
            raise TypeError("AberratedApertures cannot contain Static, Compound or Multi Apertures (or spiders).")
        if aperture.occulting:
            raise TypeError("AberratedApertures cannot be occulting.")
        if effect not in ("opd", "phase", "amplitude"):
            raise ValueError("effect must be 'opd', 'phase', or 'amplitude'.")
        self.aperture = aperture
        self.effect = effect
        self.noll_inds = [int(j) for j in noll_inds]
        if coefficients is None:
            coefficients = np.zeros(len(self.noll_inds), np.float32)
        self.coefficients = coefficients if torch.is_tensor(coefficients) else np.asarray(coefficients, np.float32)
        if tuple(self.coefficients.shape) != (len(self.noll_inds),):
            raise ValueError("coefficients must have one entry per Noll index")

    @property
    def normalise(self):
        return self.aperture.normalise

    def transmission(self, coords, pixel_scale):        # :732-733
        return self.aperture.transmission(coords, pixel_scale)

    def calc_basis(self, coords):                       # :735-752; polynomials.py:40-51: Zernikes on a circular
        from .utils.zernikes import polike_basis_torch, zernike_basis_torch   # aperture, "polikes" on n sides
        if self.aperture.transformation is not None:
            coords = self.aperture.transformation(coords)
        ext = self.aperture.extent
        ext = ext.to(coords.device, coords.dtype) if torch.is_tensor(ext) else float(ext)
        nsides = int(self.aperture.nsides)
        if nsides == 0:
            return zernike_basis_torch(self.noll_inds, coords / ext)
        return polike_basis_torch(nsides, self.noll_inds, coords / ext)

    def eval_basis(self, coords):                       # :754-771
        c = self.coefficients
        c = c.to(coords.device, coords.dtype) if torch.is_tensor(c) else torch.as_tensor(c, device=coords.device,
                                                                                          dtype=coords.dtype)
        return torch.tensordot(c, self.calc_basis(coords), dims=1)

    def __call__(self, wavefront):                      # :773-800
        wavefront = wavefront * self.transmission(wavefront.coordinates(), wavefront.pixel_scale)
        if self.normalise:
            wavefront = wavefront.normalise()
        ab = self.eval_basis(wavefront.coordinates())
        if self.effect == "phase":
            return wavefront.add_phase(ab)
        if self.effect == "opd":
            return wavefront.add_opd(ab)
        return wavefront * (1 + ab)


class SquareAperture(_DynamicAperture):
    def __init__(self, width, transformation=None, occulting=False, softening=1.0, normalise=False):
        super().__init__(transformation, occulting, softening, normalise)
        self.width = _param(width, (), "width")

    def _shape(self, coords, clip):
        return G.soft_square(coords, self.width, clip, self.occulting)

    @property
    def extent(self):                                   # apertures.py:382-383
        return math.sqrt(2) * self.width

    nsides = 4                                          # :386-387


class RectangularAperture(_DynamicAperture):
    def __init__(self, height, width, transformation=None, occulting=False, softening=1.0, normalise=False):
        super().__init__(transformation, occulting, softening, normalise)
        self.height = _param(height, (), "height")
        self.width = _param(width, (), "width")

    def _shape(self, coords, clip):
        return G.soft_rectangle(coords, self.width, self.height, clip, self.occulting)

    @property
    def extent(self):                                   # apertures.py:467-468
        if torch.is_tensor(self.height) or torch.is_tensor(self.width):
            return torch.hypot(torch.as_tensor(self.height) / 2.0, torch.as_tensor(self.width) / 2.0)
        return np.float32(np.hypot(self.height / 2.0, self.width / 2.0))

    nsides = 4                                          # :471-472


class RegPolyAperture(_DynamicAperture):
    def __init__(self, nsides: int, rmax, transformation=None, occulting=False, softening=1.0,
                 normalise=False):
        super().__init__(transformation, occulting, softening, normalise)
        self.nsides = int(nsides)
        self.rmax = _param(rmax, (), "rmax")

    def _shape(self, coords, clip):
        return G.soft_reg_polygon(coords, self.rmax, self.nsides, clip, self.occulting)

    @property
    def extent(self):                                   # apertures.py:550-551
        return self.rmax


class Spider(_DynamicAperture):
    def __init__(self, width, angles, transformation=None, occulting=False, softening=1.0, normalise=False):
        super().__init__(transformation, occulting, softening, normalise)
        self.width = _param(width, (), "width")
        self.angles = angles if torch.is_tensor(angles) else np.atleast_1d(np.asarray(angles, np.float32))

    def _shape(self, coords, clip):
        return G.soft_spider(coords, self.width, self.angles, clip, self.occulting)


class _Composite(_DynamicAperture):
    def __init__(self, apertures, transformation=None, normalise=False):
        super().__init__(transformation, False, 1.0, normalise)
        if isinstance(apertures, (list, tuple)):
            od = OrderedDict()
            for i, a in enumerate(apertures):
                key, ap = a if isinstance(a, tuple) else (f"{type(a).__name__}_{i}", a)
                od[key] = ap
            apertures = od
        for ap in apertures.values():
            if not isinstance(ap, _DynamicAperture):
                raise TypeError("apertures must be dynamic aperture layers")
        self.apertures = OrderedDict(apertures)

    def __getattr__(self, key):
        aps = self.__dict__.get("apertures", {})
        if key in aps:
            return aps[key]
        raise AttributeError(key)

    def transmissions(self, coords, pixel_scale):      # apertures.py:950-969
        return torch.stack([ap.transmission(coords, pixel_scale) for ap in self.apertures.values()])

    def transmission(self, coords, pixel_scale):
        if self.transformation is not None:
            coords = self.transformation(coords)
        return self._join(self.transmissions(coords, pixel_scale))


class CompoundAperture(_Composite):
    """Overlapping shapes multiplied together (apertures.py:1005-1069), e.g. primary x secondary
    obstruction x spiders."""

    def _join(self, ts):
        return ts.prod(0)


class MultiAperture(_Composite):
    """Separate sub-apertures summed (apertures.py:1072-1117), e.g. the holes of an NRM."""

    def _join(self, ts):
        return ts.sum(0)
