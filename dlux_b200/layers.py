"""Mirror of the pupil-plane and propagator layers that feed the MFT
(/root/reference/src/dLux/layers/optical_layers.py:80-371, layers/optics.py:77-175,
layers/propagators.py:145-217).  Same class names, constructor arguments and
``layer(wavefront) -> wavefront`` contract; arrays are torch CUDA tensors."""
from __future__ import annotations

import numpy as np
import torch

from .utils import propagation as _prop
from .wavefronts import CoordSpec, Wavefront

__all__ = ["OpticalLayer", "TransmissiveLayer", "AberratedLayer", "BasisLayer", "Tilt", "Normalise",
           "Optic", "BasisOptic", "MFT", "FFT", "UnifiedLayer", "Resize", "Rotate", "Flip", "Lambda"]


def _arr(x, device=None):
    if x is None:
        return None
    if torch.is_tensor(x):
        return x.to(dtype=torch.float32, device=device or x.device)
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=device or "cuda")


class OpticalLayer:
    """optical_layers.py:80-105: the reference's user-extension API."""

    def apply(self, wavefront: Wavefront) -> Wavefront:
        return self(wavefront)

    def __call__(self, wavefront: Wavefront) -> Wavefront:  # pragma: no cover - abstract
        raise NotImplementedError


class TransmissiveLayer(OpticalLayer):
    def __init__(self, transmission=None, normalise: bool = False, device=None):
        self.transmission = _arr(transmission, device)
        self.normalise = bool(normalise)

    def __call__(self, wavefront):                     # optical_layers.py:181-186
        if self.transmission is not None:
            wavefront = wavefront * self.transmission
        if self.normalise:
            wavefront = wavefront.normalise()
        return wavefront


class AberratedLayer(OpticalLayer):
    def __init__(self, opd=None, phase=None, device=None):
        self.opd = _arr(opd, device)
        self.phase = _arr(phase, device)
        if self.opd is not None and self.phase is not None and self.opd.shape != self.phase.shape:
            raise ValueError("opd and phase must have the same shape. Got "
                             f"shapes {tuple(self.opd.shape)} and {tuple(self.phase.shape)}.")

    def __call__(self, wavefront):                     # optical_layers.py:233-236
        return wavefront.add_opd(self.opd).add_phase(self.phase)


class BasisLayer(OpticalLayer):
    def __init__(self, basis=None, coefficients=None, effect: str = "opd", device=None,
                 coefficient_shape=None):
        self.basis = _arr(basis, device)
        if coefficients is None and self.basis is not None:    # optical_layers.py:284-296
            shape = tuple(self.basis.shape[:-2]) if coefficient_shape is None else tuple(coefficient_shape)
            coefficients = torch.zeros(shape, dtype=torch.float32, device=self.basis.device)
        self.coefficients = _arr(coefficients, device)
        if self.basis is not None and tuple(self.basis.shape[:self.coefficients.dim()]) != tuple(
                self.coefficients.shape):
            raise ValueError("The coefficient shape must match the leading basis dimensions.")
        if effect not in ("opd", "phase", "amplitude"):
            raise ValueError("effect must be 'opd', 'phase', or 'amplitude'.")
        self.effect = effect

    def eval_basis(self):                              # optical_layers.py:304-313
        return _prop.eval_basis(self.basis, self.coefficients)

    def __call__(self, wavefront):                     # optical_layers.py:319-327
        output = self.eval_basis()
        if self.effect == "phase":
            return wavefront.add_phase(output)
        if self.effect == "opd":
            return wavefront.add_opd(output)
        return wavefront * (1 + output)


class Tilt(OpticalLayer):
    def __init__(self, angles, device=None):
        self.angles = _arr(angles, device)
        if tuple(self.angles.shape) != (2,):
            raise ValueError("angles must have shape (2,).")

    def __call__(self, wavefront):                     # optical_layers.py:358-359
        return wavefront.tilt(self.angles)


class Normalise(OpticalLayer):
    def __call__(self, wavefront):                     # optical_layers.py:370-371
        return wavefront.normalise()


class Optic(TransmissiveLayer, AberratedLayer):
    def __init__(self, transmission=None, opd=None, phase=None, normalise: bool = False, device=None):
        TransmissiveLayer.__init__(self, transmission, normalise, device)
        AberratedLayer.__init__(self, opd, phase, device)
        for name in ("opd", "phase"):
            v = getattr(self, name)
            if self.transmission is not None and v is not None and v.shape != self.transmission.shape:
                raise ValueError(f"transmission and {name} must have the same shape. Got shapes "
                                 f"{tuple(self.transmission.shape)} and {tuple(v.shape)}.")

    def __call__(self, wavefront):                     # layers/optics.py:91-96
        if self.transmission is not None:
            wavefront = wavefront * self.transmission
        wavefront = wavefront.add_opd(self.opd).add_phase(self.phase)
        if self.normalise:
            wavefront = wavefront.normalise()
        return wavefront


class BasisOptic(TransmissiveLayer, BasisLayer):
    def __init__(self, basis, transmission=None, coefficients=None, normalise: bool = False,
                 effect: str = "opd", coefficient_shape=None, device=None):
        # layers/optics.py:122-152 argument order
        TransmissiveLayer.__init__(self, transmission, normalise, device)
        BasisLayer.__init__(self, basis, coefficients, effect, device, coefficient_shape)

    def __call__(self, wavefront):                     # layers/optics.py:168-175
        if self.transmission is not None:
            wavefront = wavefront * self.transmission
        wavefront = BasisLayer.__call__(self, wavefront)
        if self.normalise:
            wavefront = wavefront.normalise()
        return wavefront


class MFT(OpticalLayer):
    """layers/propagators.py:145-217."""

    def __init__(self, npixels: int, pixel_scale, focal_length=None, inverse: bool = False):
        self.npixels = int(npixels)
        self.pixel_scale = np.float32(pixel_scale)
        self.focal_length = None if focal_length is None else np.float32(focal_length)
        self.inverse = bool(inverse)

    def __call__(self, wavefront):                     # layers/propagators.py:198-217
        return wavefront.propagate(self.npixels, self.pixel_scale, self.focal_length, self.inverse)


class FFT(OpticalLayer):
    """layers/propagators.py:57-142: padded FFT propagation followed by a centre crop."""

    def __init__(self, focal_length=None, inverse: bool = False, pad: int = 1, crop: int = 1,
                 center: bool = True):
        self.focal_length = None if focal_length is None else np.float32(focal_length)
        self.inverse = bool(inverse)
        self.pad = int(pad)
        self.crop = int(crop)
        self.center = bool(center)

    def __call__(self, wavefront):                     # layers/propagators.py:118-142
        spec = CoordSpec(c=0.0) if self.center else None
        size_out = wavefront.npixels * self.pad // self.crop
        return wavefront.propagate_FFT(pad=self.pad, focal_length=self.focal_length, inverse=self.inverse,
                                       spec_out=spec).resize(size_out)


class UnifiedLayer(OpticalLayer):
    """layers/unified_layers.py:14-22: layers that act on a ``Wavefront`` or a ``PSF`` alike."""


class Resize(UnifiedLayer):
    """unified_layers.py:25-66: centre-preserving crop / zero pad to ``npixels`` (even -> even, odd -> odd)."""

    def __init__(self, npixels: int):
        self.npixels = int(npixels)

    def __call__(self, target):
        return target.resize(self.npixels)


class Rotate(UnifiedLayer):
    """unified_layers.py:69-133: rotation by ``angle`` radians through interpolation; ``complex`` picks the (real,
    imaginary) or the (amplitude, phase) fields of a wavefront and is ignored for a PSF."""

    def __init__(self, angle, method: str = "linear", complex: bool = False):
        self.angle = angle if torch.is_tensor(angle) else np.asarray(angle, dtype=np.float32)
        self.method = str(method)
        self.complex = bool(complex)

    def __call__(self, target):
        from .psfs import PSF
        if isinstance(target, PSF):
            return target.rotate(self.angle, self.method)
        return target.rotate(self.angle, self.method, self.complex)


class Flip(UnifiedLayer):
    """unified_layers.py:136-186: flip about the given axes ('ij' convention: axis 0 = y, axis 1 = x)."""

    def __init__(self, axes):
        self.axes = axes
        if isinstance(axes, tuple):
            if not all(isinstance(a, int) for a in axes):
                raise ValueError("All axes must be integers.")
        elif not isinstance(axes, int):
            raise ValueError("axes must be an int or tuple of ints.")

    def __call__(self, target):
        return target.flip(self.axes)


class Lambda(UnifiedLayer):
    """unified_layers.py:189-212: returns its input unchanged (placeholder in a layer list)."""

    def __call__(self, target):
        return target

