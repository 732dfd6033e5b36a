"""Mirror of ``dLux.wavefronts.Wavefront`` (/root/reference/src/dLux/wavefronts.py:17)
restricted to what the MFT hot path touches.  State lives in torch CUDA tensors; the
pupil-plane methods are plain elementwise plumbing for the layer-by-layer (unfused)
route -- the fused route is ``OpticalSystem.propagate`` (optical_systems.py here)."""
from __future__ import annotations

import copy
import math

import numpy as np
import torch

from .utils import propagation as _prop

__all__ = ["Wavefront", "CoordSpec"]


class CoordSpec:
    """Minimal mirror of ``dLux.coordinates.CoordSpec`` (coordinates.py:74-157): n pixels of
    size d centred on c."""

    def __init__(self, n=None, d=None, c=0.0):
        self.n, self.d, self.c = n, d, c

    def set(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    def xs(self, device):
        if self.d is None:
            raise ValueError("d must be specified to calculate coordinates.")
        idx = torch.arange(self.n, dtype=torch.float32, device=device)
        to_t = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float32), device=device)
        return to_t(self.c) + (idx - np.float32((self.n - 1) / 2)) * to_t(self.d)


class Wavefront:
    def __init__(self, wavelength, npixels: int, diameter=None, pixel_scale=None, center=None,
                 device=None):
        # wavefronts.py:96-122
        if diameter is None and pixel_scale is None:
            raise ValueError("Provide one: diameter or pixel_scale.")
        if diameter is not None and pixel_scale is not None:
            raise ValueError(
                "Cannot specify both 'diameter' and 'pixel_scale' - they are "
                "interdependent (diameter = pixel_scale × npixels). Choose one: "
                "use 'diameter' for wavefront diameter, or 'pixel_scale' for "
                "wavefront sampling.")
        device = torch.device("cuda" if device is None else device)
        f = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float32), device=device)
        # a 1-d `wavelength` makes a batched wavefront, phasor [L, N, N]: what the reference gets
        # from vmapping propagate_mono over the wavelengths (optical_systems.py:213-216)
        self.wavelength = f(wavelength)
        self.pixel_scale = f(np.float32(diameter) / np.float32(npixels)) if diameter is not None \
            else f(pixel_scale)
        amp = torch.full(tuple(self.wavelength.shape) + (npixels, npixels), 1.0 / npixels ** 2,
                         dtype=torch.float32, device=device)
        self.phasor = torch.complex(amp, torch.zeros_like(amp))
        if center is not None:
            self.center = f(center)
            if tuple(self.center.shape) != (1,):
                raise ValueError("center must have shape (1,).")
        else:
            self.center = torch.zeros(1, dtype=torch.float32, device=device)

    # ------------------------------------------------------------------ zodiax-like
    def set(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    def multiply(self, name, value):
        return self.set(**{name: getattr(self, name) * value})

    def add(self, name, value):
        return self.set(**{name: getattr(self, name) + value})

    @classmethod
    def from_phasor(cls, phasor, wavelength, pixel_scale=None, diameter=None, center=None, device=None):
        """wavefronts.py:124-167: a wavefront around an existing complex field."""
        if not torch.is_tensor(phasor):
            phasor = torch.as_tensor(np.asarray(phasor, dtype=np.complex64),
                                     device=torch.device("cuda" if device is None else device))
        wf = cls(wavelength, phasor.shape[-1], diameter, pixel_scale, center, device=phasor.device)
        return wf.set(phasor=phasor.to(torch.complex64))

    def _magic(self, other, op):                       # wavefronts.py:834-878
        if other is None:
            return self
        if isinstance(other, Wavefront):
            other = other.phasor
        elif not (torch.is_tensor(other) or isinstance(other, (np.ndarray, np.generic, float, int, complex))):
            raise TypeError(f"Unsupported type for {op}: {type(other)}. Must be an array, "
                            "Wavefront, or None.")
        if isinstance(other, np.ndarray):
            other = torch.as_tensor(other, device=self.phasor.device)
        if op == "add":
            return self.add("phasor", other)
        if op == "subtract":
            return self.add("phasor", -other)
        if op == "multiply":
            return self.multiply("phasor", other)
        return self.multiply("phasor", 1 / other)

    def __add__(self, other):
        return self._magic(other, "add")

    def __sub__(self, other):
        return self._magic(other, "subtract")

    def __mul__(self, other):
        return self._magic(other, "multiply")

    def __truediv__(self, other):
        return self._magic(other, "divide")

    __iadd__, __isub__, __imul__, __itruediv__ = __add__, __sub__, __mul__, __truediv__

    def flip(self, axis):                              # wavefronts.py:426-440 (0 = y, 1 = x)
        axes = (axis,) if isinstance(axis, int) else tuple(axis)
        nd = self.phasor.dim()
        return self.set(phasor=torch.flip(self.phasor, [a + nd - 2 if a >= 0 else a for a in axes]))

    # ------------------------------------------------------------------ properties
    @property
    def npixels(self):
        return self.phasor.shape[-1]

    @property
    def diameter(self):
        return self.npixels * self.pixel_scale

    @property
    def real(self):
        return self.phasor.real

    @property
    def imaginary(self):
        return self.phasor.imag

    @property
    def complex(self):                                 # wavefronts.py:244-254: [real, imaginary]
        return torch.stack([self.phasor.real, self.phasor.imag])

    @property
    def polar(self):                                   # wavefronts.py:256-266: [amplitude, phase]
        return torch.stack([self.amplitude, self.phase])

    @property
    def ndim(self):
        return self.pixel_scale.dim()

    @property
    def amplitude(self):
        return self.phasor.abs()

    @property
    def phase(self):
        return self.phasor.angle()

    @property
    def psf(self):                                     # wavefronts.py:279
        return self.phasor.real ** 2 + self.phasor.imag ** 2

    def to_psf(self):                                  # wavefronts.py:281-292
        from .psfs import PSF
        return PSF(self.psf, self.pixel_scale)

    @property
    def wavenumber(self):                              # wavefronts.py:303
        return np.float32(2 * math.pi) / self.wavelength

    @property
    def power(self):                                   # wavefronts.py:330 (per wavelength when batched)
        return self.psf.sum((-2, -1))

    @property
    def xs(self):                                      # coordinates.py:129
        n = self.npixels
        idx = torch.arange(n, dtype=torch.float32, device=self.phasor.device)
        return self.center + (idx - np.float32((n - 1) / 2)) * self.pixel_scale

    def coordinates(self, scale=1.0):                  # wavefronts.py:584-609
        xs = self.xs * np.float32(scale)
        X, Y = torch.meshgrid(xs, xs, indexing="xy")
        return torch.stack([X, Y])

    # ------------------------------------------------------------------ pupil ops
    def add_phase(self, phase):                        # wavefronts.py:332-349
        if phase is None:
            return self
        phase = torch.as_tensor(phase, dtype=torch.float32, device=self.phasor.device)
        return self.multiply("phasor", torch.polar(torch.ones_like(phase), phase))

    def add_opd(self, opd):                            # wavefronts.py:351-368
        if opd is None:
            return self
        opd = torch.as_tensor(opd, dtype=torch.float32, device=self.phasor.device)
        k = self.wavenumber
        return self.add_phase((k[..., None, None] if k.dim() else k) * opd)

    def tilt(self, angles, unit: str = "rad"):         # wavefronts.py:370-395
        angles = torch.as_tensor(np.asarray(angles, dtype=np.float32) if not torch.is_tensor(angles)
                                 else angles, dtype=torch.float32, device=self.phasor.device)
        if tuple(angles.shape) != (2,):
            raise ValueError("angles must be a 1d array of shape (2,).")
        if unit != "rad":
            raise ValueError("only unit='rad' is supported")
        coords = self.coordinates()
        return self.add_opd((angles[:, None, None] * coords).sum(0))

    def normalise(self, mode: str = "power", value: float = 1.0):   # wavefronts.py:397-424
        if mode == "power":
            scale = torch.sqrt(np.float32(value) / self.power)
        elif mode == "peak":
            scale = torch.sqrt(np.float32(value) / self.psf.amax((-2, -1)))
        else:
            raise ValueError("mode must be 'power' or 'peak'")
        return self.multiply("phasor", scale[..., None, None] if scale.dim() else scale)

    # ------------------------------------------------------------------ propagation
    def propagate(self, npixels: int, pixel_scale, focal_length=None, inverse: bool = False,
                  precision=None):
        """wavefronts.py:729-772 -> dlu.MFT."""
        phasor = _prop.MFT(self.phasor, self.wavelength, self.pixel_scale, npixels, pixel_scale,
                           focal_length=focal_length, inverse=bool(inverse), precision=precision)
        # a device tensor stays in the autograd graph: a later propagate reads it as pixel_scale_in
        ps = pixel_scale.to(phasor.device, torch.float32) if torch.is_tensor(pixel_scale) else torch.as_tensor(
            np.asarray(pixel_scale, dtype=np.float32), device=phasor.device)
        return self.set(phasor=phasor, pixel_scale=ps)

    def propagate_MFT(self, spec_out, focal_length=None, inverse=None, precision=None):
        """wavefronts.py:774-805; ``spec_out`` needs attributes ``n`` and ``d``."""
        return self.propagate(spec_out.n, spec_out.d, focal_length, bool(inverse), precision)

    def propagate_FFT(self, pad: int = 2, focal_length=None, inverse: bool = False, spec_out=None,
                      precision=None):
        """wavefronts.py:654-727 -> dlu.FFT, with the optional input/output phase ramps that
        re-centre the FFT grid on ``spec_out.c``."""
        wl = self.wavelength
        n_out = self.npixels * pad
        d_fft, c_fft = _prop.fft_spec(n_out, self.pixel_scale, wl, focal_length)
        in_ramp = out_ramp = None
        if spec_out is not None:
            if spec_out.d is not None:
                raise ValueError("Output spec cannot specify d; FFT output d is fixed.")
            if spec_out.n is not None:
                raise ValueError("Output spec cannot specify n; FFT output n is determined by the "
                                 "pad parameter.")
            dev = self.phasor.device
            to_t = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float32), device=dev)
            shift = to_t(c_fft) - to_t(spec_out.c)
            in_ramp = _prop.fft_phase_ramp(self.xs, wl, shift, focal_length, inverse)
            spec_out = spec_out.set(n=n_out, d=d_fft)
            shift = _prop.fft_spec(spec_out.n, spec_out.d, wl, focal_length)[1]
            out_ramp = _prop.fft_phase_ramp(spec_out.xs(dev), wl, shift, focal_length, inverse)
            center = to_t(spec_out.c).reshape(1)
        else:
            center = torch.as_tensor(np.asarray(c_fft.detach().cpu() if torch.is_tensor(c_fft) else c_fft,
                                                dtype=np.float32), device=self.phasor.device).reshape(1)
        phasor = self.phasor if in_ramp is None else self.phasor * in_ramp
        phasor, pixel_scale = _prop.FFT(phasor, wl, self.pixel_scale, focal_length, pad, inverse, precision)
        if out_ramp is not None:
            phasor = phasor * out_ramp
        ps = pixel_scale if torch.is_tensor(pixel_scale) else torch.as_tensor(
            np.asarray(pixel_scale, dtype=np.float32), device=phasor.device)
        return self.set(phasor=phasor, pixel_scale=ps, center=center)

    # ------------------------------------------------------------------ interpolating operations (off the hot path)
    def _resample(self, fn, complex: bool):
        """Applies ``fn`` to the two real fields of a monochromatic wavefront -- (real, imaginary) or (amplitude,
        phase) -- and reassembles the phasor (wavefronts.py:472-483, 511-524, 552-566)."""
        if self.phasor.dim() != 2:
            raise ValueError("dlux_b200: interpolating Wavefront operations act on a monochromatic (2-D) wavefront")
        a, b = self.complex if complex else self.polar
        a, b = fn(a), fn(b)
        return torch.complex(a, b) if complex else torch.polar(a, b)

    def scale_to(self, npixels: int, pixel_scale, complex: bool = True):        # wavefronts.py:442-483
        from .utils import interpolation as _interp
        ps = pixel_scale if torch.is_tensor(pixel_scale) else torch.as_tensor(
            np.asarray(pixel_scale, dtype=np.float32), device=self.phasor.device)
        ratio = ps / self.pixel_scale
        return self.set(phasor=self._resample(lambda f: _interp.scale(f, int(npixels), ratio), complex), pixel_scale=ps)

    def interpolate(self, transformation, method: str = "linear", complex: bool = True, fill: float = 0.0):
        from .apertures import CoordTransform                                    # wavefronts.py:485-524
        from .utils import interpolation as _interp
        if not isinstance(transformation, CoordTransform):
            raise TypeError("transformation must be a BaseCoordTransform.")
        knots = self.coordinates()
        samples = transformation(knots)
        return self.set(phasor=self._resample(lambda f: _interp.interp(f, knots, samples, method, fill), complex))

    def rotate(self, angle, method: str = "linear", complex: bool = True):       # wavefronts.py:526-566
        from .utils import interpolation as _interp
        return self.set(phasor=self._resample(lambda f: _interp.rotate(f, angle, method), complex))

    def resize(self, npixels: int):
        """wavefronts.py:562-582 -> dlu.resize: centre-preserving crop / zero pad."""
        n_in = self.npixels
        if npixels == n_in:
            return self
        if n_in % 2 != npixels % 2:
            raise ValueError("Center-preserving resizing requires parity consistency, i.e. even -> even "
                             f"or odd -> odd: {n_in} -> {npixels}.")
        if npixels < n_in:
            a, b = (n_in - npixels) // 2, (n_in + npixels) // 2
            return self.set(phasor=self.phasor[..., a:b, a:b])
        p = (npixels - n_in) // 2
        return self.set(phasor=torch.nn.functional.pad(self.phasor, (p, p, p, p)))
