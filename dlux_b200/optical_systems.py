"""Mirror of the optical-system classes around the MFT
(/root/reference/src/dLux/optical_systems.py:30-775): ``propagate_mono``,
``propagate``, ``model`` with the reference's argument checks and messages.

Two routes produce the same numbers:
 * layer-by-layer (``fused=False``): Wavefront -> layers -> ``Wavefront.propagate`` ->
   ``dlux_mft_c64`` per wavelength batch; works for any layer stack;
 * fused (default when the stack is pupil-only): one ``dlux_polypsf_fwd`` call does
   the pupil phasor, both contractions, |E|^2, spectral weights and the source sum on
   the GPU, and ``dlux_polypsf_bwd`` is its VJP.
"""
from __future__ import annotations

from collections import OrderedDict
import math

import numpy as np
import torch

from . import ops
from .layers import (AberratedLayer, BasisLayer, BasisOptic, Normalise, Optic,
                     TransmissiveLayer)
from .utils import propagation as _prop
from .psfs import PSF
from .wavefronts import Wavefront

__all__ = ["BaseOpticalSystem", "OpticalSystem", "ParametricOpticalSystem", "LayeredOpticalSystem",
           "ParametricLayeredOpticalSystem", "AngularOpticalSystem", "CartesianOpticalSystem"]


def _np32(x):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


class BaseOpticalSystem:
    """optical_systems.py:30-135: the abstract interface (propagate_mono / propagate / model)."""

    def propagate_mono(self, wavelength, offset=None, return_wf=False):  # pragma: no cover
        raise NotImplementedError

    def propagate(self, wavelengths, offset=None, weights=None, return_wf=False, return_psf=False):  # pragma: no cover
        raise NotImplementedError

    def model(self, source, return_wf=False, return_psf=False):  # pragma: no cover
        raise NotImplementedError


class OpticalSystem(BaseOpticalSystem):
    """optical_systems.py:138-231: polychromatic ``propagate`` and ``model``."""

    def propagate(self, wavelengths, offset=None, weights=None, return_wf=False, return_psf=False):
        """optical_systems.py:147-223.  Returns the PSF array, or with ``return_wf`` the vectorised
        ``Wavefront`` (phasor [L, M, M], sqrt(weight) applied, :213-220), or with ``return_psf`` a
        ``PSF(psf, mean pixel scale)`` (:221-222)."""
        if return_wf and return_psf:
            raise ValueError(
                "Cannot return both Wavefront and PSF objects. Choose one: "
                "return_wf=True for Wavefront, or return_psf=True for PSF.")
        wl_t = None
        if torch.is_tensor(wavelengths) and wavelengths.is_cuda and wavelengths.requires_grad:
            wl_t = torch.atleast_1d(wavelengths)       # differentiable leaf of the fused route
        wavelengths = np.atleast_1d(_np32(wavelengths))
        if weights is None:
            weights = np.ones_like(wavelengths) / np.float32(len(wavelengths))
        else:
            weights = np.atleast_1d(_np32(weights)) if not (torch.is_tensor(weights) and weights.is_cuda) \
                else torch.atleast_1d(weights)
        if tuple(weights.shape) != tuple(wavelengths.shape):
            raise ValueError(
                f"Wavelength and weight shape mismatch: "
                f"wavelengths {tuple(wavelengths.shape)} vs weights {tuple(weights.shape)}. "
                f"Must have same length and dimensions.")
        offset = np.zeros(2, np.float32) if offset is None else (
            offset if torch.is_tensor(offset) else _np32(offset))
        if tuple(offset.shape) != (2,):
            raise ValueError(
                f"offset must be [x, y] array of shape (2,), "
                f"got shape {tuple(offset.shape)}. "
                "Pass offset as [on_axis_x, on_axis_y] angles in radians.")
        if wl_t is not None:
            if return_wf or not (getattr(self, "fused", False) and hasattr(self, "fused_propagate")
                                 and self._can_fuse()):
                raise ValueError("differentiable wavelengths need the fused route (pupil-only layer stack)")
            off = offset if torch.is_tensor(offset) else _np32(offset)
            psf = self.fused_propagate(wl_t, off.reshape(1, 2), weights.reshape(1, -1))
            return PSF(psf, self._psf_pixel_scale_out(psf.device)) if return_psf else psf
        out = self._propagate(wavelengths, offset, weights, return_wf)
        if return_wf:
            return out
        return PSF(out, self._psf_pixel_scale_out(out.device)) if return_psf else out

    def _psf_pixel_scale_out(self, device):
        """``wf.pixel_scale.mean()`` of optical_systems.py:222 -- known without a wavefront for the
        parametric systems; a generic layer stack propagates one wavelength's geometry to find it."""
        raise NotImplementedError

    def _batchable(self):
        """True when every layer is one of this package's elementwise / MFT layers, which
        broadcast over a leading wavelength axis (user layers and FFT keep the per-wavelength
        loop)."""
        return False

    def _propagate(self, wavelengths, offset, weights, return_wf):
        if self._batchable() and len(wavelengths) > 1:
            # the reference vmaps propagate_mono over (wavelength, weight): one batched wavefront
            wf = self.propagate_mono(wavelengths, offset, return_wf=True)
            w = weights if torch.is_tensor(weights) else torch.as_tensor(weights, device=wf.phasor.device)
            fields = wf.phasor * (w.to(torch.float32) ** 0.5)[:, None, None]
        else:
            fields = []
            for wl, w in zip(wavelengths, weights):
                wf = self.propagate_mono(wl, offset, return_wf=True)
                fields.append(wf.phasor * (w ** 0.5))
            fields = torch.stack(fields)
        if return_wf:
            # the vectorised Wavefront that filter_vmap returns: every leaf gains the wavelength axis
            L = fields.shape[0]
            ps = wf.pixel_scale.reshape(-1)
            ctr = wf.center.reshape(-1, 1) if wf.center.dim() > 1 else wf.center.reshape(1, 1)
            return wf.set(phasor=fields, wavelength=torch.as_tensor(wavelengths, device=fields.device),
                          pixel_scale=ps.expand(L).clone() if ps.numel() == 1 else ps,
                          center=ctr.expand(L, 1).clone() if ctr.shape[0] == 1 else ctr)
        self.__dict__["_last_pixel_scale"] = wf.pixel_scale.reshape(-1)[:1].mean()
        return (fields.real ** 2 + fields.imag ** 2).sum(0)

    def model(self, source, return_wf=False, return_psf=False):   # optical_systems.py:225-231
        return source.model(self, return_wf, return_psf)


class ParametricOpticalSystem(OpticalSystem):
    """optical_systems.py:234-292: a system with a parametrised output sampling."""

    def _init_parametric(self, psf_npixels, psf_pixel_scale, oversample=1):
        self.psf_npixels = int(psf_npixels)
        self.oversample = int(oversample)
        # a torch scalar (requires_grad) makes the pixel scale a fitted parameter of the fused route
        self.psf_pixel_scale = (psf_pixel_scale if torch.is_tensor(psf_pixel_scale)
                                else np.float32(psf_pixel_scale))

    @property
    def fov(self):                                      # :283-292
        return self.psf_npixels * self.psf_pixel_scale


class LayeredOpticalSystem(OpticalSystem):
    """optical_systems.py:298-507."""

    def __init__(self, wf_npixels: int, diameter, layers, device=None, fused=True, precision=None,
                 sparse=False):
        # sparse=True: opt-in exact zero-block skipping on the fused route (blocks of the pupil where the
        # transmission is zero are neither contracted nor produced); bit-identical results, fewer FLOPs
        self.sparse = bool(sparse)
        self.wf_npixels = int(wf_npixels)
        self.diameter = np.float32(diameter)
        if isinstance(layers, (list, tuple)):
            od = OrderedDict()
            for i, l in enumerate(layers):
                if isinstance(l, tuple):
                    od[l[0]] = l[1]
                else:
                    od[f"{type(l).__name__}_{i}" if type(l).__name__ in od else type(l).__name__] = l
            layers = od
        self.layers = OrderedDict(layers)
        self.device = torch.device("cuda" if device is None else device)
        self.fused = bool(fused)
        self.precision = precision

    def __getattr__(self, key):                        # zodiax-style access to layers by name
        layers = self.__dict__.get("layers", {})
        if key in layers:
            return layers[key]
        raise AttributeError(key)

    def _batchable(self):
        from .layers import MFT as _MFTLayer, Tilt as _Tilt
        from .apertures import AberratedAperture, _DynamicAperture
        ok = (TransmissiveLayer, AberratedLayer, BasisLayer, Optic, BasisOptic, Normalise, _MFTLayer, _Tilt)
        return all(type(l) in ok or isinstance(l, (_DynamicAperture, AberratedAperture)) for l in self.layers.values())

    def initialise_wavefront(self, wavelength, offset=None):      # optical_systems.py:363-389
        wf = Wavefront(wavelength, self.wf_npixels, self.diameter, device=self.device)
        return wf.tilt(np.zeros(2, np.float32) if offset is None else offset)

    def propagate_mono(self, wavelength, offset=None, return_wf=False):   # :391-425
        wf = self.initialise_wavefront(wavelength, offset)
        for layer in self.layers.values():
            wf = layer(wf)
        return wf if return_wf else wf.psf

    def _psf_pixel_scale_out(self, device):
        ps = self.__dict__.get("_last_pixel_scale")
        if ps is None:
            raise ValueError("return_psf needs a propagated wavefront to read the pixel scale from")
        return ps

    def insert_layer(self, layer, index: int):           # :465-491
        items = list(self.layers.items())
        items.insert(index, layer if isinstance(layer, tuple) else (type(layer).__name__, layer))
        import copy
        new = copy.copy(self)
        new.layers = OrderedDict(items)
        return new

    def remove_layer(self, key: str):                    # :493-507
        import copy
        new = copy.copy(self)
        new.layers = OrderedDict((k, v) for k, v in self.layers.items() if k != key)
        return new


class ParametricLayeredOpticalSystem(ParametricOpticalSystem, LayeredOpticalSystem):
    """optical_systems.py:510-594: pupil-plane layers followed by one MFT ``to_focus`` whose output
    sampling is parametrised (``psf_npixels``, ``psf_pixel_scale``, ``oversample``).  Subclasses
    supply the unit conversion (``_focal_args``); this is the class the fused route hangs on."""

    def __init__(self, wf_npixels, diameter, layers, psf_npixels, psf_pixel_scale, oversample=1,
                 **kw):
        LayeredOpticalSystem.__init__(self, wf_npixels, diameter, layers, **kw)
        self._init_parametric(psf_npixels, psf_pixel_scale, oversample)

    def _psf_pixel_scale_out(self, device):
        ps = self._focal_args()[1]
        return ps.to(device) if torch.is_tensor(ps) else torch.as_tensor(np.float32(ps), device=device)

    # -- units --------------------------------------------------------------------
    def _focal_args(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def to_focus(self, wavefront):
        npix, ps, fl = self._focal_args()
        return wavefront.propagate(npix, ps, fl, precision=self.precision)

    def propagate_mono(self, wavelength, offset=None, return_wf=False):   # :560-594
        wf = LayeredOpticalSystem.propagate_mono(self, wavelength, offset, return_wf=True)
        wf = self.to_focus(wf)
        return wf if return_wf else wf.psf

    # -- fused route ----------------------------------------------------------------
    def _can_fuse(self) -> bool:
        """Structure-only version of ``_fusable`` (no layer is evaluated): is the stack pupil-only,
        with no transmission applied after a normalisation?"""
        from .apertures import AberratedAperture, _DynamicAperture
        normalise = False
        for layer in self.layers.values():
            if isinstance(layer, AberratedAperture):
                if normalise or (layer.effect == "amplitude" and layer.normalise):
                    return False                       # a transmission applied after a normalisation
                normalise = normalise or layer.normalise
                continue
            dyn = isinstance(layer, _DynamicAperture)
            if not dyn and type(layer) not in (TransmissiveLayer, AberratedLayer, BasisLayer, Optic,
                                                BasisOptic, Normalise):
                return False
            has_t = dyn or getattr(layer, "transmission", None) is not None or (
                isinstance(layer, BasisLayer) and layer.effect not in ("opd", "phase"))
            if has_t and normalise:
                return False
            if isinstance(layer, Normalise) or getattr(layer, "normalise", False):
                normalise = True
        return True

    def _fusable(self):
        """Collapse a pupil-only stack into (T, opd, phase, normalise) or None."""
        if not self._can_fuse():
            return None
        T = opd = phase = None
        normalise = False
        t_after_norm = False

        def mul(a, b):
            return b if a is None else (a if b is None else a * b)

        def add(a, b):
            return b if a is None else (a if b is None else a + b)

        from .apertures import AberratedAperture, _DynamicAperture
        for layer in self.layers.values():
            if isinstance(layer, AberratedAperture):
                coords = self.__dict__.get("_pupil_coords")
                if coords is None:
                    from .utils.geometry import pixel_coords
                    coords = self.__dict__["_pupil_coords"] = pixel_coords(
                        self.wf_npixels, float(self.diameter), device=self.device)
                ps = torch.as_tensor(np.float32(self.diameter / np.float32(self.wf_npixels)), device=self.device)
                T = mul(T, layer.transmission(coords, ps))
                ab = layer.eval_basis(coords)
                if layer.effect == "opd":
                    opd = add(opd, ab)
                elif layer.effect == "phase":
                    phase = add(phase, ab)
                else:
                    T = mul(T, 1 + ab)
                if layer.normalise:
                    normalise = True
                continue
            if isinstance(layer, _DynamicAperture):
                # a dynamic aperture evaluates its transmission on the pupil grid (torch, autograd)
                coords = self.__dict__.get("_pupil_coords")
                if coords is None:
                    from .utils.geometry import pixel_coords
                    coords = self.__dict__["_pupil_coords"] = pixel_coords(
                        self.wf_npixels, float(self.diameter), device=self.device)
                ps = torch.as_tensor(np.float32(self.diameter / np.float32(self.wf_npixels)), device=self.device)
                t = layer.transmission(coords, ps)
                if normalise:
                    t_after_norm = True
                T = mul(T, t)
                if layer.normalise:
                    normalise = True
                continue
            if type(layer) not in (TransmissiveLayer, AberratedLayer, BasisLayer, Optic, BasisOptic,
                                   Normalise):
                return None
            t = getattr(layer, "transmission", None)
            if t is not None:
                if normalise:
                    t_after_norm = True
                T = mul(T, t)
            if isinstance(layer, BasisLayer):
                out = layer.eval_basis()
                if layer.effect == "opd":
                    opd = add(opd, out)
                elif layer.effect == "phase":
                    phase = add(phase, out)
                else:
                    if normalise:
                        t_after_norm = True
                    T = mul(T, 1 + out)
            if isinstance(layer, AberratedLayer):
                opd = add(opd, layer.opd)
                phase = add(phase, layer.phase)
            if isinstance(layer, Normalise) or getattr(layer, "normalise", False):
                normalise = True
        if t_after_norm:
            return None
        return T, opd, phase, normalise

    def _geometry(self, wavelengths):
        """Per-wavelength device operands (wavenumber, scale_out, norm, wavelengths), cached:
        a fitting loop calls propagate with the same wavelength grid every step."""
        npix, ps, fl = self._focal_args()
        if torch.is_tensor(ps) or torch.is_tensor(wavelengths):
            # differentiable pixel scale / wavelengths: no cache, geometry through autograd
            ps_in = np.float32(self.diameter / np.float32(self.wf_npixels))
            wl_dev = (wavelengths.to(self.device, torch.float32) if torch.is_tensor(wavelengths)
                      else torch.as_tensor(wavelengths, device=self.device))
            ps = ps.to(self.device, torch.float32) if torch.is_tensor(ps) else ps
            scale_out, norm = _prop.mft_geometry(wl_dev, self.wf_npixels, ps_in, npix, ps, fl)
            k = np.float32(2 * math.pi) / wl_dev
            return npix, scale_out, norm, k, wl_dev
        key = (wavelengths.tobytes(), npix, float(ps), None if fl is None else float(fl),
               float(self.diameter), self.wf_npixels, str(self.device))
        cache = self.__dict__.setdefault("_geom_cache", {})
        hit = cache.get(key)
        if hit is None:
            ps_in = np.float32(self.diameter / np.float32(self.wf_npixels))
            scale_out, norm = _prop.mft_geometry(wavelengths, self.wf_npixels, ps_in, npix, ps, fl)
            k = (np.float32(2 * math.pi) / wavelengths).astype(np.float32)
            up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=self.device)
            hit = (npix, up(scale_out), up(norm), up(k), up(wavelengths))
            if len(cache) > 16:
                cache.clear()
            cache[key] = hit
        return hit

    def _upload(self, a):
        """Device copy of a small host array, cached by content: a fitting loop passes the same
        positions / weights every step, and a pageable host->device copy would serialise the
        host with the stream each time."""
        a = np.ascontiguousarray(_np32(a))
        key = (a.shape, a.tobytes())
        cache = self.__dict__.setdefault("_upload_cache", {})
        hit = cache.get(key)
        if hit is None:
            if len(cache) > 32:
                cache.clear()
            hit = cache[key] = torch.as_tensor(a, device=self.device)
        return hit

    def fused_propagate(self, wavelengths, offsets, weights):
        """psf = sum_{s,l} weights[s,l] |E_sl|^2 in one fused call.
        wavelengths [L] (host), offsets [S,2] rad (host or device), weights [S,L]."""
        parts = self._fusable()
        if parts is None:
            raise ValueError("layer stack is not pupil-only; use fused=False")
        T, opd, phase, normalise = parts
        dev = self.device
        if not (torch.is_tensor(wavelengths) and wavelengths.requires_grad):
            wavelengths = np.atleast_1d(_np32(wavelengths))
        npix, scale_out, norm, k, wl_dev = self._geometry(wavelengths)
        up = self._upload
        offsets_t = offsets.to(dev, torch.float32) if torch.is_tensor(offsets) else up(offsets)
        offsets_t = offsets_t.reshape(-1, 2)
        # tilt (wavefronts.py:370-395) folded into the output coordinates:
        # delta = theta * D / lambda, in fringes
        delta = (offsets_t[:, None, :] * self.diameter) / wl_dev[None, :, None]
        weights_t = weights.to(dev, torch.float32) if torch.is_tensor(weights) else up(weights)
        weights_t = weights_t.reshape(offsets_t.shape[0], len(wavelengths))
        cont = lambda t: None if t is None else t.contiguous()
        prec = (self.precision, True) if self.sparse else self.precision
        return ops.PolyPSFFunction.apply(cont(opd), cont(phase), weights_t.contiguous(), delta.contiguous(),
                                         cont(T), k, scale_out, norm, self.wf_npixels,
                                         npix, normalise, prec)


    def propagate_batch(self, coefficients, wavelengths, offset=None, weights=None, layer=None):
        """A batch of parameter sets through one fused call: ``coefficients`` [B, nz...] are B settings of the
        coefficients of one OPD basis layer of the stack (the only ``BasisLayer`` with effect "opd", or the
        one named by ``layer``); returns the B polychromatic PSFs [B, M, M].  Gradients flow to
        ``coefficients`` item by item.  This is the reference's ``vmap`` over parameter sets of
        ``OpticalSystem.propagate`` (optical_systems.py:147-223, docs/mask_design.md:454-488)."""
        cands = [(k, l) for k, l in self.layers.items() if isinstance(l, BasisLayer) and l.effect == "opd"]
        if layer is not None:
            cands = [(k, l) for k, l in cands if k == layer or l is layer]
        if len(cands) != 1:
            raise ValueError("propagate_batch needs exactly one OPD basis layer (name it with layer=...)")
        blayer = cands[0][1]
        if not self._can_fuse():
            raise ValueError("layer stack is not pupil-only; propagate the batch in a loop with fused=False")
        dev = self.device
        coefficients = coefficients if torch.is_tensor(coefficients) else torch.as_tensor(
            np.asarray(coefficients, dtype=np.float32), device=dev)
        if tuple(coefficients.shape[1:]) != tuple(blayer.coefficients.shape):
            raise ValueError("coefficients must be [B, *layer.coefficients.shape]")
        saved = blayer.coefficients
        try:                      # the other layers' (T, opd, phase): this layer contributes a zero OPD
            blayer.coefficients = torch.zeros_like(saved)
            T, opd, phase, normalise = self._fusable()
        finally:
            blayer.coefficients = saved
        wavelengths = np.atleast_1d(_np32(wavelengths))
        L = len(wavelengths)
        weights = np.full(L, 1.0 / L, np.float32) if weights is None else np.atleast_1d(_np32(weights))
        if weights.shape != wavelengths.shape:
            raise ValueError("Wavelength and weight shape mismatch")
        npix, scale_out, norm, k, wl_dev = self._geometry(wavelengths)
        delta = None
        if offset is not None:
            off = self._upload(np.asarray(_np32(offset)).reshape(1, 2))
            delta = ((off[:, None, :] * self.diameter) / wl_dev[None, :, None]).reshape(L, 2).contiguous()
        cont = lambda t: None if t is None else t.contiguous()
        return ops.PolyPSFBatchFunction.apply(coefficients, blayer.basis, cont(opd), cont(phase), cont(T),
                                              self._upload(weights), delta, k, scale_out, norm, self.wf_npixels,
                                              npix, normalise, self.precision)

    def _propagate(self, wavelengths, offset, weights, return_wf):
        if self.fused and not return_wf and self._can_fuse():
            off = offset if torch.is_tensor(offset) else _np32(offset)
            return self.fused_propagate(wavelengths, off.reshape(1, 2), weights.reshape(1, -1))
        return super()._propagate(wavelengths, offset, weights, return_wf)


class AngularOpticalSystem(ParametricLayeredOpticalSystem):
    """optical_systems.py:597-680: psf_pixel_scale in arcseconds."""

    def _focal_args(self):                              # :676-680
        if torch.is_tensor(self.psf_pixel_scale):
            p = self.psf_pixel_scale.to(self.device, torch.float32)
            return (self.psf_npixels * self.oversample, _prop.arcsec2rad(p / np.float32(self.oversample)), None)
        true_pixel_scale = np.float32(self.psf_pixel_scale / np.float32(self.oversample))
        return (self.psf_npixels * self.oversample, np.float32(_prop.arcsec2rad(true_pixel_scale)),
                None)


class CartesianOpticalSystem(ParametricLayeredOpticalSystem):
    """optical_systems.py:683-775: psf_pixel_scale in microns.  NOTE: the
    reference's ``to_focus`` (:771-775) does not forward ``focal_length`` to
    ``wavefront.propagate``; that behaviour is preserved, not fixed."""

    def __init__(self, wf_npixels, diameter, layers, focal_length, psf_npixels, psf_pixel_scale,
                 oversample=1, **kw):
        super().__init__(wf_npixels, diameter, layers, psf_npixels, psf_pixel_scale, oversample, **kw)
        self.focal_length = np.float32(focal_length)

    def _focal_args(self):                              # :771-775
        if torch.is_tensor(self.psf_pixel_scale):
            p = self.psf_pixel_scale.to(self.device, torch.float32)
            return (self.psf_npixels * self.oversample, p * np.float32(1e-6 / self.oversample), None)
        true_pixel_scale = np.float32(self.psf_pixel_scale / np.float32(self.oversample))
        return (self.psf_npixels * self.oversample, np.float32(1e-6 * true_pixel_scale), None)


_FocalSystem = ParametricLayeredOpticalSystem     # former private name
