"""Mirror of the slice of ``dLux.utils`` that sits on the diffraction hot path."""
from . import geometry, interpolation, propagation, zernikes
from .array_ops import downsample
from .interpolation import interp, rotate, scale
from .propagation import (MFT, FFT, calc_nfringes, mft_geometry, arcsec2rad, eval_basis, fft_spec,
                          fft_phase_ramp)

__all__ = ["propagation", "geometry", "zernikes", "MFT", "FFT", "calc_nfringes", "mft_geometry", "arcsec2rad", "eval_basis",
           "fft_spec", "fft_phase_ramp", "downsample", "interpolation", "interp", "scale", "rotate"]
