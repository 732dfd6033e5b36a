"""Mirror of the slice of ``dLux.utils`` that sits on the diffraction hot path."""
from . import propagation
from .propagation import MFT, calc_nfringes, mft_geometry, arcsec2rad, eval_basis

__all__ = ["propagation", "MFT", "calc_nfringes", "mft_geometry", "arcsec2rad", "eval_basis"]
