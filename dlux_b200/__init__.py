"""dlux_b200: B200-native (sm_100a) implementation of dLux's diffraction hot path --
the matrix-Fourier-transform pupil->focal propagation behind ``dlu.MFT``, the ``MFT``
propagator layer, ``Wavefront.propagate`` and ``OpticalSystem.propagate/model``.

Importing the package does not need a GPU; calling any operator does, and there is no
CPU fallback (the native library is loaded lazily and its absence is an error)."""
from . import utils
from .apertures import (AberratedAperture, CircularAperture, CompoundAperture, CoordTransform, MultiAperture,
                        RectangularAperture, RegPolyAperture, Spider, SquareAperture)
from .psfs import PSF
from .detectors import (AddConstant, ApplyInterpolation, ApplyJitter, ApplyPixelResponse, ApplySaturation, DetectorLayer,
                        Downsample, LayeredDetector, Telescope)
from .layers import (AberratedLayer, BasisLayer, BasisOptic, FFT, Flip, Lambda, MFT, Normalise, Optic, OpticalLayer,
                     Resize, Rotate, Tilt, TransmissiveLayer, UnifiedLayer)
from .optical_systems import (AngularOpticalSystem, BaseOpticalSystem, CartesianOpticalSystem,
                              LayeredOpticalSystem, OpticalSystem, ParametricLayeredOpticalSystem,
                              ParametricOpticalSystem)
from .sources import (BinarySource, PointResolvedSource, PointSource, PointSources, ResolvedSource,
                      Scene)
from .wavefronts import CoordSpec, Wavefront
from .graphs import GraphedFitStep, GraphedValueAndGrad

__version__ = "0.1.0"
__all__ = ["utils", "Wavefront", "OpticalLayer", "TransmissiveLayer", "AberratedLayer", "BasisLayer",
           "Tilt", "Normalise", "Optic", "BasisOptic", "MFT", "FFT", "CoordSpec", "BaseOpticalSystem",
           "OpticalSystem", "ParametricOpticalSystem",
           "LayeredOpticalSystem", "ParametricLayeredOpticalSystem", "AngularOpticalSystem", "CartesianOpticalSystem", "PointSource",
           "PointSources", "BinarySource", "ResolvedSource", "PointResolvedSource", "Scene", "CoordTransform", "AberratedAperture", "CircularAperture", "SquareAperture", "RectangularAperture",
           "RegPolyAperture", "Spider", "CompoundAperture", "MultiAperture", "PSF", "DetectorLayer",
           "ApplyInterpolation", "ApplyPixelResponse", "ApplyJitter", "ApplySaturation", "AddConstant", "Downsample",
           "LayeredDetector", "Telescope", "GraphedValueAndGrad", "GraphedFitStep", "UnifiedLayer", "Resize", "Rotate", "Flip",
           "Lambda"]
