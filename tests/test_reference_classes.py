"""Class-level pin: ``tests/golden/reference_classes.npz`` holds what the reference's OWN
``Wavefront`` / layers / ``OpticalSystem.propagate`` / ``*Source.model`` code returns (executed
unmodified on NumPy stand-ins for jax / zodiax / equinox by tests/golden/make_golden_classes.py),
plus float64 central-difference gradients through that executed code.

CPU tests here check the oracle (oracle/mft_oracle.py) and the autograd twin (oracle/torch_twin.py)
against it; the ``gpu`` tests check the CUDA path (dlux_b200 public API -> C ABI) against the same
vectors directly.  Tolerance 1e-5 relative L2 (north_star), 2e-7 for oracle-vs-reference (float32
summation order only)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2
from oracle import mft_oracle as O
from oracle import torch_twin

TOL = 1e-5


@pytest.fixture(scope="module")
def gc():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_classes.npz"))


def small(gc, **over):
    sc = gc["sm_scalars"]
    od = dict(wf_npixels=int(sc[0]), diameter=np.float32(sc[1]), psf_npixels=int(sc[2]),
              psf_pixel_scale=np.float32(sc[3]), oversample=int(sc[4]), transmission=gc["sm_transmission"],
              basis=gc["sm_basis"], coefficients=gc["sm_coefficients"], normalise=True)
    od.update(over)
    return od, np.float32(sc[5])


# ------------------------------------------------------------------ oracle vs executed reference (CPU)
def test_oracle_c1_bit_exact_against_reference_classes(gc):
    from dlux_b200 import workloads
    c1 = workloads.config("c1")
    psf = O.point_source_model(c1, c1["wavelengths"], np.zeros(2, np.float32), 1.0)
    field = O.propagate_mono(c1, c1["wavelengths"][0], None, True)
    assert np.array_equal(field, gc["c1_field"])          # every rounding of the class layer reproduced
    assert np.array_equal(psf, gc["c1_psf"])
    assert np.float32(gc["c1_pixel_scale"]) == O.arcsec2rad(np.float32(c1["psf_pixel_scale"]))


def test_oracle_c2_offset_polychromatic(gc):
    from dlux_b200 import workloads
    c2 = workloads.config("c2")
    psf = O.point_source_model(c2, c2["wavelengths"], gc["c2_position"], 1.0, c2["weights"])
    assert rel_l2(psf, gc["c2_psf"]) < 2e-7


def test_oracle_sources_and_layers(gc):
    od, flux = small(gc)
    wl, w, pos = gc["sm_wavelengths"], gc["sm_weights"], gc["sm_position"]
    assert rel_l2(O.point_source_model(od, wl, pos, flux, w), gc["sm_point_psf"]) < 2e-7
    # per-wavelength complex fields (sqrt(weight) applied): pins the tilt sign / axis convention
    wn = w / w.sum()
    fields = O.propagate(od, wl, pos, wn * flux, return_field=True)
    assert rel_l2(fields, gc["sm_point_fields"]) < 2e-7
    assert rel_l2(O.propagate(od, wl), gc["sm_propagate_default"]) < 2e-7
    stars = O.point_sources_model(od, wl, gc["sm_positions"], gc["sm_fluxes"], w)
    assert rel_l2(stars, gc["sm_stars_psf"]) < 2e-7
    # Optic(transmission, opd, phase, normalise)
    od2, _ = small(gc, basis=None, coefficients=None, opd=gc["sm_opd"], phase=gc["sm_phase"])
    assert rel_l2(O.point_source_model(od2, wl, pos, flux, w), gc["sm_optic_psf"]) < 2e-7
    # Scene = PointSource + PointSources
    assert rel_l2(O.point_source_model(od, wl, pos, flux, w) + stars, gc["sm_scene_psf"]) < 2e-7


def test_oracle_cartesian_quirk(gc):
    # CartesianOpticalSystem.to_focus does NOT pass its focal length (optical_systems.py:771-775):
    # the MFT runs in angular units with pixel_scale = 1e-6 * psf_pixel_scale / oversample
    od, flux = small(gc)
    wl, w, pos = gc["sm_wavelengths"], gc["sm_weights"], gc["sm_position"]
    wn = (w / w.sum()) * flux
    M = od["psf_npixels"] * 2
    ps = np.float32(1e-6) * (np.float32(0.4) / np.float32(2))        # weak Python float * f32 -> f32
    psf = 0
    for l in range(len(wl)):
        wf = O.OracleWavefront(wl[l], od["wf_npixels"], od["diameter"])
        wf.tilt(pos)
        O.apply_optic(wf, od["transmission"], None, None, od["basis"], od["coefficients"], True)
        wf.propagate(M, ps)
        psf = psf + wn[l] * wf.psf
    assert rel_l2(psf, gc["sm_cartesian_psf"]) < 2e-7


def test_twin_gradients_match_finite_differences_of_the_executed_reference(gc):
    """The autograd twin (float64) against central differences through the reference's own classes run
    in x64 mode: pins the gradient path (what jax.grad differentiates) for coefficients, position,
    flux and spectral weights (through the spectrum normalisation, spectra.py:113-117)."""
    od, flux = small(gc)
    wl = gc["sm_wavelengths"]
    G = torch.tensor(gc["sm_G"], dtype=torch.float64)
    c = torch.tensor(gc["sm_coefficients"], dtype=torch.float64, requires_grad=True)
    p = torch.tensor(gc["sm_position"], dtype=torch.float64, requires_grad=True)
    f = torch.tensor(float(flux), dtype=torch.float64, requires_grad=True)
    w = torch.tensor(gc["sm_weights"], dtype=torch.float64, requires_grad=True)
    sc = gc["sm_scalars"]          # the x64 run took diameter / pixel scale as Python floats (float64)
    psf = torch_twin.poly_psf(od["transmission"], None, wl, (w / w.sum()) * f, diameter=float(sc[1]),
                              psf_npixels=od["psf_npixels"],
                              pixel_scale_rad=O.arcsec2rad(float(sc[3]), np.float64), offset=p,
                              basis=od["basis"], coefficients=c, dtype=np.float64)
    assert rel_l2(psf.detach().numpy(), gc["sm_x64_psf"]) < 1e-12
    loss = (psf * G).sum()
    assert abs(loss.item() - float(gc["sm_x64_loss"])) < 1e-12 * abs(float(gc["sm_x64_loss"])) + 1e-15
    loss.backward()
    assert rel_l2(c.grad.numpy(), gc["sm_fd_grad_coefficients"]) < 1e-6
    assert rel_l2(p.grad.numpy(), gc["sm_fd_grad_position"]) < 1e-6
    assert abs(f.grad.item() - float(gc["sm_fd_grad_flux"])) < 1e-7 * abs(float(gc["sm_fd_grad_flux"]))
    assert rel_l2(w.grad.numpy(), gc["sm_fd_grad_weights"]) < 1e-6


# ------------------------------------------------------------------ CUDA path vs executed reference (GPU)
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dlux_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _angular(dl, od, dev, fused=True, coeffs=None, precision=None):
    co = od["coefficients"] if coeffs is None else coeffs
    layer = dl.BasisOptic(od["basis"], od["transmission"], co, normalise=True, effect="opd", device=dev)
    return dl.AngularOpticalSystem(od["wf_npixels"], od["diameter"], [("pupil", layer)], od["psf_npixels"],
                                   od["psf_pixel_scale"], od["oversample"], device=dev, fused=fused,
                                   precision=precision)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_cuda_c1_against_reference_classes(dev, gc, fused):
    import dlux_b200 as dl
    from dlux_b200 import workloads
    c1 = workloads.config("c1")
    optics = _angular(dl, c1, dev, fused)
    psf = optics.model(dl.PointSource(c1["wavelengths"], np.zeros(2, np.float32), 1.0))
    assert rel_l2(psf.cpu().numpy(), gc["c1_psf"]) < TOL
    wf = optics.propagate_mono(c1["wavelengths"][0], return_wf=True)
    assert rel_l2(wf.phasor.cpu().numpy(), gc["c1_field"]) < TOL


@pytest.mark.gpu
def test_cuda_c2_against_reference_classes(dev, gc):
    import dlux_b200 as dl
    from dlux_b200 import workloads
    c2 = workloads.config("c2")
    optics = _angular(dl, c2, dev)
    psf = optics.model(dl.PointSource(c2["wavelengths"], gc["c2_position"], 1.0, weights=c2["weights"]))
    assert rel_l2(psf.cpu().numpy(), gc["c2_psf"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_cuda_sources_against_reference_classes(dev, gc, fused):
    import dlux_b200 as dl
    od, flux = small(gc)
    wl, w, pos = gc["sm_wavelengths"], gc["sm_weights"], gc["sm_position"]
    optics = _angular(dl, od, dev, fused)
    src = dl.PointSource(wl, pos, flux, weights=w)
    assert rel_l2(optics.model(src).cpu().numpy(), gc["sm_point_psf"]) < TOL
    assert rel_l2(optics.propagate(wl).cpu().numpy(), gc["sm_propagate_default"]) < TOL
    stars = dl.PointSources(wl, gc["sm_positions"], gc["sm_fluxes"], weights=w)
    assert rel_l2(optics.model(stars).cpu().numpy(), gc["sm_stars_psf"]) < TOL
    binary = dl.BinarySource(wl, pos, 2.0, 4.0e-7, 0.7, 3.0, weights=gc["sm_w2"])
    assert rel_l2(optics.model(binary).cpu().numpy(), gc["sm_binary_psf"]) < TOL
    resolved = dl.ResolvedSource(wl, pos, 1.7, gc["sm_distribution"], weights=w)
    assert rel_l2(optics.model(resolved).cpu().numpy(), gc["sm_resolved_psf"]) < TOL
    pres = dl.PointResolvedSource(wl, pos, 1.7, gc["sm_distribution"], 5.0, weights=gc["sm_w2"])
    assert rel_l2(optics.model(pres).cpu().numpy(), gc["sm_point_resolved_psf"]) < TOL
    scene = dl.Scene([("a", src), ("b", stars)])
    assert rel_l2(optics.model(scene).cpu().numpy(), gc["sm_scene_psf"]) < TOL
    # the Wavefront / PSF return modes of the reference
    wfs = optics.model(src, return_wf=True)
    assert rel_l2(wfs.phasor.cpu().numpy(), gc["sm_point_fields"]) < TOL
    pobj = optics.model(src, return_psf=True)
    assert rel_l2(pobj.data.cpu().numpy(), gc["sm_point_psf"]) < TOL
    assert abs(float(pobj.pixel_scale) - float(gc["sm_point_psf_pixel_scale"])) < 1e-6 * float(gc["sm_point_psf_pixel_scale"])
    wf_stars = optics.model(stars, return_wf=True)
    assert rel_l2(wf_stars.phasor.cpu().numpy(), gc["sm_stars_fields"]) < TOL


@pytest.mark.gpu
def test_cuda_layer_stacks_against_reference_classes(dev, gc):
    import dlux_b200 as dl
    od, flux = small(gc)
    wl, w, pos = gc["sm_wavelengths"], gc["sm_weights"], gc["sm_position"]
    src = dl.PointSource(wl, pos, flux, weights=w)
    optic = dl.Optic(od["transmission"], gc["sm_opd"], gc["sm_phase"], normalise=True, device=dev)
    sys_ = dl.AngularOpticalSystem(od["wf_npixels"], od["diameter"], [("pupil", optic)], od["psf_npixels"],
                                   od["psf_pixel_scale"], od["oversample"], device=dev)
    assert rel_l2(sys_.model(src).cpu().numpy(), gc["sm_optic_psf"]) < TOL
    lay = dl.LayeredOpticalSystem(od["wf_npixels"], od["diameter"], [
        ("t", dl.TransmissiveLayer(od["transmission"], device=dev)),
        ("a", dl.AberratedLayer(gc["sm_opd"], gc["sm_phase"], device=dev)),
        ("tilt", dl.Tilt(gc["sm_tilt_angles"])),
        ("n", dl.Normalise()),
        ("mft", dl.MFT(40, O.arcsec2rad(np.float32(0.07)))),
    ], device=dev)
    assert rel_l2(lay.propagate(wl, pos, w).cpu().numpy(), gc["sm_layered_psf"]) < TOL
    amp = dl.BasisOptic(od["basis"] * np.float32(1e7), od["transmission"], od["coefficients"], effect="amplitude",
                        normalise=True, device=dev)
    sys_ = dl.AngularOpticalSystem(od["wf_npixels"], od["diameter"], [("pupil", amp)], od["psf_npixels"],
                                   od["psf_pixel_scale"], od["oversample"], device=dev)
    assert rel_l2(sys_.model(src).cpu().numpy(), gc["sm_amplitude_psf"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_cuda_cartesian_system_against_reference_classes(dev, gc, fused):
    # quirk F9 (optical_systems.py:771-775) preserved
    import dlux_b200 as dl
    od, flux = small(gc)
    layer = dl.BasisOptic(od["basis"], od["transmission"], od["coefficients"], normalise=True, effect="opd", device=dev)
    cart = dl.CartesianOpticalSystem(od["wf_npixels"], od["diameter"], [("pupil", layer)], 2.5,
                                     od["psf_npixels"], 0.4, 2, device=dev, fused=fused)
    assert isinstance(cart, dl.ParametricLayeredOpticalSystem)
    src = dl.PointSource(gc["sm_wavelengths"], gc["sm_position"], flux, weights=gc["sm_weights"])
    assert rel_l2(cart.model(src).cpu().numpy(), gc["sm_cartesian_psf"]) < TOL


@pytest.mark.gpu
def test_cuda_gradients_against_reference_finite_differences(dev, gc):
    """CUDA gradients (float32 inputs) against the float64 central differences through the executed
    reference.  The reference's own float32 forward differs from its float64 forward by 6.8e-7 on this
    case, so the bar stays at 1e-5."""
    import dlux_b200 as dl
    od, flux = small(gc)
    c = torch.as_tensor(od["coefficients"], device=dev).requires_grad_(True)
    p = torch.as_tensor(gc["sm_position"], device=dev).requires_grad_(True)
    f = torch.tensor(float(flux), device=dev, requires_grad=True)
    optics = _angular(dl, od, dev, True, coeffs=c)
    psf = optics.model(dl.PointSource(gc["sm_wavelengths"], p, f, weights=gc["sm_weights"]))
    (psf * torch.as_tensor(gc["sm_G"], device=dev)).sum().backward()
    assert rel_l2(psf.detach().cpu().numpy(), gc["sm_x64_psf"]) < TOL
    assert rel_l2(c.grad.cpu().numpy(), gc["sm_fd_grad_coefficients"]) < TOL
    assert rel_l2(p.grad.cpu().numpy(), gc["sm_fd_grad_position"]) < TOL
    assert abs(f.grad.item() - float(gc["sm_fd_grad_flux"])) < TOL * abs(float(gc["sm_fd_grad_flux"]))
