"""GPU parity tests: the CUDA path (through the C ABI, via dlux_b200.ops) against the
oracle, the committed golden vectors and size-independent properties.

Tolerance (BASELINE.json north_star): relative L2 <= 1e-5 on fields, PSFs and
gradients, against the reference's complex64 arithmetic."""
import numpy as np
import pytest
import torch

from conftest import check, rel_l2, rel_scalar
from oracle import mft_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5
PRECS = ["fp32", "3xtf32"]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dlux_b200 import _lib
    _lib.load()          # raises loudly if the native library is missing
    return torch.device("cuda:0")


def _geom(g):
    n_in, n_out, wl, psi, pso, fl, sx, sy, pixel, inverse = g
    return dict(n_in=int(n_in), n_out=int(n_out), wl=wl, psi=psi, pso=pso,
                fl=None if fl < 0 else fl, shift=(sx, sy), pixel=bool(pixel), inverse=bool(inverse))


def _rand_c64(rng, *shape):
    return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / shape[-1]).astype(np.complex64)


# ------------------------------------------------------------------ bit-exact pieces
def test_coords_bit_exact(dev):
    from dlux_b200 import ops
    cases = [(32, 16, 0.25, 1.0), (96, 40, 0.0817, -2.0), (256, 128, 0.24240684, 0.0),
             (1024, 512, 0.12207031, 0.5), (100, 37, 1.7, 3.25)]
    for n_in, n_out, s, shift in cases:
        sc = torch.tensor([s, s * 1.1], dtype=torch.float32, device=dev)
        sh = torch.tensor([[shift, -shift], [0.0, 2 * shift]], dtype=torch.float32, device=dev)
        xin, uout = ops.mft_coords(n_in, n_out, sc, sh)
        xin, uout = xin.cpu().numpy(), uout.cpu().numpy()
        for b in range(2):
            for ax in range(2):
                sft = np.float32(sh[b, ax].item())
                sv = np.float32(sc[b].item())
                ref_in = O.nd_coords_1d(n_in, 1.0 / n_in, sft * np.float32(1.0 / n_in))
                ref_out = O.nd_coords_1d(n_out, sv, sft * sv)
                assert np.array_equal(xin[b, ax], ref_in), (n_in, b, ax)
                assert np.array_equal(uout[b, ax], ref_out), (n_out, b, ax)


# ------------------------------------------------------------------ MFT vs reference source
@pytest.mark.parametrize("prec", PRECS)
def test_mft_golden_reference_vectors(dev, golden, prec):
    import dlux_b200 as dl
    for g in range(int(golden["n_geoms"])):
        p = _geom(golden[f"mft_{g}_geom"])
        ph = torch.as_tensor(golden[f"mft_{g}_in"], device=dev)
        out = dl.utils.MFT(ph, np.float32(p["wl"]), np.float32(p["psi"]), p["n_out"], np.float32(p["pso"]),
                           None if p["fl"] is None else np.float32(p["fl"]),
                           np.asarray(p["shift"], np.float32), p["pixel"], p["inverse"], precision=prec)
        err = rel_l2(out.cpu().numpy(), golden[f"mft_{g}_out"])
        assert err < TOL, (g, prec, err)


@pytest.mark.parametrize("prec", PRECS)
def test_mft_reference_test_fixture(dev, golden, prec):
    # /root/reference/tests/utils/test_propagation.py:98-139 (+ values)
    import dlux_b200 as dl
    ones32 = torch.ones((32, 32), dtype=torch.complex64, device=dev)
    k = 0
    for fl in (None, np.float32(2.0)):
        for inverse in (False, True):
            for pixel in (True, False):
                out = dl.utils.MFT(ones32, np.float32(1.0), np.float32(0.1), 16, np.float32(0.05), fl,
                                   np.asarray([1.0, -2.0], np.float32), pixel, inverse, precision=prec)
                assert out.shape == (16, 16)
                assert not torch.isnan(out.real).any()
                assert rel_l2(out.cpu().numpy(), golden[f"reftest_{k}"]) < TOL, k
                k += 1


@pytest.mark.parametrize("prec", PRECS)
def test_mft_shift_units_identity(dev, prec):
    # /root/reference/tests/utils/test_propagation.py:141-176
    import dlux_b200 as dl
    ph = torch.ones((32, 32), dtype=torch.complex64, device=dev)
    a = dl.utils.MFT(ph, 1.0, 0.1, 16, 0.05, 2.0, np.asarray([1.0, -2.0], np.float32), True, precision=prec)
    b = dl.utils.MFT(ph, 1.0, 0.1, 16, 0.05, 2.0, np.asarray([0.05, -0.1], np.float32), False, precision=prec)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n_in,n_out,nl", [(256, 128, 1), (512, 256, 3), (200, 72, 2), (130, 131, 1)])
def test_mft_vs_oracle_batched(dev, prec, n_in, n_out, nl):
    import dlux_b200 as dl
    rng = np.random.default_rng(n_in + n_out)
    ph = _rand_c64(rng, nl, n_in, n_in)
    wls = np.linspace(0.9e-6, 1.1e-6, nl).astype(np.float32)
    ps_in = np.float32(1.0 / n_in)
    pso = O.arcsec2rad(0.05)
    out = dl.utils.MFT(torch.as_tensor(ph, device=dev), wls, ps_in, n_out, pso, precision=prec)
    assert out.shape == (nl, n_out, n_out)
    for l in range(nl):
        ref = O.MFT(ph[l], wls[l], ps_in, n_out, pso)
        err = rel_l2(out[l].cpu().numpy(), ref)
        assert err < TOL, (l, err)


@pytest.mark.parametrize("prec", PRECS)
def test_adjoint_is_conjugate_transpose(dev, prec):
    from dlux_b200 import ops
    rng = np.random.default_rng(7)
    n_in, n_out = 96, 40
    x = torch.as_tensor(_rand_c64(rng, 2, n_in, n_in), device=dev)
    y = torch.as_tensor(_rand_c64(rng, 2, n_out, n_out), device=dev)
    s = torch.tensor([0.21, 0.33], device=dev)
    sh = torch.tensor([[0.5, -1.0], [0.0, 2.0]], device=dev)
    nrm = torch.tensor([0.7, 1.3], device=dev)
    Ax = ops.mft_c64(x, s, n_out, sh, None, nrm, False, False, prec)
    Aty = ops.mft_c64(y, s, n_in, sh, None, nrm, False, True, prec)
    lhs = torch.sum(torch.conj(y) * Ax, dim=(1, 2)).cpu().numpy()
    rhs = torch.sum(torch.conj(Aty) * x, dim=(1, 2)).cpu().numpy()
    assert np.allclose(lhs, rhs, rtol=2e-5, atol=1e-8), (lhs, rhs)
    # against the oracle's matrices
    xin, uout = ops.mft_coords(n_in, n_out, s, sh)
    xin, uout = xin.cpu().numpy().astype(np.float64), uout.cpu().numpy().astype(np.float64)
    for b in range(2):
        Ay = np.exp(-2j * np.pi * np.outer(xin[b, 1], uout[b, 1]))
        Axm = np.exp(-2j * np.pi * np.outer(xin[b, 0], uout[b, 0]))
        ref = float(nrm[b]) * (np.conj(Ay) @ y[b].cpu().numpy().astype(np.complex128) @ np.conj(Axm).T)
        assert rel_l2(Aty[b].cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("prec", PRECS)
def test_mft_autograd_matches_adjoint(dev, prec):
    import dlux_b200 as dl
    rng = np.random.default_rng(8)
    ph = torch.as_tensor(_rand_c64(rng, 64, 64), device=dev).requires_grad_(True)
    out = dl.utils.MFT(ph, 1e-6, 1.0 / 64, 32, 1.5e-7, precision=prec)
    g = torch.as_tensor(_rand_c64(rng, 32, 32), device=dev)
    loss = (out.real * g.real + out.imag * g.imag).sum()
    loss.backward()
    tm = O.transfer_matrix(1e-6, 64, 1.0 / 64, 32, 1.5e-7, dtype=np.float64)
    nrm = O.mft_norm(O.calc_nfringes(1e-6, 64, 1.0 / 64, 32, 1.5e-7, dtype=np.float64), 64, 32, np.float64)
    ref = nrm * (np.conj(tm) @ g.cpu().numpy().astype(np.complex128) @ np.conj(tm).T)
    assert rel_l2(ph.grad.cpu().numpy(), ref) < TOL


# ------------------------------------------------------------------ fused poly-PSF
def _optics_dict(N, M, nz, seed, oversample=1, pscale=0.05):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[:N, :N]
    r = np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) / (N / 2)
    T = (r <= 1).astype(np.float32)
    basis = (rng.standard_normal((nz, N, N)).astype(np.float32) * T) * np.float32(2e-8)
    coeffs = rng.standard_normal(nz).astype(np.float32)
    return dict(wf_npixels=N, diameter=1.0, psf_npixels=M, psf_pixel_scale=pscale,
                oversample=oversample, transmission=T, basis=basis, coefficients=coeffs,
                normalise=True)


def _system(od, dev, fused=True, prec=None):
    import dlux_b200 as dl
    layer = dl.BasisOptic(od["basis"], od["transmission"], od["coefficients"], normalise=True, effect="opd",
                          device=dev)
    return dl.AngularOpticalSystem(od["wf_npixels"], od["diameter"], [("aperture", layer)],
                                   od["psf_npixels"], od["psf_pixel_scale"], od["oversample"],
                                   device=dev, fused=fused, precision=prec)


@pytest.mark.parametrize("prec", PRECS)
def test_config1_angular_system_psf(dev, prec):
    # BASELINE config 1: 256 px circular aperture + 10-term basis OPD, 1 source, 1 wavelength,
    # MFT to 128x128
    od = _optics_dict(256, 128, 10, 0)
    sys_ = _system(od, dev, True, prec)
    psf = sys_.propagate(np.array([1.0e-6], np.float32))
    ref = O.propagate(od, [1.0e-6])
    assert psf.shape == (128, 128)
    assert rel_l2(psf.cpu().numpy(), ref) < TOL
    # the layer-by-layer route gives the same numbers
    psf2 = _system(od, dev, False, prec).propagate(np.array([1.0e-6], np.float32))
    assert rel_l2(psf2.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("prec", PRECS)
def test_polychromatic_offset_weights(dev, prec):
    od = _optics_dict(128, 64, 5, 1, oversample=2, pscale=0.08)
    sys_ = _system(od, dev, True, prec)
    wls = np.linspace(0.9e-6, 1.1e-6, 5).astype(np.float32)
    w = np.array([0.1, 0.3, 0.2, 0.25, 0.15], np.float32)
    off = np.array([3.0e-7, -2.0e-7], np.float32)
    psf = sys_.propagate(wls, off, w)
    ref = O.propagate(od, wls, off, w)
    assert psf.shape == (128, 128)
    assert rel_l2(psf.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("prec", PRECS)
def test_point_sources_model(dev, prec):
    import dlux_b200 as dl
    od = _optics_dict(96, 48, 4, 2)
    sys_ = _system(od, dev, True, prec)
    wls = np.linspace(0.95e-6, 1.05e-6, 3).astype(np.float32)
    pos = np.array([[1e-7, -2e-7], [-3e-7, 0.5e-7], [0.0, 0.0]], np.float32)
    flux = np.array([3.0, 0.25, 1.0], np.float32)
    src = dl.PointSources(wls, pos, flux)
    psf = sys_.model(src)
    ref = O.point_sources_model(od, wls, pos, flux)
    assert rel_l2(psf.cpu().numpy(), ref) < TOL
    one = dl.PointSource(wls, pos[0], 3.0)
    assert rel_l2(sys_.model(one).cpu().numpy(), O.point_source_model(od, wls, pos[0], 3.0)) < TOL


@pytest.mark.parametrize("prec", PRECS)
def test_phase_retrieval_gradient(dev, prec):
    # BASELINE config 2 (scaled down): grad of a PSF loss w.r.t. the basis coefficients
    import dlux_b200 as dl
    from oracle import torch_twin
    N, M, nz = 128, 64, 6
    od = _optics_dict(N, M, nz, 3)
    wls = np.linspace(0.9e-6, 1.1e-6, 4).astype(np.float32)
    w = np.full(4, 0.25, np.float32)
    off = np.array([2.0e-7, 1.0e-7], np.float32)
    rng = np.random.default_rng(30)
    G = rng.standard_normal((M, M)).astype(np.float32)

    coeffs = torch.as_tensor(od["coefficients"], device=dev).requires_grad_(True)
    wt = torch.as_tensor(w, device=dev).requires_grad_(True)
    layer = dl.BasisOptic(od["basis"], od["transmission"], coeffs, normalise=True, effect="opd", device=dev)
    sys_ = dl.AngularOpticalSystem(N, 1.0, [("aperture", layer)], M, 0.05, device=dev, precision=prec)
    psf = sys_.propagate(wls, off, wt)
    loss = (psf * torch.as_tensor(G, device=dev)).sum()
    loss.backward()

    c_ref = torch.tensor(od["coefficients"], dtype=torch.float64, requires_grad=True)
    w_ref = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    psf_ref = torch_twin.poly_psf(od["transmission"], None, wls, w_ref, diameter=1.0, psf_npixels=M,
                                  pixel_scale_rad=O.arcsec2rad(0.05), offset=off, basis=od["basis"],
                                  coefficients=c_ref, dtype=np.float64)
    (psf_ref * torch.tensor(G, dtype=torch.float64)).sum().backward()
    assert rel_l2(psf.detach().cpu().numpy(), psf_ref.detach().numpy()) < TOL
    assert rel_l2(coeffs.grad.cpu().numpy(), c_ref.grad.numpy()) < TOL
    assert rel_l2(wt.grad.cpu().numpy(), w_ref.grad.numpy()) < TOL


@pytest.mark.parametrize("prec", PRECS)
def test_point_sources_position_and_flux_gradients(dev, prec):
    # PointSources.position / flux gradients (north_star: "PointSources position/flux sum")
    import dlux_b200 as dl
    from oracle import torch_twin
    N, M, nz = 96, 48, 3
    od = _optics_dict(N, M, nz, 6)
    wls = np.linspace(0.95e-6, 1.05e-6, 3).astype(np.float32)
    w = np.full(3, 1 / 3, np.float32)
    pos0 = np.array([[1.5e-7, -2.0e-7], [-3.0e-7, 0.5e-7]], np.float32)
    flux0 = np.array([2.0, 0.75], np.float32)
    G = np.random.default_rng(31).standard_normal((M, M)).astype(np.float32)

    pos = torch.as_tensor(pos0, device=dev).requires_grad_(True)
    flux = torch.as_tensor(flux0, device=dev).requires_grad_(True)
    sys_ = _system(od, dev, True, prec)
    psf = sys_.model(dl.PointSources(wls, pos, flux))
    (psf * torch.as_tensor(G, device=dev)).sum().backward()

    pos_r = torch.tensor(pos0, dtype=torch.float64, requires_grad=True)
    flux_r = torch.tensor(flux0, dtype=torch.float64, requires_grad=True)
    tot = 0.0
    for s in range(2):
        tot = tot + torch_twin.poly_psf(od["transmission"], None, wls, torch.tensor(w, dtype=torch.float64) * flux_r[s],
                                        diameter=1.0, psf_npixels=M, pixel_scale_rad=O.arcsec2rad(0.05),
                                        offset=pos_r[s], basis=od["basis"], coefficients=od["coefficients"],
                                        dtype=np.float64)
    (tot * torch.tensor(G, dtype=torch.float64)).sum().backward()
    assert rel_l2(psf.detach().cpu().numpy(), tot.detach().numpy()) < TOL
    assert rel_l2(flux.grad.cpu().numpy(), flux_r.grad.numpy()) < TOL
    check(f"position grad [{prec}]", rel_l2(pos.grad.cpu().numpy(), pos_r.grad.numpy()), TOL)


@pytest.mark.parametrize("prec", PRECS)
def test_transmission_and_phase_gradients(dev, prec):
    # gradients w.r.t. the remaining pupil-plane leaves: transmission (through the power
    # normalisation too) and an additive phase screen
    import dlux_b200 as dl
    from oracle import torch_twin
    N, M = 64, 32
    rng = np.random.default_rng(40)
    yy, xx = np.mgrid[:N, :N]
    r = np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) / (N / 2)
    T0 = ((r <= 1) * rng.uniform(0.5, 1.0, (N, N))).astype(np.float32)
    opd0 = (rng.standard_normal((N, N)) * 3e-8).astype(np.float32)
    wls = np.linspace(0.9e-6, 1.1e-6, 3).astype(np.float32)
    w = np.array([0.2, 0.5, 0.3], np.float32)
    off = np.array([1.0e-7, -3.0e-7], np.float32)
    G = rng.standard_normal((M, M)).astype(np.float32)
    for normalise in (True, False):
        T = torch.as_tensor(T0, device=dev).requires_grad_(True)
        opd = torch.as_tensor(opd0, device=dev).requires_grad_(True)
        layer = dl.Optic(T, opd, None, normalise=normalise, device=dev)
        sys_ = dl.AngularOpticalSystem(N, 1.0, [("optic", layer)], M, 0.05, device=dev, precision=prec)
        psf = sys_.propagate(wls, off, w)
        (psf * torch.as_tensor(G, device=dev)).sum().backward()
        T_r = torch.tensor(T0, dtype=torch.float64, requires_grad=True)
        o_r = torch.tensor(opd0, dtype=torch.float64, requires_grad=True)
        ref = torch_twin.poly_psf(T_r, o_r, wls, w, diameter=1.0, psf_npixels=M,
                                  pixel_scale_rad=O.arcsec2rad(0.05), offset=off, normalise=normalise,
                                  dtype=np.float64)
        (ref * torch.tensor(G, dtype=torch.float64)).sum().backward()
        assert rel_l2(psf.detach().cpu().numpy(), ref.detach().numpy()) < TOL
        assert rel_l2(opd.grad.cpu().numpy(), o_r.grad.numpy()) < TOL, normalise
        check(f"transmission grad [{prec}, normalise={normalise}]", rel_l2(T.grad.cpu().numpy(), T_r.grad.numpy()), TOL)


def test_layered_system_with_mft_layers_batched_route(dev):
    # unfused route: a LayeredOpticalSystem whose stack holds its own MFT layers (pupil -> focal ->
    # pupil -> focal, a Lyot-style chain) runs all wavelengths as ONE batched wavefront (the
    # reference vmaps propagate_mono, optical_systems.py:213-216); compared with the per-wavelength
    # loop and, for the first leg, with the oracle
    import dlux_b200 as dl
    N, M = 96, 48
    od = _optics_dict(N, M, 3, 11)
    wls = np.linspace(0.9e-6, 1.1e-6, 4).astype(np.float32)
    w = np.array([0.1, 0.2, 0.3, 0.4], np.float32)
    off = np.array([1.0e-7, 2.0e-7], np.float32)
    ps = O.arcsec2rad(np.float32(0.05))
    mk_optic = lambda c: dl.BasisOptic(od["basis"], od["transmission"], c, normalise=True, effect="opd", device=dev)
    stop = torch.as_tensor((np.hypot(*np.mgrid[:N, :N] - (N - 1) / 2) <= 0.4 * N).astype(np.float32), device=dev)
    c = torch.as_tensor(od["coefficients"], device=dev).requires_grad_(True)
    layers = [("optic", mk_optic(c)), ("to_focal", dl.MFT(M, ps)),
              ("to_pupil", dl.MFT(N, np.float32(1.0 / N), inverse=True)), ("lyot", dl.Optic(stop, device=dev)),
              ("to_focal2", dl.MFT(M, ps))]
    sys_ = dl.LayeredOpticalSystem(N, 1.0, layers, device=dev)
    assert sys_._batchable()
    psf = sys_.propagate(wls, off, w)
    psf.sum().backward()
    g_batched = c.grad.clone()
    # the same stack, one wavelength at a time
    c2 = torch.as_tensor(od["coefficients"], device=dev).requires_grad_(True)
    layers2 = [("optic", mk_optic(c2))] + layers[1:]
    sys2 = dl.LayeredOpticalSystem(N, 1.0, layers2, device=dev)
    ref = sum(sys2.propagate(wls[l:l + 1], off, w[l:l + 1]) for l in range(len(wls)))
    ref.sum().backward()
    assert rel_l2(psf.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    assert rel_l2(g_batched.cpu().numpy(), c2.grad.cpu().numpy()) < 1e-5
    # first leg against the oracle
    first = dl.LayeredOpticalSystem(N, 1.0, layers[:2], device=dev).propagate(wls, off, w)
    want = O.propagate(od, wls, off, w)
    assert rel_l2(first.detach().cpu().numpy(), want) < TOL


def test_wavefront_container_api(dev):
    # the rest of the Wavefront surface next to the path (wavefronts.py:124-167, 196-266, 426-440,
    # 834-920): from_phasor, real/imaginary/complex/polar, flip, + - * / between wavefronts / arrays
    import dlux_b200 as dl
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))).astype(np.complex64)
    wf = dl.Wavefront.from_phasor(x, 1e-6, diameter=1.0, device=dev)
    assert wf.npixels == 8 and abs(float(wf.pixel_scale) - 0.125) < 1e-7 and wf.ndim == 0
    np.testing.assert_array_equal(wf.real.cpu().numpy(), x.real)
    np.testing.assert_array_equal(wf.complex.cpu().numpy(), np.stack([x.real, x.imag]))
    np.testing.assert_allclose(wf.polar.cpu().numpy(), np.stack([np.abs(x), np.angle(x)]), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(wf.flip(0).phasor.cpu().numpy(), np.flip(x, 0))
    np.testing.assert_array_equal(wf.flip((0, 1)).phasor.cpu().numpy(), np.flip(x, (0, 1)))
    np.testing.assert_allclose((wf + wf).phasor.cpu().numpy(), 2 * x, rtol=1e-6)
    np.testing.assert_allclose((wf - x).phasor.cpu().numpy(), 0 * x, atol=1e-7)
    np.testing.assert_allclose((wf / 2.0).phasor.cpu().numpy(), x / 2, rtol=1e-6)
    assert (wf * None) is wf and (wf + None) is wf
    with pytest.raises(TypeError):
        wf + "a"
    with pytest.raises(ValueError):
        dl.Wavefront.from_phasor(x, 1e-6, device=dev)
    # a propagated from_phasor wavefront equals the functional MFT
    out = wf.propagate(6, np.float32(2e-7)).phasor.cpu().numpy()
    assert rel_l2(out, O.MFT(x, 1e-6, np.float32(0.125), 6, np.float32(2e-7))) < TOL


def test_dynamic_apertures_fused_and_differentiable(dev):
    # NEXT-3: parametrised soft-edged apertures produce the transmission of the fused route; the
    # transmission cotangent of dlux_polypsf_bwd makes their parameters fitted parameters.
    # Forward vs the oracle fed with the same transmission; gradients w.r.t. the primary radius and
    # the spider rotation vs float64 autograd of the oracle twin through the same geometry code.
    import dlux_b200 as dl
    from dlux_b200.utils import geometry as G
    from oracle import torch_twin
    N, M = 96, 48
    rng = np.random.default_rng(12)
    wls = np.linspace(0.9e-6, 1.1e-6, 3).astype(np.float32)
    w = np.array([0.3, 0.3, 0.4], np.float32)
    off = np.array([1.0e-7, -2.0e-7], np.float32)
    Gc = rng.standard_normal((M, M))

    def build(radius, rot, dtype, device):
        xf = dl.CoordTransform(translation=np.array([0.01, -0.02], np.float32), rotation=rot)
        return dl.CompoundAperture([
            ("primary", dl.CircularAperture(radius, softening=2.0)),
            ("secondary", dl.CircularAperture(np.float32(0.12), occulting=True, softening=2.0)),
            ("spiders", dl.Spider(np.float32(0.02), [0.0, 120.0, 240.0], softening=2.0)),
        ], transformation=xf, normalise=True)

    radius = torch.tensor(0.45, device=dev, requires_grad=True)
    rot = torch.tensor(0.2, device=dev, requires_grad=True)
    ap = build(radius, rot, torch.float32, dev)
    for fused in (True, False):
        radius.grad = rot.grad = None
        sys_ = dl.AngularOpticalSystem(N, 1.0, [("aperture", ap)], M, 0.05, device=dev, fused=fused)
        psf = sys_.propagate(wls, off, w)
        (psf * torch.as_tensor(Gc.astype(np.float32), device=dev)).sum().backward()
        # reference: same geometry in float64 on the CPU -> oracle twin
        r64 = torch.tensor(0.45, dtype=torch.float64, requires_grad=True)
        q64 = torch.tensor(0.2, dtype=torch.float64, requires_grad=True)
        c64 = G.pixel_coords(N, 1.0, dtype=torch.float64)
        T64 = build(r64, q64, torch.float64, "cpu").transmission(c64, torch.tensor(1.0 / N, dtype=torch.float64))
        ref = torch_twin.poly_psf(T64, None, wls, w, diameter=1.0, psf_npixels=M,
                                  pixel_scale_rad=O.arcsec2rad(0.05), offset=off, normalise=True, dtype=np.float64)
        (ref * torch.tensor(Gc)).sum().backward()
        check(f"dynamic aperture psf [fused={fused}]", rel_l2(psf.detach().cpu().numpy(), ref.detach().numpy()), TOL)
        # a shape parameter's gradient is <T_bar, dT/d(param)> restricted to the soft-edge pixels: they carry
        # 7 % of the norm of T_bar and the inner product cancels 11.5-fold (measured with the float64 twin), so
        # the transmission cotangent's own 2.5e-6 (test_transmission_and_phase_gradients) shows up as 1-2e-5
        # here.  Conditioning of the derived quantity, not a looser kernel
        check(f"radius grad [fused={fused}]", rel_scalar(radius.grad, r64.grad), 5e-5)
        check(f"rotation grad [fused={fused}]", rel_scalar(rot.grad, q64.grad), 5e-5)


def test_second_order_through_layer_route(dev):
    # NEXT-1, second order: MFTFunction / BasisEvalFunction are closed under differentiation
    # (each backward is the other linear operator as an autograd Function), so the Hessian of a
    # loss w.r.t. the basis coefficients comes out of double backward on the layer-by-layer
    # route (what zdx.hessian / a Fisher matrix needs).  Against the float64 Hessian of the twin.
    import dlux_b200 as dl
    from oracle import torch_twin
    N, M, nz = 48, 24, 3
    od = _optics_dict(N, M, nz, 21)
    od["basis"] = od["basis"] * np.float32(4.0)       # stronger curvature
    wls = np.array([0.95e-6, 1.05e-6], np.float32)
    w = np.array([0.5, 0.5], np.float32)
    rng = np.random.default_rng(22)
    target = rng.uniform(0.5, 1.5, (M, M))
    basis_d = torch.as_tensor(od["basis"], device=dev)

    def loss_gpu(c):
        layer = dl.BasisOptic(basis_d, od["transmission"], c, normalise=True, effect="opd", device=dev)
        sys_ = dl.AngularOpticalSystem(N, 1.0, [("a", layer)], M, 0.05, device=dev, fused=False)
        psf = sys_.propagate(wls, None, w)
        return ((psf * 1e3 - torch.as_tensor(target.astype(np.float32), device=dev)) ** 2).sum()

    def loss_ref(c):
        psf = torch_twin.poly_psf(od["transmission"], None, wls, w, diameter=1.0, psf_npixels=M,
                                  pixel_scale_rad=O.arcsec2rad(0.05), basis=od["basis"], coefficients=c,
                                  dtype=np.float64)
        return ((psf * 1e3 - torch.tensor(target)) ** 2).sum()

    c0 = od["coefficients"]
    H = torch.autograd.functional.hessian(loss_gpu, torch.as_tensor(c0, device=dev)).cpu().numpy().astype(np.float64)
    Href = torch.autograd.functional.hessian(loss_ref, torch.tensor(c0, dtype=torch.float64)).numpy()
    check("hessian symmetry (layer route)", rel_l2(H, H.T), TOL)
    check("hessian vs float64 twin (layer route)", rel_l2(H, Href), TOL)


def test_binary_source(dev):
    # BinarySource (sources.py:524-635) = two point sources from (position, separation, position
    # angle, mean flux, contrast); forward vs the oracle's PointSources, gradient w.r.t. the
    # separation and contrast vs float64 autograd of the twin
    import dlux_b200 as dl
    from oracle import torch_twin
    N, M = 64, 32
    od = _optics_dict(N, M, 3, 31)
    wls = np.linspace(0.9e-6, 1.1e-6, 3).astype(np.float32)
    w2 = np.array([[1.0, 2.0, 1.0], [3.0, 1.0, 1.0]], np.float32)
    pos0, sep0, pa0, mf0, con0 = np.array([1e-7, -1e-7], np.float32), 6e-7, 0.7, 2.0, 3.0
    G = np.random.default_rng(32).standard_normal((M, M))
    sep = torch.tensor(sep0, device=dev, requires_grad=True)
    con = torch.tensor(con0, device=dev, requires_grad=True)
    src = dl.BinarySource(wls, pos0, mf0, sep, pa0, con, weights=w2)
    sys_ = _system(od, dev)
    psf = src.model(sys_)
    (psf * torch.as_tensor(G.astype(np.float32), device=dev)).sum().backward()
    # oracle: the same two stars as PointSources with per-star spectra
    wn = w2 / w2.sum(-1)[:, None]
    s64 = torch.tensor(sep0, dtype=torch.float64, requires_grad=True)
    c64 = torch.tensor(con0, dtype=torch.float64, requires_grad=True)
    vec = torch.stack([s64 / 2 * np.sin(pa0), s64 / 2 * np.cos(pa0)])
    positions = torch.stack([torch.tensor(pos0, dtype=torch.float64) + vec, torch.tensor(pos0, dtype=torch.float64) - vec])
    flux = 2 * torch.stack([c64 * mf0, torch.tensor(mf0, dtype=torch.float64)]) / (1 + c64)
    ref = 0
    for s in range(2):
        ref = ref + torch_twin.poly_psf(od["transmission"], None, wls, torch.tensor(wn[s], dtype=torch.float64) * flux[s],
                                        diameter=1.0, psf_npixels=M, pixel_scale_rad=O.arcsec2rad(0.05),
                                        offset=positions[s], basis=od["basis"], coefficients=od["coefficients"],
                                        dtype=np.float64)
    (ref * torch.tensor(G)).sum().backward()
    assert rel_l2(psf.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    check("binary separation grad", rel_scalar(sep.grad, s64.grad), TOL)
    check("binary contrast grad", rel_scalar(con.grad, c64.grad), TOL)


def test_resolved_sources_and_scene(dev):
    # image-plane sources (sources.py:414-521, 638-748, 751-850): PSF (*) distribution after the
    # fused PSF; against the oracle PSF convolved with scipy
    import dlux_b200 as dl
    from scipy.signal import convolve
    N, M = 64, 32
    od = _optics_dict(N, M, 3, 41)
    wls = np.linspace(0.9e-6, 1.1e-6, 3).astype(np.float32)
    pos = np.array([1e-7, -2e-7], np.float32)
    dist = np.random.default_rng(42).uniform(0, 1, (5, 5)).astype(np.float32)
    dist[0, 0] = -1.0                                   # floored by normalise (sources.py:470-474)
    sys_ = _system(od, dev)
    base = O.point_source_model(od, wls, pos, 2.0).astype(np.float64)
    dn = dist / dist.sum()
    dn = np.maximum(dn, 0)
    dn = dn / dn.sum()
    res = dl.ResolvedSource(wls, pos, 2.0, dist).model(sys_).cpu().numpy()
    assert rel_l2(res, convolve(base, dn, mode="same")) < TOL
    pr = dl.PointResolvedSource(wls, pos, 2.0, dist, contrast=3.0).model(sys_).cpu().numpy()
    f = 2 * np.array([3.0 * 2.0, 2.0]) / 4.0
    # sources.py:728-743: the reference propagates with its DEFAULT 1/L spectral weights and then
    # applies the per-component weights, hence the extra 1/L (pinned by reference_classes.npz)
    L = len(wls)
    want = (base / 2.0 * f[0] + convolve(base / 2.0 * f[1], dn, mode="same")) / L
    assert rel_l2(pr, want) < TOL
    scene = dl.Scene([("star", dl.PointSource(wls, pos, 2.0)), ("disk", dl.ResolvedSource(wls, pos, 2.0, dist))])
    assert rel_l2(scene.model(sys_).cpu().numpy(), base + convolve(base, dn, mode="same")) < TOL
    with pytest.raises(NotImplementedError):
        dl.ResolvedSource(wls, pos, 2.0, dist).model(sys_, return_wf=True)


def test_telescope_pipeline(dev):
    # instruments.py:140-172: optics -> Scene of sources -> detector, differentiable end to end
    import dlux_b200 as dl
    N, M = 64, 16
    od = _optics_dict(N, M, 3, 51, oversample=4, pscale=0.2)
    wls = np.linspace(0.9e-6, 1.1e-6, 3).astype(np.float32)
    c = torch.as_tensor(od["coefficients"], device=dev).requires_grad_(True)
    layer = dl.BasisOptic(od["basis"], od["transmission"], c, normalise=True, effect="opd", device=dev)
    optics = dl.AngularOpticalSystem(N, 1.0, [("a", layer)], M, 0.2, 4, device=dev)
    sources = [("star", dl.PointSource(wls, np.array([1e-7, 0.0], np.float32), 5.0)),
               ("binary", dl.BinarySource(wls, None, 2.0, 8e-7, 0.3, 2.0))]
    det = dl.LayeredDetector([dl.ApplyJitter(0.8, 5), dl.Downsample(4), dl.AddConstant(0.001)])
    tel = dl.Telescope(optics, sources, det)
    img = tel.model()
    assert tuple(img.shape) == (M, M)
    # the same chain by hand from the oracle PSFs
    from scipy.signal import convolve
    base = O.point_source_model(od, wls, np.array([1e-7, 0.0], np.float32), 5.0).astype(np.float64)
    vec = np.array([4e-7 * np.sin(0.3), 4e-7 * np.cos(0.3)], np.float32)
    fl = 2 * np.array([2.0 * 2.0, 2.0]) / 3.0
    base = base + O.point_sources_model(od, wls, np.stack([vec, -vec]), fl.astype(np.float32)).astype(np.float64)
    k = det.layers["ApplyJitter_0"].kernel().numpy().astype(np.float64)
    want = convolve(base, k, mode="same").reshape(M, 4, M, 4).sum((1, 3)) + 0.001
    check("telescope image", rel_l2(img.detach().cpu().numpy(), want), TOL)
    img.sum().backward()
    assert torch.isfinite(c.grad).all() and float(c.grad.abs().sum()) > 0
    assert abs(float(tel.model(return_psf=True).pixel_scale) - float(O.arcsec2rad(np.float32(0.2)))) < 1e-12


def test_config4_like_large_pupil_many_sources(dev):
    # BASELINE config 4 shape (scaled down in sources/wavelengths): 2048 px pupil with a binary
    # 0/pi phase mask, several stars, MFT to 256x256
    import dlux_b200 as dl
    N, M = 2048, 256
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[:N, :N]
    r = np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) / (N / 2)
    T = (r <= 1).astype(np.float32)
    f = np.fft.fft2(rng.standard_normal((N, N)))
    f[40:-40, :] = 0
    f[:, 40:-40] = 0
    phase = (np.pi * (np.fft.ifft2(f).real > 0)).astype(np.float32)
    od = dict(wf_npixels=N, diameter=0.125, psf_npixels=M, psf_pixel_scale=0.7, oversample=1,
              transmission=T, phase=phase, normalise=True)
    wls = np.linspace(5.3e-7, 6.4e-7, 3).astype(np.float32)
    pos = (rng.uniform(-0.35, 0.35, (3, 2)) * M * O.arcsec2rad(0.7)).astype(np.float32)
    flux = np.array([1.0, 30.0, 500.0], np.float32)
    layer = dl.Optic(T, None, phase, normalise=True, device=dev)
    sys_ = dl.AngularOpticalSystem(N, 0.125, [("mask", layer)], M, 0.7, device=dev)
    psf = sys_.model(dl.PointSources(wls, pos, flux))
    ref = O.point_sources_model(od, wls, pos, flux)
    assert rel_l2(psf.cpu().numpy(), ref) < TOL


def test_multi_chunk_paths(dev):
    # force the item-chunk loops (DLUX_B200_CHUNK_MB is read once per process, so run a child)
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import numpy as np, torch, sys
        sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
        from test_gpu_parity import _optics_dict, _system
        from oracle import mft_oracle as O
        import dlux_b200 as dl
        dev = torch.device('cuda:0')
        od = _optics_dict(64, 32, 3, 7)
        wls = np.linspace(0.9e-6, 1.1e-6, 5).astype(np.float32)
        pos = np.array([[1e-7, -2e-7], [-3e-7, 0.5e-7], [0, 0]], np.float32)
        flux = np.array([3.0, 0.25, 1.0], np.float32)
        c = torch.as_tensor(od['coefficients'], device=dev).requires_grad_(True)
        layer = dl.BasisOptic(od['basis'], od['transmission'], c, normalise=True, effect="opd", device=dev)
        s = dl.AngularOpticalSystem(64, 1.0, [('a', layer)], 32, 0.05, device=dev)
        psf = s.model(dl.PointSources(wls, pos, flux))
        psf.sum().backward()
        ref = O.point_sources_model(od, wls, pos, flux)
        err = np.linalg.norm(psf.detach().cpu().numpy() - ref) / np.linalg.norm(ref)
        x = torch.as_tensor((np.random.default_rng(0).standard_normal((7, 64, 64)) + 0j).astype(np.complex64), device=dev)
        out = dl.utils.MFT(x, np.full(7, 1e-6, np.float32), np.float32(1 / 64), 32, np.float32(2e-7))
        e2 = max(np.linalg.norm(out[i].cpu().numpy() - O.MFT(x[i].cpu().numpy(), 1e-6, 1 / 64, 32, 2e-7)) /
                 np.linalg.norm(out[i].cpu().numpy()) for i in range(7))
        print('ERR', err, e2, float(c.grad.abs().sum()))
        assert err < 1e-5 and e2 < 1e-5
    """ % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
           os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    env = dict(os.environ, DLUX_B200_CHUNK_MB="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ERR" in r.stdout


# ------------------------------------------------------------------ full-size properties (C3)
@pytest.mark.parametrize("prec", ["3xtf32"])
def test_c3_size_properties(dev, prec):
    """1024 -> 512 (BASELINE config 3 shape): linearity, inverse symmetry and agreement
    with the fp32 CUDA-core path -- properties that do not need the oracle at full size."""
    from dlux_b200 import ops
    rng = np.random.default_rng(11)
    N, M = 1024, 512
    x = torch.as_tensor(_rand_c64(rng, 2, N, N), device=dev)
    s = torch.tensor([0.1221, 0.1180], device=dev)
    nrm = s / N
    a = ops.mft_c64(x, s, M, None, None, nrm, False, False, prec)
    lin = ops.mft_c64((2.0 * x[0] - 0.5j * x[1])[None], s[:1], M, None, None, nrm[:1], False, False, prec)
    a1 = ops.mft_c64(x[1:2], s[:1], M, None, None, nrm[:1], False, False, prec)
    assert rel_l2(lin[0].cpu().numpy(), (2.0 * a[0] - 0.5j * a1[0]).cpu().numpy()) < TOL
    inv = ops.mft_c64(torch.conj(x), s, M, None, None, nrm, True, False, prec)
    assert rel_l2(torch.conj(inv).cpu().numpy(), a.cpu().numpy()) < TOL
    ref = ops.mft_c64(x, s, M, None, None, nrm, False, False, "fp32")
    assert rel_l2(a.cpu().numpy(), ref.cpu().numpy()) < TOL
    # one wavelength against the oracle at full size (a few seconds of CPU)
    ps_in = np.float32(6.6 / N)
    pso = O.arcsec2rad(0.0656 / 4)
    import dlux_b200 as dl
    out = dl.utils.MFT(x[0], np.float32(4.3e-6), ps_in, M, pso, precision=prec)
    assert rel_l2(out.cpu().numpy(), O.MFT(x[0].cpu().numpy(), 4.3e-6, ps_in, M, pso)) < TOL


# ------------------------------------------------------------------ FFT (SURVEY 8f NEXT-2)
@pytest.mark.parametrize("prec", PRECS)
def test_fft_golden_reference_vectors(dev, golden, prec):
    import dlux_b200 as dl
    for j in range(int(golden["n_fft"])):
        pad, inverse, ps = golden[f"fft_{j}_meta"]
        ph = torch.as_tensor(golden[f"fft_{j}_in"], device=dev)
        out, gps = dl.utils.FFT(ph, np.float32(1.0e-6), np.float32(0.01), None, int(pad), bool(inverse),
                                precision=prec)
        ref = golden[f"fft_{j}_out"]
        assert tuple(out.shape) == ref.shape
        assert rel_l2(out.cpu().numpy(), ref) < TOL, (j, rel_l2(out.cpu().numpy(), ref))
        assert abs(float(gps) - ps) <= 1e-6 * abs(ps)


def test_fft_large_sizes_against_numpy_fft(dev):
    # the reference's FFT is an FFT: no float32 phase-matrix rounding.  The exact-DFT mode of the kernels must hold
    # 1e-5 at sizes where the MFT's float32 phase form would not (1e-4 at N_pad = 2048)
    import dlux_b200 as dl
    rng = np.random.default_rng(3)
    for n, pad, inverse in ((1024, 2, False), (512, 3, True), (200, 2, False), (129, 3, False)):
        x = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
        out, ps = dl.utils.FFT(torch.as_tensor(x, device=dev), np.float32(1e-6), np.float32(0.01), None, pad, inverse)
        ref, ps_ref = O.FFT(x.astype(np.complex128), 1e-6, 0.01, None, pad, inverse, dtype=np.float64)
        check(f"FFT n={n} pad={pad} inverse={inverse}", rel_l2(out.cpu().numpy(), ref), TOL)
        assert abs(float(ps) - float(ps_ref)) < 1e-6 * float(ps_ref)


def test_fft_roundtrip_and_layer(dev):
    # /root/reference/tests/utils/test_propagation.py:74-92 and tests/layers/test_propagators.py:48-63
    import dlux_b200 as dl
    ph = torch.ones((32, 32), dtype=torch.complex64, device=dev)
    f, _ = dl.utils.FFT(ph, 1.0, 0.1, 2.0, pad=1)
    b, _ = dl.utils.FFT(f, 1.0, 0.1, 2.0, pad=1, inverse=True)
    assert torch.allclose(b, ph, rtol=1e-5, atol=1e-5)
    # odd sizes and padding follow numpy's fftshift conventions
    rng = np.random.default_rng(12)
    for n, pad in ((31, 2), (32, 3), (15, 3)):
        x = _rand_c64(rng, n, n)
        out, _ = dl.utils.FFT(torch.as_tensor(x, device=dev), 1e-6, 0.01, None, pad)
        ref, _ = O.FFT(x, 1e-6, 0.01, None, pad)
        assert rel_l2(out.cpu().numpy(), ref) < TOL, (n, pad)
    wf = dl.Wavefront(1e-6, 16, diameter=1.0, device=dev)
    out = dl.FFT(pad=2, crop=2, center=False)(wf)
    assert isinstance(out, dl.Wavefront) and out.phasor.shape == (16, 16)
    cen = dl.FFT(pad=2, center=True)(wf)
    assert cen.phasor.shape == (32, 32)
    # re-centring only applies phase ramps: a centred FFT has the intensity of the centred MFT
    mft = wf.propagate(32, float(cen.pixel_scale))
    assert rel_l2(cen.psf.cpu().numpy(), mft.psf.cpu().numpy()) < 1e-4
    with pytest.raises(ValueError, match="cannot specify d"):
        wf.propagate_FFT(spec_out=dl.CoordSpec(d=1.0))


# ------------------------------------------------------------------ API behaviour
def test_api_errors_and_shapes(dev):
    import dlux_b200 as dl
    wf = dl.Wavefront(1e-6, 16, diameter=1.0, device=dev)
    out = wf.propagate(8, 1e-7)
    assert isinstance(out, dl.Wavefront) and out.phasor.shape == (8, 8)
    assert isinstance(dl.MFT(8, 1e-7)(wf), dl.Wavefront)
    with pytest.raises(ValueError):
        dl.Wavefront(1e-6, 16)
    with pytest.raises(ValueError):
        dl.Wavefront(1e-6, 16, diameter=1.0, pixel_scale=0.1)
    sys_ = _system(_optics_dict(16, 8, 2, 5), dev)
    with pytest.raises(ValueError, match="shape mismatch"):
        sys_.propagate(np.array([1e-6, 2e-6]), None, np.ones(3))
    with pytest.raises(ValueError, match="offset must be"):
        sys_.propagate(np.array([1e-6]), np.zeros(3))
    with pytest.raises(ValueError, match="Cannot return both"):
        sys_.propagate(np.array([1e-6]), return_wf=True, return_psf=True)
    with pytest.raises(TypeError):
        dl.utils.MFT(torch.ones((8, 8), dtype=torch.complex128, device=dev), 1e-6, 0.1, 4, 1e-7)
    with pytest.raises(ValueError):
        dl.utils.MFT(torch.ones((8, 8), dtype=torch.complex64), 1e-6, 0.1, 4, 1e-7)  # CPU tensor
