"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares,
host-side geometry matches the oracle bit for bit, and the mirrored API validates its
arguments like the reference.  No compute calls (there is no GPU here)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import mft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    from dlux_b200 import build, _lib
    build.build()                       # no-op when up to date; nvcc cross-compiles on CPU
    return _lib.load()


def test_header_symbols_exported(lib):
    from dlux_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dlux_b200.h")).read()
    declared = set(re.findall(r"DLUX_API\s+[\w\s\*]+?\b(dlux_\w+)\s*\(", hdr))
    assert len(declared) >= 12
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dlux_\w+)", nm))
    assert declared <= exported, declared - exported
    assert lib.dlux_abi_version() == 2
    assert lib.dlux_error_string(-3) == b"scratch buffer too small"


def test_scratch_sizes_and_arg_checks(lib):
    import ctypes as C
    from dlux_b200._lib import MftDesc, PolyPsfDesc
    d = MftDesc(1024, 512, 64, 0, 0, 0)
    n = lib.dlux_mft_scratch_bytes(C.byref(d))
    assert 64 * 1024 * 1024 < n < 4 * 1024 ** 3
    assert lib.dlux_mft_scratch_bytes(C.byref(MftDesc(0, 8, 1, 0, 0, 0))) == 0
    p = PolyPsfDesc(1024, 512, 64, 1, 1, 0, 1, 0)
    assert lib.dlux_polypsf_scratch_bytes(C.byref(p)) > 64 * 1024 * 1024
    # null pointers are rejected before any CUDA call
    assert lib.dlux_mft_c64(C.byref(d), None, None, None, None, None, None, None, 0, None) == -1
    assert lib.dlux_mft_c64(C.byref(MftDesc(8, 8, 1, 0, 0, 7)), None, None, None, None, None, None,
                            None, 0, None) == -1
    assert lib.dlux_basis_eval(0, 10, None, None, None, None, None) == -1


def test_round2_descriptors_host_checks(lib):
    """ABI v2 additions, checked on the host side only (no CUDA call is reached): the exact-DFT period of
    ``dlux_mft_desc``, the parameter-batch descriptor and the scratch queries that size the fused launch's ring."""
    import ctypes as C
    from dlux_b200._lib import MftDesc, PolyPsfBatchDesc, PolyPsfDesc
    assert lib.dlux_abi_version() == 2
    # dlu.FFT through the MFT kernels: period = padded size; negative periods and index products that are not
    # exact in float32 are refused
    assert lib.dlux_mft_scratch_bytes(C.byref(MftDesc(1024, 2048, 1, 0, 0, 0, 2048, 0))) > 0
    assert lib.dlux_mft_scratch_bytes(C.byref(MftDesc(64, 128, 1, 0, 0, 0, -1, 0))) == 0
    assert lib.dlux_mft_scratch_bytes(C.byref(MftDesc(4096, 8192, 1, 0, 0, 0, 8192, 0))) == 0
    assert lib.dlux_mft_c64(C.byref(MftDesc(64, 128, 1, 0, 0, 0, -1, 0)), None, None, None, None, None, None,
                            None, 0, None) == -1
    # the sparse option does not change the scratch layout; more items never need less scratch
    dense = lib.dlux_polypsf_scratch_bytes(C.byref(PolyPsfDesc(1024, 512, 64, 1, 1, 0, 1, 0)))
    assert lib.dlux_polypsf_scratch_bytes(C.byref(PolyPsfDesc(1024, 512, 64, 1, 1, 0, 1, 1))) == dense
    assert lib.dlux_polypsf_scratch_bytes(C.byref(PolyPsfDesc(1024, 512, 64, 4, 1, 0, 1, 0))) >= dense
    assert lib.dlux_polypsf_scratch_bytes(C.byref(PolyPsfDesc(1024, 512, 0, 1, 1, 0, 1, 0))) == 0
    b = PolyPsfBatchDesc()
    b.n_pupil, b.n_psf, b.n_wavels, b.n_batch, b.n_basis, b.normalise = 1024, 256, 32, 128, 10, 1
    n128 = lib.dlux_polypsf_batch_scratch_bytes(C.byref(b))
    assert n128 > 0
    b.n_batch = 4096                                  # chunked by whole batch elements: bounded scratch
    assert n128 <= lib.dlux_polypsf_batch_scratch_bytes(C.byref(b)) < 40 * 1024 ** 3
    b.n_basis = 0
    assert lib.dlux_polypsf_batch_scratch_bytes(C.byref(b)) == 0


def test_missing_library_is_loud(monkeypatch):
    from dlux_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdlux_b200.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_cpu_tensors_rejected():
    import dlux_b200 as dl
    with pytest.raises(ValueError, match="CUDA tensor"):
        dl.utils.MFT(torch.ones((8, 8), dtype=torch.complex64), 1e-6, 0.1, 4, 1e-7)


def test_geometry_scalars_match_oracle_bits():
    from dlux_b200.utils import propagation as P
    rng = np.random.default_rng(0)
    for _ in range(50):
        wl = np.float32(rng.uniform(4e-7, 5e-6))
        n_in, n_out = int(rng.integers(16, 2049)), int(rng.integers(8, 1025))
        psi = np.float32(rng.uniform(0.1, 8.0) / n_in)
        pso = np.float32(rng.uniform(1e-8, 1e-6))
        fl = None if rng.random() < 0.5 else np.float32(rng.uniform(0.5, 20.0))
        s, nrm = P.mft_geometry(wl, n_in, psi, n_out, pso, fl)
        assert np.float32(s) == O.mft_scalars(wl, n_in, psi, pso, fl)
        nf = O.calc_nfringes(wl, n_in, psi, n_out, pso, fl)
        assert np.float32(P.calc_nfringes(wl, n_in, psi, n_out, pso, fl)) == nf
        assert np.float32(nrm) == O.mft_norm(nf, n_in, n_out)
    # vectorised over wavelengths
    wls = np.linspace(0.9e-6, 1.1e-6, 7).astype(np.float32)
    s, nrm = P.mft_geometry(wls, 512, np.float32(1 / 512), 256, np.float32(1.2e-7))
    for i, wl in enumerate(wls):
        assert s[i] == O.mft_scalars(wl, 512, np.float32(1 / 512), np.float32(1.2e-7))
    assert P.arcsec2rad(np.float32(0.05)) == O.arcsec2rad(0.05)


def test_fusable_logic_and_validation():
    import dlux_b200 as dl
    T = np.ones((8, 8), np.float32)
    basis = np.zeros((2, 8, 8), np.float32)
    mk = lambda layers: dl.AngularOpticalSystem(8, 1.0, layers, 4, 0.1, device="cpu")
    s1 = mk([dl.Optic(T, np.zeros((8, 8), np.float32), normalise=True, device="cpu")])
    T_, opd, phase, norm = s1._fusable()
    assert norm and phase is None and opd.shape == (8, 8) and T_.shape == (8, 8)
    # a transmission after the normalisation changes the power: not fusable
    s2 = mk([dl.Optic(T, normalise=True, device="cpu"), dl.TransmissiveLayer(T * 0.5, device="cpu")])
    assert s2._fusable() is None
    # an MFT layer in the stack is not pupil-only
    s3 = mk([dl.Optic(T, device="cpu"), dl.MFT(4, 1e-7)])
    assert s3._fusable() is None
    assert s1._can_fuse() and not s2._can_fuse() and not s3._can_fuse()
    # evaluating a basis needs the CUDA kernel: loud error on a CPU tensor
    s4 = mk([dl.BasisOptic(basis, T, np.zeros(2, np.float32), normalise=True, device="cpu")])
    with pytest.raises(ValueError, match="CUDA tensor"):
        s4._fusable()
    with pytest.raises(ValueError, match="effect must be"):
        dl.BasisLayer(basis, effect="bogus", device="cpu")
    with pytest.raises(ValueError, match="same shape"):
        dl.AberratedLayer(np.zeros((4, 4)), np.zeros((5, 5)), device="cpu")
    with pytest.raises(ValueError, match="2d array"):
        dl.PointSources(np.array([1e-6]), np.zeros(2))
    with pytest.raises(ValueError, match="Length of flux"):
        dl.PointSources(np.array([1e-6]), np.zeros((3, 2)), np.ones(2))
    src = dl.PointSource(np.array([1e-6, 2e-6]), weights=np.array([2.0, 6.0]))
    assert np.allclose(src.normalised_weights(), [0.25, 0.75])
    npix, ps, fl = s1._focal_args()
    assert npix == 4 and fl is None and ps == O.arcsec2rad(np.float32(0.1))


def test_downsample_matches_reference_reduction_order():
    # dlu.downsample (utils/array_ops.py:124-161): block mean / sum, columns first
    import torch
    from dlux_b200.utils import downsample
    rng = np.random.default_rng(3)
    a = rng.standard_normal((12, 12)).astype(np.float32)

    def ref(array, n, mean):
        method = np.mean if mean else np.sum
        size_in, size_out = array.shape[0], array.shape[0] // n
        array = method(array.reshape((size_in * size_out, n)), 1).reshape(size_in, size_out).T
        return method(array.reshape((size_out * size_out, n)), 1).reshape(size_out, size_out).T

    for n in (2, 3, 4):
        for mean in (True, False):
            got = downsample(torch.as_tensor(a), n, mean).numpy()
            np.testing.assert_allclose(got, ref(a, n, mean), rtol=1e-6, atol=1e-7)
    assert downsample(torch.as_tensor(np.stack([a, 2 * a])), 4, False).shape == (2, 3, 3)
    with pytest.raises(ValueError):
        downsample(torch.zeros(10, 10), 3)
    with pytest.raises(ValueError):
        downsample(torch.zeros(10, 8), 2)


def test_convolve_same_matches_scipy():
    # jax.scipy.signal.convolve(mode="same") semantics used by ResolvedSource (sources.py:516)
    import torch
    from scipy.signal import convolve
    from dlux_b200.sources import convolve_same
    rng = np.random.default_rng(8)
    img = rng.standard_normal((17, 17))
    for shape in ((3, 3), (4, 4), (5, 2), (1, 1)):
        k = rng.uniform(0, 1, shape)
        got = convolve_same(torch.as_tensor(img), torch.as_tensor(k)).numpy()
        np.testing.assert_allclose(got, convolve(img, k, mode="same"), rtol=1e-10, atol=1e-12)


def test_unified_layers_on_wavefront_and_psf():
    # layers/unified_layers.py:25-66, 136-212; psfs.py:159-190; wavefronts.py:426-440, 568-582
    import dlux_b200 as dl
    rng = np.random.default_rng(3)
    img = rng.random((6, 6)).astype(np.float32)
    psf = dl.PSF(img, np.float32(0.1))
    wf = dl.Wavefront(1e-6, 6, diameter=1.0, device="cpu") * torch.as_tensor(img)
    for target, get in ((psf, lambda t: t.data.numpy()), (wf, lambda t: t.amplitude.numpy() * 36)):
        np.testing.assert_allclose(get(dl.Resize(4)(target)), img[1:5, 1:5], rtol=1e-6)
        padded = get(dl.Resize(10)(target))
        assert padded.shape == (10, 10) and np.all(padded[:2] == 0)
        np.testing.assert_allclose(padded[2:8, 2:8], img, rtol=1e-6)
        np.testing.assert_allclose(get(dl.Flip(0)(target)), img[::-1], rtol=1e-6)
        np.testing.assert_allclose(get(dl.Flip((0, 1))(target)), img[::-1, ::-1], rtol=1e-6)
        assert dl.Lambda()(target) is target and dl.Resize(6)(target) is target
        with pytest.raises(ValueError):
            dl.Resize(5)(target)                       # even -> odd is not centre preserving
    with pytest.raises(ValueError):
        dl.Flip(0.5)
    with pytest.raises(ValueError):
        dl.Flip((0, "x"))
    # in a layer list, in front of the propagator
    layers = [("pad", dl.Resize(8)), ("noop", dl.Lambda())]
    out = wf
    for _, layer in layers:
        out = layer(out)
    assert out.npixels == 8 and isinstance(out, dl.Wavefront)


def test_interpolating_operations_against_scipy_bilinear():
    """utils/interpolation.py:13-107, wavefronts.py:442-566, psfs.py:112-157, unified_layers.py:69-133.  The reference's
    arithmetic lives in interpax (absent): the bilinear definition is pinned to SciPy's RegularGridInterpolator."""
    from scipy.interpolate import RegularGridInterpolator
    import dlux_b200 as dl
    from dlux_b200.utils import geometry as G, interpolation as I
    rng = np.random.default_rng(11)
    n = 16
    img = rng.standard_normal((n, n))

    def scipy_at(image, knots, samples, fill=0.0):
        xs, ys = knots[0][0].numpy(), knots[1][:, 0].numpy()
        f = RegularGridInterpolator((ys, xs), image, method="linear", bounds_error=False, fill_value=fill)
        return f(np.stack([samples[1].numpy().ravel(), samples[0].numpy().ravel()], 1)).reshape(samples[0].shape)

    t = torch.as_tensor(img)
    # rotate: unit-pixel centred grid, samples = rotated grid, zeros outside
    knots = G.pixel_coords(n, float(n), dtype=torch.float64)
    np.testing.assert_allclose(I.rotate(t, 0.4).numpy(), scipy_at(img, knots, G.rotate_coords(knots, 0.4)), atol=1e-13)
    np.testing.assert_allclose(I.rotate(t, 0.0).numpy(), img, atol=1e-13)
    np.testing.assert_allclose(I.rotate(t, np.pi / 2).numpy()[1:-1, 1:-1], np.rot90(img, 1)[1:-1, 1:-1], atol=1e-12)
    # scale: npixels_out pixels, each `ratio` input pixels wide
    k_in = G.pixel_coords(n, 1.0, dtype=torch.float64)
    k_out = G.pixel_coords(24, 1.0, dtype=torch.float64) * (0.5 * 24 / n)
    np.testing.assert_allclose(I.scale(t, 24, 0.5).numpy(), scipy_at(img, k_in, k_out), atol=1e-13)
    np.testing.assert_allclose(I.scale(t, n, 1.0).numpy(), img, atol=1e-13)
    with pytest.raises(NotImplementedError):
        I.rotate(t, 0.1, method="cubic")
    # Wavefront: both field decompositions, new pixel scale, layer form; differentiable in the angle
    ph = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    wf = dl.Wavefront.from_phasor(torch.as_tensor(ph), 1e-6, pixel_scale=np.float32(0.1))
    r = wf.rotate(np.float32(0.3))
    k32 = G.pixel_coords(n, float(n))
    want = scipy_at(ph.real, k32.double(), G.rotate_coords(k32, 0.3).double()) + \
        1j * scipy_at(ph.imag, k32.double(), G.rotate_coords(k32, 0.3).double())
    np.testing.assert_allclose(r.phasor.numpy(), want, atol=2e-5)
    rp = dl.Rotate(np.float32(0.3))(wf)                        # layer default: amplitude / phase fields
    want_p = scipy_at(np.abs(ph), k32.double(), G.rotate_coords(k32, 0.3).double()) * np.exp(
        1j * scipy_at(np.angle(ph), k32.double(), G.rotate_coords(k32, 0.3).double()))
    np.testing.assert_allclose(rp.phasor.numpy(), want_p, atol=2e-5)
    s = wf.scale_to(20, np.float32(0.05))
    assert s.npixels == 20 and abs(float(s.pixel_scale) - 0.05) < 1e-8
    np.testing.assert_allclose(s.phasor.real.numpy(), I.scale(torch.as_tensor(ph.real), 20, 0.5).numpy(), atol=1e-6)
    tf = dl.CoordTransform(translation=np.array([0.12, -0.07], np.float32), rotation=np.float32(0.2))
    wi = wf.interpolate(tf, fill=0.0)
    kw = wf.coordinates()
    np.testing.assert_allclose(wi.phasor.real.numpy(), scipy_at(ph.real, kw.double(), tf(kw).double()), atol=2e-5)
    with pytest.raises(TypeError):
        wf.interpolate("shift")
    ang = torch.tensor(0.3, requires_grad=True)
    wf.rotate(ang).phasor.real.sum().backward()
    assert ang.grad is not None and torch.isfinite(ang.grad)
    # PSF
    psf = dl.PSF(np.abs(ph) ** 2, np.float32(0.1))
    np.testing.assert_allclose(dl.Rotate(np.float32(0.3))(psf).data.numpy(),
                               scipy_at(np.abs(ph) ** 2, k32.double(), G.rotate_coords(k32, 0.3).double()), atol=2e-5)
    kp = G.pixel_coords(n, n * 0.1)
    np.testing.assert_allclose(psf.interpolate(tf).data.numpy(),
                               scipy_at(np.abs(ph) ** 2, kp.double(), tf(kp).double()), atol=2e-5)
    np.testing.assert_allclose(dl.ApplyInterpolation(tf, fill=0.5)(psf).data.numpy(),       # detector_layers.py:68-97
                               scipy_at(np.abs(ph) ** 2, kp.double(), tf(kp).double(), fill=0.5), atol=2e-5)
    with pytest.raises(TypeError):
        dl.ApplyInterpolation("shift")


def test_bench_clock_sampler_window():
    """bench.ClockSampler reports the samples of the timed window only (each sample carries its wall-clock time)."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.source = "nvml"
    s.rows = [(100.0 + 0.01 * i, 1965.0 if i < 50 else 1650.0, 1965.0, 300.0 + i, ["sw_power_cap"] if i >= 50 else [])
              for i in range(100)]
    early = s.stop(100.0, 100.2)
    assert early["samples"] == 21 and early["sm_mhz"] == 1965.0 and early["reasons"] == []
    late = s.stop(100.6, 101.0)
    assert late["samples"] == 40 and late["sm_mhz"] == 1650.0 and late["reasons"] == ["sw_power_cap"]
    assert s.stop(200.0, 201.0)["samples"] == 0 and s.stop()["samples"] == 100
    assert bench.mft_flops(1024, 512) == 8 * 512 * 1024 * (1024 + 512)


def test_detector_layers_cpu():
    # layers/detector_layers.py:100-296, detectors.py:103-128, psfs.py:74-110 on CPU tensors
    import torch
    from scipy.signal import convolve
    from scipy.stats import norm
    import dlux_b200 as dl
    rng = np.random.default_rng(9)
    img = rng.uniform(0, 1, (24, 24)).astype(np.float32)
    psf = dl.PSF(img, 0.1)
    resp = rng.uniform(0.9, 1.1, (24, 24)).astype(np.float32)
    jit = dl.ApplyJitter(1.5, kernel_size=5, oversample=3)
    # the kernel: normal pdf on linspace(-5, 5, 15), outer product, normalised, summed 3x3
    g = norm.pdf(np.linspace(-5, 5, 15), scale=1.5)
    k = np.outer(g, g)
    k = (k / k.sum()).reshape(5, 3, 5, 3).sum((1, 3))
    np.testing.assert_allclose(jit.kernel().numpy(), k, rtol=1e-5)
    det = dl.LayeredDetector([("resp", dl.ApplyPixelResponse(resp)), ("jitter", jit), ("sat", dl.ApplySaturation(0.9)),
                              ("bias", dl.AddConstant(0.01)), ("bin", dl.Downsample(4))])
    want = np.minimum(convolve(img * resp, k, mode="same"), 0.9) + 0.01
    want = want.reshape(6, 4, 6, 4).sum((1, 3))
    got = det.model(psf)
    np.testing.assert_allclose(got.numpy(), want, rtol=2e-5, atol=1e-6)
    out = det(psf, return_psf=True)
    assert abs(float(out.pixel_scale) - 0.4) < 1e-6 and out.npixels == 6 and det.jitter is jit
    with pytest.raises(ValueError):
        dl.ApplyPixelResponse(np.ones(3))
    with pytest.raises(TypeError):
        dl.LayeredDetector([object()])
