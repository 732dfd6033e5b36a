"""The oracle (oracle/mft_oracle.py) against (a) golden vectors produced by executing
the reference's own source (tests/golden/make_golden.py) and (b) analytic known
answers (SURVEY.md 8c i-vii).  CPU only."""
import numpy as np
import pytest

from oracle import mft_oracle as O
from conftest import rel_l2


def _geom(g):
    n_in, n_out, wl, psi, pso, fl, sx, sy, pixel, inverse = g
    return dict(n_in=int(n_in), n_out=int(n_out), wl=wl, psi=psi, pso=pso,
                fl=None if fl < 0 else fl, shift=(sx, sy), pixel=bool(pixel),
                inverse=bool(inverse))


def test_reference_own_test_fixture(golden):
    # /root/reference/tests/utils/test_propagation.py:12-49,101-139 (values from the
    # reference source itself; the reference's test only checks shape/NaN)
    ones32 = np.ones((32, 32), np.complex64)
    k = 0
    for fl in (None, 2.0):
        for inverse in (False, True):
            for pixel in (True, False):
                got = O.MFT(ones32, 1.0, 0.1, 16, 0.05, fl, (1.0, -2.0), pixel, inverse)
                ref = golden[f"reftest_{k}"]
                assert got.shape == (16, 16) and not np.isnan(got).any()
                assert rel_l2(got, ref) < 2e-6, (k, rel_l2(got, ref))
                k += 1


def test_mft_against_reference_source(golden):
    for g in range(int(golden["n_geoms"])):
        p = _geom(golden[f"mft_{g}_geom"])
        got = O.MFT(golden[f"mft_{g}_in"], p["wl"], p["psi"], p["n_out"], p["pso"], p["fl"],
                    p["shift"], p["pixel"], p["inverse"])
        ref = golden[f"mft_{g}_out"]
        assert rel_l2(got, ref) < 2e-6, (g, rel_l2(got, ref))
        nf = O.calc_nfringes(p["wl"], p["n_in"], p["psi"], p["n_out"], p["pso"], p["fl"])
        assert nf == golden[f"mft_{g}_nfringes"]


def test_transfer_matrix_bits(golden):
    # same argument formation => identical float32 phase arguments; cos/sin are the
    # same NumPy kernels on both sides, so the matrices must match exactly.
    for g in range(int(golden["n_geoms"])):
        p = _geom(golden[f"mft_{g}_geom"])
        tm = O.transfer_matrix(p["wl"], p["n_in"], p["psi"], p["n_out"], p["pso"],
                               p["shift"][0] if p["pixel"] else 0.0, p["fl"], 0.0, p["inverse"])
        assert np.array_equal(tm, golden[f"mft_{g}_tmx"]), g


def test_nd_coords_bits(golden):
    for j in range(int(golden["n_coords"])):
        n, sc, off = golden[f"coords_{j}_args"]
        got = O.nd_coords_1d(int(n), sc, off)
        assert np.array_equal(got, golden[f"coords_{j}"]), j


def test_fft_against_reference_source(golden):
    for j in range(int(golden["n_fft"])):
        pad, inverse, ps = golden[f"fft_{j}_meta"]
        got, gps = O.FFT(golden[f"fft_{j}_in"], 1.0e-6, 0.01, None, int(pad), bool(inverse))
        assert got.shape == golden[f"fft_{j}_out"].shape
        assert rel_l2(got, golden[f"fft_{j}_out"]) < 1e-6
        assert abs(gps - ps) <= 1e-7 * abs(ps)


def test_fft_roundtrip_reference_test():
    # /root/reference/tests/utils/test_propagation.py:74-92
    ph = np.ones((32, 32), np.complex64)
    f, _ = O.FFT(ph, 1.0, 0.1, 2.0, pad=1)
    b, _ = O.FFT(f, 1.0, 0.1, 2.0, pad=1, inverse=True)
    assert np.allclose(b, ph, rtol=1e-6, atol=1e-6)


def test_shift_units_identity_reference_test():
    # /root/reference/tests/utils/test_propagation.py:141-176
    ph = np.ones((32, 32), np.complex64)
    a = O.MFT(ph, 1.0, 0.1, 16, 0.05, 2.0, (1.0, -2.0), True)
    b = O.MFT(ph, 1.0, 0.1, 16, 0.05, 2.0, (0.05, -0.1), False)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- analytic known answers
def test_dirichlet_kernel():
    # (i) all-ones pupil: E[a,b] = (s/N) D(u_a) D(u_b), D(u) = sin(pi u)/sin(pi u/N)
    N, M = 64, 32
    wl, psi, pso = 1.0e-6, 1.0 / N, 1.5e-7
    E = O.MFT(np.ones((N, N)), wl, psi, M, pso, dtype=np.float64)
    s = pso * (N * psi) / wl
    u = (np.arange(M) - (M - 1) / 2) * s
    with np.errstate(invalid="ignore", divide="ignore"):
        D = np.where(np.abs(u) < 1e-15, N, np.sin(np.pi * u) / np.sin(np.pi * u / N))
    ref = (s / N) * np.outer(D, D)
    assert rel_l2(E, ref) < 1e-12
    E32 = O.MFT(np.ones((N, N)), wl, psi, M, pso)
    assert rel_l2(E32, ref) < 3e-6


def test_unitary_case_power_conserved():
    # (ii) M = N, s = 1  -> unitary DFT
    N = 48
    rng = np.random.default_rng(0)
    P = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    E = O.MFT(P, 1.0, 1.0 / N, N, 1.0, dtype=np.float64)
    assert abs((np.abs(E) ** 2).sum() / (np.abs(P) ** 2).sum() - 1) < 1e-12


def test_tilt_equals_coordinate_shift():
    # (iii) SURVEY F6: a pupil tilt theta equals evaluating the DFT phasors at output
    # coordinates shifted by theta/(lambda/D) fringes.
    N, M = 64, 32
    wl, D, pso = 1.0e-6, 1.0, 1.3e-7
    theta = np.array([2.3e-7, -4.1e-7])
    rng = np.random.default_rng(1)
    P = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    wf = O.OracleWavefront(wl, N, D, np.float64)
    wf.phasor = P.copy()
    wf.tilt(theta)
    E_tilt = O.MFT(wf.phasor, wl, D / N, M, pso, dtype=np.float64)
    s = pso * D / wl
    xs = (np.arange(N) - (N - 1) / 2) / N
    ux = (np.arange(M) - (M - 1) / 2) * s - theta[0] * D / wl
    uy = (np.arange(M) - (M - 1) / 2) * s - theta[1] * D / wl
    Ax = np.exp(-2j * np.pi * np.outer(xs, ux))
    Ay = np.exp(-2j * np.pi * np.outer(xs, uy))
    E_shift = (Ay.T @ P @ Ax) * (s / N)
    assert rel_l2(E_tilt, E_shift) < 1e-12


def test_inverse_is_conjugate_symmetry():
    # (vii) MFT(inverse=True)(P) = conj(MFT(conj(P)))
    N, M = 40, 24
    rng = np.random.default_rng(2)
    P = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    a = O.MFT(P, 1e-6, 1.0 / N, M, 2e-7, inverse=True, dtype=np.float64)
    b = np.conj(O.MFT(np.conj(P), 1e-6, 1.0 / N, M, 2e-7, dtype=np.float64))
    assert rel_l2(a, b) < 1e-13


def test_cropped_fft_equals_mft():
    # (vi) FFT(pad) cropped == MFT at ps_out = fringe/pad with a half-pixel shift
    N, pad = 32, 2
    rng = np.random.default_rng(3)
    P = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    wl, psi = 1e-6, 0.01
    F, ps = O.FFT(P, wl, psi, None, pad, dtype=np.float64)
    M = N * pad
    # FFT centring: in index n - N_pad/2, out index k - N_pad/2 on the padded grid
    # <=> MFT with shift = +0.5 px in both planes; the in-plane half pixel is a pure
    # linear phase which we apply analytically.
    E = O.MFT(P, wl, psi, M, ps, shift=(0.5, 0.5), dtype=np.float64)
    # compare intensities (centring conventions differ by linear phases only)
    assert rel_l2(np.abs(E) ** 2, np.abs(F) ** 2) < 1e-10


def test_f32_oracle_error_budget_c3_like():
    # error of the float32 restatement against its float64 twin, C3-like fringes
    N, M = 256, 128
    rng = np.random.default_rng(4)
    P = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))) / N
    wl, D = 4.3e-6, 6.6
    pso = O.arcsec2rad(0.0656 / 4, np.float64) * 4  # same nfringes as C3 (~62)
    a = O.MFT(P, wl, D / N, M, pso)
    b = O.MFT(P, wl, D / N, M, pso, dtype=np.float64)
    # ~4e-6: the reference's own float32 phase-argument rounding (|arg| ~ 100 rad).
    # This is why the CUDA generator reproduces the reference's two float32 roundings
    # of the argument instead of evaluating the phase "exactly" (DESIGN.md).
    assert rel_l2(a, b) < 1e-5


def test_propagate_and_sources_consistency():
    N, M = 32, 16
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[:N, :N]
    r = np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) / (N / 2)
    optics = dict(wf_npixels=N, diameter=1.0, psf_npixels=M, psf_pixel_scale=0.05,
                  oversample=1, transmission=(r <= 1).astype(np.float32),
                  basis=rng.standard_normal((3, N, N)).astype(np.float32) * 1e-8,
                  coefficients=np.array([1.0, -2.0, 0.5], np.float32))
    wls = np.linspace(0.9e-6, 1.1e-6, 3)
    psf = O.propagate(optics, wls)
    assert psf.shape == (M, M) and psf.dtype == np.float32 and (psf >= 0).all()
    # PointSource(flux) == flux * propagate(normalised weights)
    ps1 = O.point_source_model(optics, wls, (1e-7, -2e-7), flux=3.0)
    ps2 = O.propagate(optics, wls, (1e-7, -2e-7), np.full(3, 1 / 3, np.float32) * np.float32(3.0))
    assert rel_l2(ps1, ps2) < 1e-6
    # PointSources == sum of PointSource
    pos = np.array([[1e-7, -2e-7], [-3e-7, 0.5e-7]])
    fl = np.array([3.0, 0.25])
    both = O.point_sources_model(optics, wls, pos, fl)
    summed = sum(O.point_source_model(optics, wls, p, f) for p, f in zip(pos, fl))
    assert rel_l2(both, summed) < 1e-6
    with pytest.raises(ValueError):
        O.propagate(optics, wls, None, np.ones(2))
