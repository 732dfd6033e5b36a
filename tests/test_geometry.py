"""dlux_b200.utils.geometry (torch) against what the reference's own utils/geometry.py and
utils/coordinates.py return (tests/golden/reference_geometry.npz, made by executing those files:
tests/golden/make_golden_geometry.py).  CPU tests: the functions are device-agnostic torch."""
import os

import numpy as np
import pytest
import torch

from dlux_b200.utils import geometry as G

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_geometry.npz"))


def _close(a, b, atol=2e-6):
    np.testing.assert_allclose(a.numpy() if torch.is_tensor(a) else a, b, rtol=0, atol=atol)


def test_coordinate_transforms(gold):
    c = G.pixel_coords(48, 2.0)
    _close(c, gold["coords"], 1e-7)
    tr = G.translate_coords(c, [0.11, -0.07])
    sh = G.shear_coords(tr, [0.05, -0.02])
    cm = G.compress_coords(sh, [1.1, 0.9])
    ro = G.rotate_coords(cm, 0.3)
    for got, key in ((tr, "translated"), (sh, "sheared"), (cm, "compressed"), (ro, "rotated")):
        _close(got, gold[key], 1e-6)
    _close(G.cart2polar(c), gold["polar"], 1e-6)


@pytest.mark.parametrize("inv", [False, True])
@pytest.mark.parametrize("cname", ["plain", "xf"])
def test_shapes_match_reference(gold, inv, cname):
    c = torch.as_tensor(gold["coords"] if cname == "plain" else gold["rotated"])
    clip = np.float32(2.0 / 48 * 1.5 / 2)
    fns = {
        "soft_circle": lambda: G.soft_circle(c, 0.7, clip, inv),
        "soft_square": lambda: G.soft_square(c, 1.1, clip, inv),
        "soft_rectangle": lambda: G.soft_rectangle(c, 1.3, 0.6, clip, inv),
        "soft_hexagon": lambda: G.soft_reg_polygon(c, 0.8, 6, clip, inv),
        "soft_pentagon": lambda: G.soft_reg_polygon(c, 0.75, 5, clip, inv),
        "soft_spider": lambda: G.soft_spider(c, 0.08, [0.0, 120.0, 240.0], clip, inv),
        "circle": lambda: G.circle(c, 0.7, inv),
        "square": lambda: G.square(c, 1.1, inv),
        "rectangle": lambda: G.rectangle(c, 1.3, 0.6, inv),
        "hexagon": lambda: G.reg_polygon(c, 0.8, 6, inv),
    }
    for name, fn in fns.items():
        want = gold[f"{name}_{int(inv)}_{cname}"]
        got = fn().numpy()
        if name.startswith("soft"):
            # soft edges: float32 rounding of the distance moves a value by ~1e-5 of the ramp
            np.testing.assert_allclose(got, want, rtol=0, atol=2e-4, err_msg=name)
        else:
            # hard edges: identical except for pixels whose centre sits on the edge to rounding
            assert (got != want).mean() < 2e-3, name


def test_soften_constant_support(gold):
    _close(G.soften(torch.full((4, 4), 2.0), 0.5), gold["soften_constant"])


def test_shape_parameters_are_differentiable():
    c = G.pixel_coords(32, 2.0, dtype=torch.float64)
    r = torch.tensor(0.7, dtype=torch.float64, requires_grad=True)
    t = torch.tensor([0.05, -0.02], dtype=torch.float64, requires_grad=True)
    T = G.soft_circle(G.translate_coords(c, t), r, 0.05)
    T.sum().backward()
    # area grows with the radius: d(sum T)/dr ~ circumference / pixel area
    assert float(r.grad) > 0 and torch.isfinite(t.grad).all()
    eps = 1e-6
    f = lambda rr: float(G.soft_circle(G.translate_coords(c, t.detach()), rr, 0.05).sum())
    fd = (f(0.7 + eps) - f(0.7 - eps)) / (2 * eps)
    assert abs(float(r.grad) - fd) < 1e-4 * abs(fd)


def test_zernike_basis_matches_reference(gold):
    # utils/zernikes.py:99-119, 297-315 executed by make_golden_geometry.py: Noll (n, m) table and
    # the first 15 polynomials (the reference evaluates factorials as exp(lgamma) in float32,
    # hence the 2e-5 tolerance relative to the polynomial's peak)
    from dlux_b200.utils.zernikes import noll_indices, zernike_basis
    nm = gold["noll_nm_1_21"]
    assert [tuple(r) for r in nm] == [noll_indices(j) for j in range(1, 22)]
    want = gold["zernikes_1_15"]
    got = zernike_basis(range(1, 16), gold["coords"], 2.0)
    assert got.shape == want.shape and got.dtype == np.float32
    for j in range(15):
        np.testing.assert_allclose(got[j], want[j], rtol=0, atol=2e-5 * np.abs(want[j]).max(), err_msg=f"Z{j + 1}")
    # unit rms over the disk (piston excluded): the normalisation the coefficients are quoted in
    inside = np.hypot(*gold["coords"]) <= 1.0
    for j in range(1, 15):
        assert abs(np.sqrt((got[j][inside] ** 2).mean()) - 1.0) < 0.06


def test_polike_basis_matches_reference(gold):
    # utils/zernikes.py:318-416 executed by make_golden_geometry.py: the Zernikes stretched onto 4-, 5- and 6-sided
    # polygons, plain and on transformed coordinates with another diameter; NumPy and differentiable torch forms
    import torch
    from dlux_b200.utils.zernikes import polike, polike_basis, polike_basis_torch
    for key, ns, coords, diam in (("polike_4_1_10", 4, gold["coords"], 2.0), ("polike_5_1_10", 5, gold["coords"], 2.0),
                                  ("polike_6_1_10", 6, gold["coords"], 2.0),
                                  ("polike_6_1_10_xf", 6, gold["rotated"].astype(np.float32), 1.6)):
        want = gold[key]
        got = polike_basis(ns, range(1, 11), coords, diam)
        got_t = polike_basis_torch(ns, range(1, 11), torch.as_tensor(coords, dtype=torch.float32), diam).numpy()
        assert got.shape == want.shape and got.dtype == np.float32
        for j in range(10):
            # pixels exactly on the polygon's edge (stretched radius within one float32 rounding of 1) may fall on
            # either side: allowed for a handful of pixels, and only as "inside here, outside there"
            tol = 2e-5 * np.abs(want[j]).max()
            for name, g, frac in (("numpy", got[j], 0.005), ("torch", got_t[j], 0.02)):
                bad = np.abs(g - want[j]) > tol
                assert bad.mean() < frac, (key, name, j, int(bad.sum()))
                assert np.all((g[bad] == 0) | (want[j][bad] == 0)), (key, name, j)
    with pytest.raises(ValueError):
        polike(2, 1, gold["coords"])


def test_aberrated_polygon_aperture_on_a_cpu_wavefront():
    # layers/apertures.py:643-800 with a RegPolyAperture: the aberration basis follows the aperture's transformation
    # and extent, and is the polike basis of its number of sides (polynomials.py:40-51)
    import torch
    import dlux_b200 as dl
    from dlux_b200.utils.zernikes import polike_basis_torch
    wf = dl.Wavefront(1e-6, 48, diameter=2.0, device="cpu")
    tfm = dl.CoordTransform(translation=np.array([0.05, -0.03], np.float32), rotation=np.float32(0.2))
    hexa = dl.RegPolyAperture(6, np.float32(0.8), tfm, normalise=True)
    coeffs = torch.tensor([2e-8, -1e-8, 3e-8], dtype=torch.float32, requires_grad=True)
    ab = dl.AberratedAperture(hexa, [4, 5, 6], coeffs, effect="opd")
    coords = wf.coordinates()
    want_basis = polike_basis_torch(6, [4, 5, 6], tfm(coords) / 0.8)
    np.testing.assert_allclose(ab.calc_basis(coords).detach().numpy(), want_basis.numpy(), rtol=1e-6, atol=1e-7)
    out = ab(wf)
    want = wf * hexa.transmission(coords, wf.pixel_scale)
    want = want.normalise().add_opd(torch.tensordot(coeffs.detach(), want_basis.to(torch.float32), dims=1))
    np.testing.assert_allclose(out.phasor.detach().numpy(), want.phasor.numpy(), rtol=1e-5, atol=1e-9)
    out.phasor.real.sum().backward()                     # the coefficients stay differentiable through the layer
    assert coeffs.grad is not None and float(coeffs.grad.abs().sum()) > 0
    sq = dl.AberratedAperture(dl.SquareAperture(np.float32(1.0)), [1, 2, 3])
    assert sq.calc_basis(coords).shape == (3, 48, 48)    # nsides = 4, extent = sqrt(2) * width
    with pytest.raises(TypeError):
        dl.AberratedAperture(dl.Spider(np.float32(0.05), [0.0, 90.0]), [1])


def test_aperture_layers_on_a_cpu_wavefront():
    # layers/apertures.py:134-152, 1005-1117: a dynamic aperture multiplies the wavefront by its
    # transmission on the wavefront's own coordinates; Compound = product, Multi = sum
    import dlux_b200 as dl
    wf = dl.Wavefront(1e-6, 32, diameter=2.0, device="cpu")
    primary = dl.CircularAperture(np.float32(0.8), softening=2.0)
    secondary = dl.CircularAperture(np.float32(0.2), occulting=True, softening=2.0)
    spider = dl.Spider(np.float32(0.05), [0.0, 90.0, 180.0, 270.0])
    comp = dl.CompoundAperture([("primary", primary), ("secondary", secondary), ("spider", spider)], normalise=True)
    coords, ps = wf.coordinates(), wf.pixel_scale
    want = primary.transmission(coords, ps) * secondary.transmission(coords, ps) * spider.transmission(coords, ps)
    out = comp(wf)
    np.testing.assert_allclose(out.amplitude.numpy() ** 2 / (out.amplitude.numpy() ** 2).sum(),
                               (want.numpy() ** 2) / (want.numpy() ** 2).sum(), rtol=1e-5, atol=1e-9)
    assert abs(float(out.power) - 1.0) < 1e-5 and comp.primary is primary
    holes = [dl.CircularAperture(np.float32(0.15), dl.CoordTransform(translation=np.array([x, y], np.float32)))
             for x, y in ((0.4, 0.0), (-0.3, 0.35), (0.0, -0.5))]
    multi = dl.MultiAperture(holes)
    t = multi.transmission(coords, ps)
    np.testing.assert_allclose(t.numpy(), sum(h.transmission(coords, ps) for h in holes).numpy(), rtol=1e-6)
    assert 0.99 < float(t.max()) <= 1.0 + 1e-6
    hexa = dl.RegPolyAperture(6, np.float32(0.7), dl.CoordTransform(rotation=np.float32(0.1)))
    sq = dl.SquareAperture(np.float32(1.0))
    rect = dl.RectangularAperture(np.float32(0.5), np.float32(1.2))
    for ap in (hexa, sq, rect):
        tt = ap.transmission(coords, ps)
        assert tt.shape == (32, 32) and float(tt.min()) >= 0 and float(tt.max()) <= 1 + 1e-6
    assert float(rect.transmission(coords, ps).sum()) < float(sq.transmission(coords, ps).sum())
    p = out.to_psf()                                     # wavefronts.py:281-292
    assert isinstance(p, dl.PSF) and float(p.pixel_scale) == float(out.pixel_scale)
    np.testing.assert_allclose(p.data.numpy(), out.psf.numpy())
    with pytest.raises(TypeError):
        dl.CircularAperture(0.5, transformation="shift")
    with pytest.raises(ValueError):
        dl.CircularAperture(0.5, softening=0.0)
    with pytest.raises(ValueError):
        dl.CoordTransform(translation=[1.0])
    with pytest.raises(TypeError):
        dl.CompoundAperture([object()])
