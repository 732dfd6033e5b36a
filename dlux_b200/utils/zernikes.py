"""Zernike polynomials on a pupil grid (Noll order, unit-rms normalisation): the OPD basis the
BASELINE configurations C1 / C2 put behind ``BasisOptic``.  A setup-time producer, evaluated once
in float64 and stored as float32; mirrors ``zernike_basis`` of /root/reference/src/dLux/utils/
zernikes.py:297-315 (``noll_indices`` :99-119, radial polynomial :122-176, azimuthal part and
normalisation :179-204, aperture support rho <= 1 :227-254), pinned by tests/golden/
reference_geometry.npz."""
from __future__ import annotations

import math

import numpy as np

__all__ = ["noll_indices", "zernike", "zernike_basis", "zernike_basis_torch", "polike", "polike_basis",
           "polike_basis_torch"]


def noll_indices(j: int):
    """Noll index j >= 1 -> (n, m); m < 0 are the sine terms (odd j)."""
    if j < 1:
        raise ValueError("The Zernike index must be greater than 0.")
    n = int(math.ceil((-1 + math.sqrt(1 + 8 * j)) / 2) - 1)
    first = n * (n + 1) // 2 + 1                  # smallest j of radial order n
    # |m| values of row n in Noll order: n even: 0, 2, 2, 4, 4, ...; n odd: 1, 1, 3, 3, ...
    idx = j - first
    m_abs = 2 * ((idx + 1) // 2) if n % 2 == 0 else 2 * (idx // 2) + 1
    if m_abs == 0:
        return n, 0
    return n, (m_abs if j % 2 == 0 else -m_abs)    # even j: cosine term


def zernike(j: int, coordinates, diameter: float = 2.0):
    """Z_j on cartesian `coordinates` [2, npix, npix] (x first), zero outside the unit disk of
    the given diameter."""
    # the support test rho <= 1 is made on float32 radii, as the reference's float32 arrays make it (pixels that
    # sit exactly on the edge -- every polygon edge pixel of a symmetric grid -- would otherwise fall outside);
    # the polynomial itself is evaluated in float64 and rounded once
    c32 = np.asarray(coordinates, dtype=np.float32) / np.float32(float(diameter) / 2)
    inside = np.hypot(c32[0], c32[1]) <= np.float32(1.0)
    c = c32.astype(np.float64)
    rho, theta = np.hypot(c[0], c[1]), np.arctan2(c[1], c[0])
    n, m = noll_indices(j)
    ma = abs(m)
    radial = np.zeros_like(rho)
    for k in range((n - ma) // 2 + 1):
        coef = ((-1) ** k * math.factorial(n - k)
                / (math.factorial(k) * math.factorial((n + ma) // 2 - k) * math.factorial((n - ma) // 2 - k)))
        radial += coef * rho ** (n - 2 * k)
    norm = math.sqrt(n + 1) * (math.sqrt(2) if m != 0 else 1.0)
    az = np.cos(ma * theta) if m >= 0 else np.sin(ma * theta)
    return (inside * radial * norm * az).astype(np.float32)


def zernike_basis(js, coordinates, diameter: float = 2.0):
    return np.stack([zernike(int(j), coordinates, diameter) for j in js])


def zernike_basis_torch(js, coordinates, diameter: float = 2.0):
    """The same basis as differentiable torch arithmetic on a (possibly transformed) coordinate tensor
    [2, npix, npix]: what ``DynamicZernikeBasis.calculate_basis`` evaluates inside ``AberratedAperture``
    (/root/reference/src/dLux/polynomials.py:85-115, utils/zernikes.py:179-254)."""
    import torch
    c = coordinates / (float(diameter) / 2)
    rho, theta = torch.hypot(c[0], c[1]), torch.atan2(c[1], c[0])
    inside = (rho <= 1.0).to(c.dtype)
    out = []
    for j in js:
        n, m = noll_indices(int(j))
        ma = abs(m)
        radial = torch.zeros_like(rho)
        for k in range((n - ma) // 2 + 1):
            coef = ((-1) ** k * math.factorial(n - k)
                    / (math.factorial(k) * math.factorial((n + ma) // 2 - k) * math.factorial((n - ma) // 2 - k)))
            radial = radial + coef * rho ** (n - 2 * k)
        norm = math.sqrt(n + 1) * (math.sqrt(2) if m != 0 else 1.0)
        az = torch.cos(ma * theta) if m >= 0 else torch.sin(ma * theta)
        out.append(inside * radial * norm * az)
    return torch.stack(out)


def _polike_radius(theta, nsides: int, xp):
    """r_alpha(theta): the polygon's edge distance along direction theta in units of its circumradius
    (/root/reference/src/dLux/utils/zernikes.py:343-348, 390-394): the Zernikes are evaluated on coordinates
    stretched by 1 / r_alpha so that the unit disk maps onto the n-sided polygon."""
    alpha = math.pi / nsides
    phi = theta + alpha
    wedge = xp.floor((phi + alpha) / (2.0 * alpha))
    u_alpha = phi - wedge * (2 * alpha)
    return math.cos(alpha) / xp.cos(u_alpha)


def polike(nsides: int, j: int, coordinates, diameter: float = 2.0):
    """Z_j on an n-sided regular polygon ("polike", utils/zernikes.py:318-349): ``1 / r_alpha * Z_j(c / r_alpha)``."""
    if nsides < 3:
        raise ValueError(f"nsides must be >= 3, not {nsides}.")
    # float32 throughout the coordinate stretch, in the reference's order of operations: which edge pixels are
    # inside depends on these roundings
    F = np.float32
    c = np.asarray(coordinates, dtype=np.float32) / F(float(diameter) / 2)
    alpha = math.pi / nsides
    phi = np.arctan2(c[1], c[0]) + F(alpha)
    wedge = np.floor((phi + F(alpha)) / F(2.0 * alpha))
    u_alpha = phi - wedge * F(2 * alpha)
    r_alpha = F(math.cos(alpha)) / np.cos(u_alpha)
    return (F(1) / r_alpha * zernike(j, c / r_alpha, 2.0)).astype(np.float32)


def polike_basis(nsides: int, js, coordinates, diameter: float = 2.0):
    return np.stack([polike(nsides, int(j), coordinates, diameter) for j in js])


def polike_basis_torch(nsides: int, js, coordinates, diameter: float = 2.0):
    """Differentiable torch form on a (possibly transformed) coordinate tensor: what ``DynamicZernikeBasis``
    evaluates for an aperture with ``nsides > 0`` (/root/reference/src/dLux/polynomials.py:40-51 ->
    ``polike_fast``, utils/zernikes.py:352-395)."""
    import torch
    if nsides < 3:
        raise ValueError(f"nsides must be >= 3, not {nsides}.")
    c = coordinates / (float(diameter) / 2)
    r_alpha = _polike_radius(torch.atan2(c[1], c[0]), nsides, torch)
    return 1.0 / r_alpha * zernike_basis_torch(js, c / r_alpha, 2.0)

