"""BASELINE config 4 at its per-GPU size: 2048 px Toliman-like diffractive pupil, 125 of the 1000
stars (the share of one of 8 GPUs) x 64 wavelengths, PSF + gradients w.r.t. OPD-basis coefficients,
star positions and fluxes through the public API.  python tools/config4_probe.py [n_stars] [n_psf]"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
S = int(sys.argv[1]) if len(sys.argv) > 1 else 125
M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N, L, nz = 2048, 64, 6
dev = torch.device("cuda:0")
rng = np.random.default_rng(4)
yy, xx = np.mgrid[:N, :N]
r = np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) / (N / 2)
T = (r <= 1).astype(np.float32)
# binary diffractive pupil: 0 / half-wave OPD cells of 64 px
cells = rng.integers(0, 2, (N // 64, N // 64)).astype(np.float32)
opd0 = np.kron(cells, np.ones((64, 64), np.float32)) * np.float32(0.5 * 585e-9) * T
basis = (rng.standard_normal((nz, N, N)).astype(np.float32) * T) * np.float32(2e-8)
wls = np.linspace(530e-9, 640e-9, L).astype(np.float32)
coeffs = torch.zeros(nz, device=dev, requires_grad=True)
pos = torch.as_tensor((rng.uniform(-1, 1, (S, 2)) * 2e-6).astype(np.float32), device=dev).requires_grad_(True)
flux = torch.as_tensor(rng.uniform(0.5, 2.0, S).astype(np.float32), device=dev).requires_grad_(True)
layer = dl.BasisOptic(basis, T, coeffs, normalise=True, effect="opd", device=dev)
layer2 = dl.Optic(None, opd0, None, device=dev)
optics = dl.AngularOpticalSystem(N, 0.125, [("mask", layer2), ("aber", layer)], M, 0.375, device=dev)
G = torch.as_tensor(rng.standard_normal((M, M)).astype(np.float32), device=dev)
def step():
    for t in (coeffs, pos, flux):
        t.grad = None
    psf = dl.PointSources(wls, pos, flux).model(optics)
    (psf * G).sum().backward()
    return psf
torch.cuda.reset_peak_memory_stats()
psf = step(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    psf = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
flops = 4 * S * L * 8.0 * M * N * (N + M)
print(f"config 4 share: {S} stars x {L} wavelengths at {N}->{M}: {dt*1e3:.1f} ms per PSF+grad step "
      f"({S*L/dt:.0f} source-wavelength MFT pairs/s, {flops/dt/1e12:.0f} TFLOP/s algorithmic), "
      f"peak memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
print("finite:", bool(torch.isfinite(psf).all()), bool(torch.isfinite(coeffs.grad).all()),
      bool(torch.isfinite(pos.grad).all()), bool(torch.isfinite(flux.grad).all()), "psf sum", float(psf.sum()), "flux sum", float(flux.sum()))
