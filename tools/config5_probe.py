"""BASELINE config 5 shape: parameter perturbations x 32 wavelengths at 1024->256, PSF + gradient per
perturbation (a loop of fused calls; every perturbation has its own pupil).  python tools/config5_probe.py [n]"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, M, L, nz = 1024, 256, 32, 10
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
yy, xx = np.mgrid[:N, :N]
T = (np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) <= N / 2).astype(np.float32)
basis = torch.as_tensor((rng.standard_normal((nz, N, N)).astype(np.float32) * T) * np.float32(2e-8), device=dev)
wls = np.linspace(0.9e-6, 1.1e-6, L).astype(np.float32)
w = np.full(L, 1.0 / L, np.float32)
pert = torch.as_tensor(rng.standard_normal((B, nz)).astype(np.float32), device=dev)
G = torch.as_tensor(rng.standard_normal((M, M)).astype(np.float32), device=dev)
layer = dl.BasisOptic(basis, T, pert[0], normalise=True, effect="opd", device=dev)
optics = dl.AngularOpticalSystem(N, 1.0, [("a", layer)], M, 0.05, device=dev)
def sweep():
    grads = torch.empty((B, nz), device=dev)
    for b in range(B):
        c = pert[b].clone().requires_grad_(True)
        layer.coefficients = c
        psf = optics.propagate(wls, None, w)
        (psf * G).sum().backward()
        grads[b] = c.grad
    return grads
sweep(); torch.cuda.synchronize()
t0 = time.perf_counter(); g = sweep(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
flops = 4 * B * L * 8.0 * M * N * (N + M)
print(f"config 5 shape: {B} perturbations x {L} wavelengths at {N}->{M}: {dt/B*1e3:.3f} ms per perturbation "
      f"({B/dt:.0f} PSF+grad/s, {flops/dt/1e12:.0f} TFLOP/s algorithmic); 4096 would take {4096*dt/B:.1f} s on one GPU")
