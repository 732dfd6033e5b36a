"""C3 through the layer-by-layer route (fused=False): one batched wavefront, torch elementwise
pupil ops, batched dlux_mft_c64 + its adjoint -- against the fused route."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
from dlux_b200 import workloads
cfg = workloads.config("c3"); dev = torch.device("cuda:0")
N, M = cfg["wf_npixels"], cfg["psf_npixels"] * cfg["oversample"]
basis_d = torch.as_tensor(cfg["basis"], device=dev); T_d = torch.as_tensor(cfg["transmission"], device=dev)
G = torch.as_tensor(cfg["G"], device=dev)
for fused in (True, False):
    c = torch.as_tensor(cfg["coefficients"], device=dev).requires_grad_(True)
    layer = dl.BasisOptic(basis_d, T_d, c, normalise=True, effect="opd", device=dev)
    optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("p", layer)], cfg["psf_npixels"], cfg["psf_pixel_scale"],
                                     cfg["oversample"], device=dev, fused=fused)
    def step():
        c.grad = None
        psf = optics.propagate(cfg["wavelengths"], None, cfg["weights"])
        (psf * G).sum().backward()
        return psf
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): p = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"fused={fused}: {dt*1e3:.2f} ms per PSF+grad, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB, psf sum {float(p.sum()):.6f}, grad0 {float(c.grad[0]):.6e}")
