"""BASELINE configs 1 and 2 through the public API on one GPU (PSF + gradient w.r.t. the Zernike
coefficients), fused route."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
from dlux_b200 import workloads
dev = torch.device("cuda:0")
for name in ("c1", "c2"):
    cfg = workloads.config(name)
    N, M = cfg["wf_npixels"], cfg["psf_npixels"] * cfg["oversample"]
    c = torch.as_tensor(cfg["coefficients"], device=dev).requires_grad_(True)
    layer = dl.BasisOptic(torch.as_tensor(cfg["basis"], device=dev), torch.as_tensor(cfg["transmission"], device=dev),
                          c, normalise=True, effect="opd", device=dev)
    optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("p", layer)], cfg["psf_npixels"], cfg["psf_pixel_scale"],
                                     cfg["oversample"], device=dev)
    G = torch.as_tensor(cfg["G"], device=dev)
    src = dl.PointSource(cfg["wavelengths"], cfg["positions"][0], 1.0, cfg["weights"])
    def step():
        c.grad = None
        layer.coefficients = c
        psf = src.model(optics)
        (psf * G).sum().backward()
    for _ in range(5): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 200
    for _ in range(n): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    L = len(cfg["wavelengths"])
    print(f"{name}: {N}->{M}, {L} wavelength(s): {dt*1e3:.3f} ms per PSF+grad ({1/dt:.0f}/s, "
          f"{4*L*8.0*M*N*(N+M)/dt/1e12:.1f} TFLOP/s algorithmic)")
    # the same step captured once into a CUDA graph (dlux_b200.GraphedValueAndGrad)
    def loss_fn(cc):
        layer.coefficients = cc
        return (src.model(optics) * G).sum()
    gstep = dl.GraphedValueAndGrad(loss_fn, [c.detach()])
    val, grads = gstep(c.detach())
    torch.cuda.synchronize()
    step()
    print("   graph == eager gradient:", bool(torch.allclose(grads[0], c.grad, rtol=1e-5, atol=0)))
    t0 = time.perf_counter()
    for _ in range(n): gstep(c.detach())
    torch.cuda.synchronize(); dtg = (time.perf_counter() - t0) / n
    print(f"   CUDA graph: {dtg*1e3:.3f} ms per PSF+grad ({1/dtg:.0f}/s)")
