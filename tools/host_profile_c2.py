"""cProfile of the public-API step at config 2 (host-bound)."""
import cProfile, pstats, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
from dlux_b200 import workloads
dev = torch.device("cuda:0")
cfg = workloads.config("c2")
N = cfg["wf_npixels"]
c = torch.as_tensor(cfg["coefficients"], device=dev).requires_grad_(True)
layer = dl.BasisOptic(torch.as_tensor(cfg["basis"], device=dev), torch.as_tensor(cfg["transmission"], device=dev), c, normalise=True, effect="opd", device=dev)
optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("p", layer)], cfg["psf_npixels"], cfg["psf_pixel_scale"], cfg["oversample"], device=dev)
G = torch.as_tensor(cfg["G"], device=dev)
src = dl.PointSource(cfg["wavelengths"], cfg["positions"][0], 1.0, cfg["weights"])
def step():
    c.grad = None
    psf = src.model(optics)
    (psf * G).sum().backward()
for _ in range(20): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(500): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
