"""Host-side profile of the public-API step (the bench's e2e arm): where the CPU time of one
PSF+gradient goes.  python tools/e2e_profile.py"""
import cProfile, pstats, sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
from dlux_b200 import workloads, distributed as D
cfg = workloads.config("c3")
dev = torch.device("cuda:0")
N, M = cfg["wf_npixels"], cfg["psf_npixels"] * cfg["oversample"]
basis_d = torch.as_tensor(cfg["basis"], device=dev); T_d = torch.as_tensor(cfg["transmission"], device=dev)
coeffs_h = torch.as_tensor(cfg["coefficients"]).pin_memory(); G_h = torch.as_tensor(cfg["G"]).pin_memory()
psf_h = torch.empty((M, M)).pin_memory(); grad_h = torch.empty(len(cfg["coefficients"])).pin_memory()
layer = dl.BasisOptic(basis_d, T_d, torch.as_tensor(cfg["coefficients"], device=dev), normalise=True, effect="opd", device=dev)
optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("pupil", layer)], cfg["psf_npixels"], cfg["psf_pixel_scale"], cfg["oversample"], device=dev)
def step(sync=True):
    c = coeffs_h.to(dev, non_blocking=True).requires_grad_(True)
    G = G_h.to(dev, non_blocking=True)
    layer.coefficients = c
    psf = D.sharded_point_sources_model(optics, cfg["wavelengths"], cfg["positions"], cfg["fluxes"], cfg["weights"])
    (psf * G).sum().backward()
    psf_h.copy_(psf.detach(), non_blocking=True); grad_h.copy_(c.grad, non_blocking=True)
    if sync: torch.cuda.current_stream().synchronize()
for _ in range(10): step()
t0 = time.perf_counter()
for _ in range(200): step()
print("sync every step: %.3f ms/step" % ((time.perf_counter() - t0) * 5))
t0 = time.perf_counter()
for _ in range(200): step(False)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("no sync: host enqueue %.3f ms/step, total %.3f ms/step" % ((t1 - t0) * 5, (time.perf_counter() - t0) * 5))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): step()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
