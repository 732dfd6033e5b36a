"""Forward-only C3 PSF (no gradient): the EPI_PSF epilogue against the field + reduce path."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
from dlux_b200 import workloads
dev = torch.device("cuda:0")
cfg = workloads.config("c3")
layer = dl.BasisOptic(torch.as_tensor(cfg["basis"], device=dev), torch.as_tensor(cfg["transmission"], device=dev),
                      torch.as_tensor(cfg["coefficients"], device=dev), normalise=True, effect="opd", device=dev)
optics = dl.AngularOpticalSystem(cfg["wf_npixels"], cfg["diameter"], [("p", layer)], cfg["psf_npixels"], cfg["psf_pixel_scale"], cfg["oversample"], device=dev)
out = {}
for name, env in (("epi_psf", None), ("field_reduce", "1")):
    if env: os.environ["DLUX_B200_NO_EPI_PSF"] = env
    for _ in range(5): optics.propagate(cfg["wavelengths"], None, cfg["weights"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): optics.propagate(cfg["wavelengths"], None, cfg["weights"])
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / 50
    os.environ.pop("DLUX_B200_NO_EPI_PSF", None)
out["forward_only_psf_per_s"] = 1e3 / out["epi_psf_ms"]
print(json.dumps(out))
