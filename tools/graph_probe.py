"""Eager launches vs one CUDA graph for the device-resident C3 step (forward + gradient)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlux_b200 import ops, workloads, _lib
from dlux_b200.utils import propagation as P
dev = torch.device("cuda:0"); _lib.load()
cfg = workloads.config("c3")
N, M = cfg["wf_npixels"], cfg["psf_npixels"] * cfg["oversample"]; L = len(cfg["wavelengths"])
up = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
T_d, basis_d, G_d, coeffs_d = up(cfg["transmission"]), up(cfg["basis"]), up(cfg["G"]), up(cfg["coefficients"])
ps_in = np.float32(np.float32(cfg["diameter"]) / np.float32(N))
ps_out = P.arcsec2rad(np.float32(cfg["psf_pixel_scale"]) / np.float32(cfg["oversample"]))
s_h, nrm_h = P.mft_geometry(cfg["wavelengths"], N, ps_in, M, ps_out)
k_d = up((np.float32(2 * np.pi) / cfg["wavelengths"]).astype(np.float32))
s_d, nrm_d = up(s_h.astype(np.float32)), up(nrm_h.astype(np.float32))
w_d = up(cfg["weights"].astype(np.float32)).reshape(1, L)
delta_d = torch.zeros((1, L, 2), device=dev)
def step():
    opd = ops.basis_eval(basis_d, coeffs_d)
    psf, field = ops.polypsf_fwd(T_d, opd, None, k_d, s_d, nrm_d, w_d, delta_d, N, M, True, None, True)
    opd_bar = ops.polypsf_bwd(T_d, opd, None, k_d, s_d, nrm_d, w_d, delta_d, field, G_d, N, M, True, None, True, False, False)[0]
    return psf, ops.basis_reduce(basis_d, opd_bar, coeffs_d.shape)
def timed(fn, n=200):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for _ in range(5): step()
print("eager  %.4f ms/step" % timed(step))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    out = step()
torch.cuda.synchronize()
ref = step()
g.replay(); torch.cuda.synchronize()
print("graph == eager:", bool(torch.equal(out[0], ref[0])), bool(torch.allclose(out[1], ref[1], rtol=1e-5)))
print("graph  %.4f ms/step" % timed(g.replay))
print("eager  %.4f ms/step" % timed(step))
