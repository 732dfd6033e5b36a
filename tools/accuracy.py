"""Prints the relative L2 error of the CUDA path against the oracle (float32 restatement and
its float64 twin) for a few shapes; run on the GPU box: python tools/accuracy.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
from oracle import mft_oracle as O

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
print(f"{'shape':>14s} {'prec':>7s} {'vs f32 oracle':>14s} {'vs f64 oracle':>14s} {'f32 vs f64 oracle':>18s}")
for n_in, n_out, wl, D, pscale in [(256, 128, 1e-6, 1.0, 0.05), (512, 256, 1e-6, 1.0, 0.025),
                                   (1024, 512, 4.3e-6, 6.6, 0.0656 / 4), (2048, 256, 5.85e-7, 0.125, 0.7)]:
    x = ((rng.standard_normal((n_in, n_in)) + 1j * rng.standard_normal((n_in, n_in))) / n_in).astype(np.complex64)
    ps_in = np.float32(D / n_in)
    pso = O.arcsec2rad(pscale)
    r32 = O.MFT(x, wl, ps_in, n_out, pso)
    r64 = O.MFT(x, wl, ps_in, n_out, pso, dtype=np.float64)
    for prec in ("fp32", "3xtf32"):
        out = dl.utils.MFT(torch.as_tensor(x, device=dev), np.float32(wl), ps_in, n_out, pso, precision=prec).cpu().numpy()
        print(f"{n_in:6d}->{n_out:<6d} {prec:>7s} {rel(out, r32):14.3e} {rel(out, r64):14.3e} {rel(r32, r64):18.3e}")
