import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import dlux_b200 as dl
from oracle import mft_oracle as O
from test_gpu_parity import _optics_dict
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
x = ((rng.standard_normal((2, 130, 130)) + 1j*rng.standard_normal((2, 130, 130)))/130).astype(np.complex64)
out = dl.utils.MFT(torch.as_tensor(x, device=dev), np.array([1e-6, 1.1e-6], np.float32), np.float32(1/130), 71, np.float32(2e-7))
ref = O.MFT(x[1], 1.1e-6, 1/130, 71, 2e-7)
print('mft err', np.linalg.norm(out[1].cpu().numpy()-ref)/np.linalg.norm(ref))
od = _optics_dict(96, 48, 3, 1)
c = torch.as_tensor(od['coefficients'], device=dev).requires_grad_(True)
layer = dl.BasisOptic(od['basis'], od['transmission'], c, normalise=True, effect="opd", device=dev)
s = dl.AngularOpticalSystem(96, 1.0, [('a', layer)], 48, 0.05, device=dev)
pos = torch.as_tensor(np.array([[1e-7, -2e-7], [0, 0]], np.float32), device=dev).requires_grad_(True)
psf = s.model(dl.PointSources(np.linspace(0.9e-6, 1.1e-6, 3).astype(np.float32), pos, np.array([1.0, 2.0], np.float32)))
psf.sum().backward(); torch.cuda.synchronize()
print('ok', float(psf.sum()), float(c.grad.abs().sum()), float(pos.grad.abs().sum()))
