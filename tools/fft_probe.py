"""`dlu.FFT` through the MFT kernels (O(N^2 N_pad), tensor cores) against torch.fft.fft2 (radix FFT) on the same
box: tells a user when the `FFT` layer's MFT route loses."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl
dev = torch.device("cuda:0")
rows = []
for n, pad in ((256, 2), (512, 2), (1024, 2), (1024, 1)):
    x = torch.randn(n, n, dtype=torch.complex64, device=dev)
    def ours():
        return dl.utils.FFT(x, np.float32(1e-6), np.float32(0.01), None, pad, False)[0]
    def ref():
        p = (n * (pad - 1)) // 2
        return torch.fft.fftshift(torch.fft.fft2(torch.fft.ifftshift(torch.nn.functional.pad(x, (p, p, p, p))))) / (n * pad)
    out = {}
    for name, fn in (("mft_route_ms", ours), ("torch_fft_ms", ref)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 20
    err = float((ours() - ref()).abs().max() / ref().abs().max())
    rows.append(dict(n=n, pad=pad, **out, max_rel_diff=err))
print(json.dumps(rows))
