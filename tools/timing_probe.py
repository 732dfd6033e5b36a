"""In-kernel cycle counters (build with DLUX_NVCC_EXTRA=-DDLUX_DEBUG_TIMING): one batched MFT n_in -> n_out
(default 1024 -> 512, the forward shape of config 3; 512 -> 1024 is its adjoint's shape); prints what CTA 0 measured.

    python tools/timing_probe.py [batch [n_in n_out]]"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import dlux_b200 as dl
dev = torch.device('cuda:0')
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_in, n_out = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 512)
rng = np.random.default_rng(0)
x = torch.as_tensor((rng.standard_normal((batch, n_in, n_in)) + 1j * rng.standard_normal((batch, n_in, n_in))).astype(np.complex64) / n_in, device=dev)
wl = np.linspace(4.1e-6, 4.5e-6, batch).astype(np.float32)
print(f"==== batch {batch}  {n_in} -> {n_out}", flush=True)
for rep in range(2):
    out = dl.utils.MFT(x, wl, np.float32(6.6 / n_in), n_out, np.float32(8e-8 * 512 / n_out))
    torch.cuda.synchronize()
    print("----", flush=True)
