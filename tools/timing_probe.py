"""In-kernel cycle counters (build with DLUX_NVCC_EXTRA=-DDLUX_DEBUG_TIMING): forward MFT
1024->512 with `batch` items; prints what block 0 measured."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import dlux_b200 as dl
dev = torch.device('cuda:0')
for batch in [int(a) for a in sys.argv[1:]] or [1, 64]:
    rng = np.random.default_rng(0)
    x = torch.as_tensor((rng.standard_normal((batch, 1024, 1024)) + 1j * rng.standard_normal((batch, 1024, 1024))).astype(np.complex64) / 1024, device=dev)
    wl = np.linspace(4.1e-6, 4.5e-6, batch).astype(np.float32)
    print(f"==== batch {batch}", flush=True)
    for rep in range(2):
        out = dl.utils.MFT(x, wl, np.float32(6.6 / 1024), 512, np.float32(8e-8))
        torch.cuda.synchronize()
        print("----", flush=True)
