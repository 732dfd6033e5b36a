"""BASELINE config 5 (Fisher / mask-design sweep): 4096 coefficient perturbations x 32 wavelengths at
1024 -> 256 px, per-item PSF + per-item coefficient gradient through ONE fused call per direction
(``propagate_batch`` -> dlux_polypsf_batch_fwd / _bwd).  The batch is sharded batch-major over the
ranks (no collective: outputs are per item).

    python tools/config5_probe.py [B_per_gpu] [--loop]          # 1 GPU
    torchrun --nproc-per-node N tools/config5_probe.py [B_per_gpu]

Prints one JSON line (rank 0): items/s over all ranks (max-over-ranks device time), per-GPU TFLOP/s, and
--loop adds the round-1 figure (a Python loop of single fused calls) measured on the same box.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlux_b200 as dl  # noqa: E402
from dlux_b200 import workloads  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    loop = "--loop" in sys.argv
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = workloads.config("c5")
    B = int(args[0]) if args else 256
    N, M, L = cfg["wf_npixels"], cfg["psf_npixels"], len(cfg["wavelengths"])
    lo = (rank * B) % 4096
    pert = torch.as_tensor(np.roll(cfg["perturbations"], -lo, 0)[:B].copy(), device=dev)
    basis = torch.as_tensor(cfg["basis"], device=dev)
    T = torch.as_tensor(cfg["transmission"], device=dev)
    G = torch.as_tensor(cfg["G"], device=dev)
    layer = dl.BasisOptic(basis, T, pert[0], normalise=True, effect="opd", device=dev)
    optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("pupil", layer)], cfg["psf_npixels"],
                                     cfg["psf_pixel_scale"], cfg["oversample"], device=dev)

    def step():
        c = pert.clone().requires_grad_(True)
        psfs = optics.propagate_batch(c, cfg["wavelengths"], weights=cfg["weights"])
        (psfs * G[None]).sum().backward()
        return psfs, c.grad

    def timed(fn, n):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    step()
    step()
    ms = timed(step, 3)
    flops = 2 * L * 8.0 * M * N * (N + M) * B
    out = {"config": "c5: 1024->256, 32 wavelengths, per-item PSF + coefficient gradient", "n_gpus": world,
           "batch_per_gpu": B, "ms_per_batch": ms, "items_per_s": world * B * 1e3 / ms,
           "tflops_algorithmic_per_gpu": flops / (ms * 1e-3) / 1e12,
           "seconds_for_4096": 4096.0 / (world * B * 1e3 / ms), "collective": "none (per-item outputs)",
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    if loop and world == 1:
        nl = min(B, 32)

        def loop_step():
            for b in range(nl):
                c = pert[b].clone().requires_grad_(True)
                layer.coefficients = c
                psf = optics.propagate(cfg["wavelengths"], None, cfg["weights"])
                (psf * G).sum().backward()
        loop_step()
        msl = timed(loop_step, 2) / nl
        out["loop_of_single_calls_items_per_s"] = 1e3 / msl
        out["speedup_vs_loop"] = out["items_per_s"] / (1e3 / msl)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
