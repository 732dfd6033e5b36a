"""BASELINE config 4 at full size: 2048 px binary-phase-mask pupil, 1000 stars x 64 wavelengths -> 256x256,
PSF + gradients w.r.t. the star positions, fluxes and the phase mask, stars sharded over the GPUs of one box
with one NCCL all-reduce each way (bench.run_c4_strong with all 1000 stars).

    python tools/config4_full.py [n_stars]                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_full.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    out = bench.run_c4_strong(dev, world, rank, steps=2, n_stars=n)
    out["n_gpus"] = world
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
