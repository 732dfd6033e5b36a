"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the partition, the
differentiable all-reduce and the sharded PointSources orchestration.  The per-shard
compute is injected (the oracle), because the product compute path is CUDA-only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlux_b200 import distributed as D
from oracle import mft_oracle as O


def test_partition_covers_range():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = D.partition(n, world, r)
                assert 0 <= a <= b <= n
                got += list(range(a, b))
            assert got == list(range(n))
            sizes = [D.partition(n, world, r)[1] - D.partition(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.partition(4, 2, 2)
    assert D.shard_sources_or_wavelengths(8, 64, 4, 1) == (slice(2, 4), slice(0, 64))
    assert D.shard_sources_or_wavelengths(1, 64, 4, 3) == (slice(0, 1), slice(48, 64))


def _optics():
    N, M = 32, 16
    yy, xx = np.mgrid[:N, :N]
    r = np.hypot(xx - (N - 1) / 2, yy - (N - 1) / 2) / (N / 2)
    rng = np.random.default_rng(5)
    return dict(wf_npixels=N, diameter=1.0, psf_npixels=M, psf_pixel_scale=0.05, oversample=1,
                transmission=(r <= 1).astype(np.float32),
                basis=rng.standard_normal((3, N, N)).astype(np.float32) * 1e-8,
                coefficients=np.array([1.0, -2.0, 0.5], np.float32), normalise=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StubOptics:
    """What an empty shard needs from the optics: the output size and the device."""
    device = "cpu"

    def _focal_args(self):
        return 16, None, None


def _worker(rank, world, port, n_sources, q, n_wavels=4):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    od = _optics()
    wls = np.linspace(0.9e-6, 1.1e-6, n_wavels).astype(np.float32)
    rng = np.random.default_rng(1)
    pos = (rng.uniform(-1, 1, (n_sources, 2)) * 3e-7).astype(np.float32)
    flux = rng.uniform(0.5, 2.0, n_sources).astype(np.float32)
    scale = torch.tensor(2.0, requires_grad=True)      # a replicated parameter

    def model_fn(w, p, w_sl):                          # oracle stands in for the CUDA shard compute
        out = np.zeros((16, 16), np.float32)
        for s in range(len(p)):
            out += O.propagate(od, w, p[s], np.asarray(w_sl[s]))
        return torch.as_tensor(out) * scale

    psf = D.sharded_point_sources_model(_StubOptics(), wls, pos, flux, model_fn=model_fn)
    loss = (psf ** 2).sum()
    loss.backward()
    D.all_reduce_grads([scale])
    q.put((rank, psf.detach().numpy(), float(scale.grad)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_sources,n_wavels", [(1, 4), (3, 4), (1, 1)])
def test_sharded_model_matches_single_process(n_sources, n_wavels):
    # (1, 1): more ranks than sources and than wavelengths -- one rank owns an EMPTY shard and must
    # still join the all-reduce with zeros
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sources, q, n_wavels)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    od = _optics()
    wls = np.linspace(0.9e-6, 1.1e-6, n_wavels).astype(np.float32)
    rng = np.random.default_rng(1)
    pos = (rng.uniform(-1, 1, (n_sources, 2)) * 3e-7).astype(np.float32)
    flux = rng.uniform(0.5, 2.0, n_sources).astype(np.float32)
    ref = O.point_sources_model(od, wls, pos, flux) * 2.0
    for rank, psf, g in res:
        assert np.allclose(psf, ref, rtol=2e-5, atol=1e-9), rank
        # d/dscale sum((scale*P)^2) = 2*scale*sum(P^2) = 2*sum(psf^2)/scale
        assert abs(g - 2 * float((ref.astype(np.float64) ** 2).sum()) / 2.0) <= 1e-4 * abs(g)
