"""Multi-GPU sharding of the polychromatic PSF: one process per GPU
(``torch.distributed``, NCCL over NVLink/NVSwitch; gloo in the CPU tests).

Every (source, wavelength) MFT is independent; the only coupling is the
final sum over sources and wavelengths (/root/reference/src/dLux/sources.py:409-411,
optical_systems.py:222-223).  So the (S x L) items are partitioned across ranks -- by
source when there are at least as many sources as ranks, by wavelength otherwise --
each rank runs the fused kernels on its shard, and ONE all-reduce(sum) of the [M, M]
PSF joins them.  In the backward pass every rank holds the same dL/dpsf, back-propagates
its shard, and the (small) parameter gradients are all-reduced with
:func:`all_reduce_grads`.  The reference has no distributed code at all.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["partition", "shard_sources_or_wavelengths", "all_reduce_sum", "all_reduce_grads",
           "sharded_point_sources_model"]


def partition(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced partition of range(n): the first n % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sources_or_wavelengths(n_sources: int, n_wavels: int, world: int, rank: int):
    """Returns (source_slice, wavelength_slice) owned by `rank`."""
    if n_sources >= world:
        a, b = partition(n_sources, world, rank)
        return slice(a, b), slice(0, n_wavels)
    a, b = partition(n_wavels, world, rank)
    return slice(0, n_sources), slice(a, b)


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        # every rank evaluates the same loss on the same reduced PSF, so dL/dpsf is
        # already replicated: the local shard's cotangent is g itself
        return g, None


def all_reduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSum.apply(x, group)


def all_reduce_grads(params: Sequence[torch.Tensor], group=None) -> None:
    """Sum the .grad of replicated parameters across ranks (one flat all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    params = list(params)
    if not params:
        return
    # a rank whose shard is empty has no .grad: it contributes zeros, so that every rank enters the
    # collective with the same buffer
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        p.grad.copy_(flat[off:off + p.numel()].reshape(p.shape))
        off += p.numel()


def sharded_point_sources_model(optics, wavelengths, positions, fluxes, weights=None, group=None,
                                model_fn: Optional[Callable] = None, reduce: bool = True):
    """``PointSources(wavelengths, positions, fluxes, weights).model(optics)`` with the
    (source x wavelength) items sharded over the ranks of `group`.

    ``model_fn(wavelengths, positions, weights_SL) -> psf`` defaults to the fused CUDA
    path ``optics.fused_propagate``; the CPU tests inject the oracle here.  ``reduce=False`` returns this
    rank's partial image without the all-reduce (for callers that capture the local work into a CUDA graph and
    issue the collective themselves)."""
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=np.float32))
    positions = positions if torch.is_tensor(positions) else np.asarray(positions, dtype=np.float32)
    fluxes = fluxes if torch.is_tensor(fluxes) else np.asarray(fluxes, dtype=np.float32)
    L, S = len(wavelengths), len(positions)
    if weights is None:
        weights = np.ones(L, np.float32) / np.float32(L)
    weights = np.asarray(weights, dtype=np.float32)
    weights = weights / weights.sum()                               # spectra.py:113-117
    if torch.is_tensor(fluxes):
        w_sl = torch.as_tensor(weights, device=fluxes.device)[None, :] * fluxes[:, None]
    else:
        w_sl = weights[None, :] * fluxes[:, None]                   # sources.py:398
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ss, ls = shard_sources_or_wavelengths(S, L, world, rank)
    fn = model_fn if model_fn is not None else optics.fused_propagate
    if ss.stop - ss.start == 0 or ls.stop - ls.start == 0:
        # more ranks than sources and than wavelengths: this rank owns nothing but must still take
        # part in the all-reduce (the others would block in it otherwise)
        npix = optics._focal_args()[0]
        psf = torch.zeros((npix, npix), dtype=torch.float32, device=getattr(optics, "device", "cpu"),
                          requires_grad=True)
    else:
        psf = fn(wavelengths[ls], positions[ss], w_sl[ss, ls])
    return all_reduce_sum(psf, group) if reduce else psf
