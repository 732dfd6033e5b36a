"""CUDA-graph capture of a whole fitting step through the public API.

At small sizes (BASELINE config 1: 256 -> 128 px, one wavelength = 0.1 GFLOP) a PSF + gradient is ~20 kernel
launches of a few microseconds each and the step time is Python / ctypes / launch latency, not GPU work.
``GraphedValueAndGrad`` captures ``loss_fn(*params)`` and its backward pass ONCE into a CUDA graph (the C ABI
is capture-safe: it enqueues on the given stream only, never synchronises or allocates, and passes its TMA
descriptors by value) and replays it with new parameter values -- the role ``jax.jit`` plays for the reference's
``eqx.filter_value_and_grad`` loop (docs/phase_retrieval.md:269-287)."""
from __future__ import annotations

from typing import Callable, Sequence

import torch

__all__ = ["GraphedValueAndGrad"]


class GraphedValueAndGrad:
    """``value, grads = step(*params)`` for a scalar ``loss_fn`` of CUDA tensors, replayed from one CUDA graph.

    The parameters must keep their shapes; everything else the loss closes over (optics, sources, data) must be
    unchanged between calls -- exactly the structure of an optimisation loop."""

    def __init__(self, loss_fn: Callable, params: Sequence[torch.Tensor], warmup: int = 3, has_aux: bool = False,
                 argnums: Sequence[int] | None = None):
        """``has_aux``: ``loss_fn`` returns ``(loss, aux)`` and ``aux`` (a tensor or tuple of tensors, e.g. the model
        image) is kept as a static output; ``argnums``: the parameters to differentiate (default: all)."""
        self.argnums = list(range(len(params))) if argnums is None else list(argnums)
        self.static = [p.detach().clone().requires_grad_(i in self.argnums) for i, p in enumerate(params)]
        self.has_aux = bool(has_aux)
        self.aux = None
        dev = self.static[0].device
        wrt = [self.static[i] for i in self.argnums]

        def run():
            out = loss_fn(*self.static)
            loss, aux = out if self.has_aux else (out, None)
            return loss, aux, torch.autograd.grad(loss, wrt)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):         # fills the host-side caches (uploads, scratch, geometry)
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.loss, self.aux, self.grads = run()
        torch.cuda.synchronize(dev)

    def __call__(self, *params):
        with torch.no_grad():
            for s, p in zip(self.static, params):
                if p is not s:
                    s.copy_(p)
        self.graph.replay()
        if self.has_aux:
            return (self.loss, self.aux), self.grads
        return self.loss, self.grads
