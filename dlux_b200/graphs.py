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

__all__ = ["GraphedValueAndGrad", "GraphedFitStep"]


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


class GraphedFitStep:
    """One data-fitting step ``image = model_fn(*params); loss = loss_fn(image, *data); grads = d loss / d params``
    captured into ONE CUDA graph **together with its host traffic**: the parameters and the data arrive from
    pinned host tensors and the image and the gradients leave to pinned host tensors through memcpy nodes of the
    graph.  The data (e.g. the observed image) is only needed once the model image exists and the image is only
    needed by the host after the step, so their copies sit on side branches of the graph: the host-to-device
    copy of the data runs under the forward pass and the device-to-host copy of the image under the backward
    pass, instead of bracketing the step on the compute stream.

    ``step()`` replays the graph and returns after it has completed; results are in ``host_image`` / ``host_grads``
    (and in ``.image`` / ``.grads`` on the device)."""

    def __init__(self, model_fn: Callable, loss_fn: Callable, params: Sequence[torch.Tensor],
                 data: Sequence[torch.Tensor], host_params: Sequence[torch.Tensor],
                 host_data: Sequence[torch.Tensor], host_image: torch.Tensor, host_grads: Sequence[torch.Tensor],
                 warmup: int = 3):
        for h in list(host_params) + list(host_data) + [host_image] + list(host_grads):
            if not h.is_pinned():
                raise ValueError("GraphedFitStep: host tensors must be pinned (graph memcpy nodes read / write them)")
        self.params = [p.detach().clone().requires_grad_(True) for p in params]
        self.data = [d.detach().clone() for d in data]
        dev = self.params[0].device
        cap = torch.cuda.Stream(device=dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def run():
            cur = torch.cuda.current_stream(dev)
            s_in.wait_stream(cur)
            with torch.cuda.stream(s_in):                       # branch 1: the data, under the forward pass
                for d, h in zip(self.data, host_data):
                    d.copy_(h, non_blocking=True)
            with torch.no_grad():
                for p, h in zip(self.params, host_params):
                    p.copy_(h, non_blocking=True)
            image = model_fn(*self.params)
            s_out.wait_stream(cur)
            with torch.cuda.stream(s_out):                      # branch 2: the image, under the backward pass
                host_image.copy_(image.detach(), non_blocking=True)
            cur.wait_stream(s_in)
            loss = loss_fn(image, *self.data)
            grads = torch.autograd.grad(loss, self.params)
            for g, h in zip(grads, host_grads):
                h.copy_(g, non_blocking=True)
            cur.wait_stream(s_out)
            return image, loss, grads

        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream(dev).wait_stream(cap)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=cap):
            self.image, self.loss, self.grads = run()
        torch.cuda.synchronize(dev)
        self._dev = dev

    def step(self):
        self.graph.replay()
        torch.cuda.current_stream(self._dev).synchronize()
