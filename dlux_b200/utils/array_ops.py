"""Image-plane helpers next to the path (SURVEY 8f NEXT-4): what turns the oversampled PSF the
MFT produces into detector pixels."""
from __future__ import annotations

import torch

__all__ = ["downsample"]


def downsample(array: torch.Tensor, n: int, mean: bool = True) -> torch.Tensor:
    """``dlu.downsample`` (utils/array_ops.py:124-161): n x n block mean (or sum) of a square
    array; ``PSF.downsample`` (psfs.py:74-91) uses ``mean=False``.  Leading dimensions batch."""
    if array.shape[-1] != array.shape[-2]:
        raise ValueError(f"Input array has shape {tuple(array.shape)}, which is not square")
    size_in = array.shape[-1]
    if size_in % n != 0:
        raise ValueError(f"Input array has {size_in} pixels, which is not divisible by {n}")
    size_out = size_in // n
    blocks = array.reshape(array.shape[:-2] + (size_out, n, size_out, n))
    # the reference reduces the column blocks first, then the row blocks
    out = blocks.mean(-1) if mean else blocks.sum(-1)
    return out.mean(-2) if mean else out.sum(-2)
