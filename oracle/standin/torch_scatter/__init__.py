"""CPU stand-in for the third-party ``torch_scatter`` package (TEST INFRASTRUCTURE ONLY).

The reference imports ``scatter_add`` and ``scatter_max`` from ``torch_scatter``
(reference ``relgat_projector/core/model/layer.py:6``; call sites ``:284, :290,
:308, :316``).  The dependency is un-pinned (``requirements.txt:7``,
``setup.py:41``), is not vendored under ``/root/reference`` and is not installed
in this image, so its two functions are restated here from the published
behaviour of rusty1s/pytorch_scatter (2.1.x):

* ``scatter_add`` / ``scatter_sum``: broadcast ``index`` to ``src``'s shape along
  ``dim``, allocate zeros of ``dim_size`` and ``scatter_add_`` into them.
* ``scatter_max``: segment maximum; segments that receive no element are 0 (the
  package fills with the dtype's lowest value, then zeroes untouched slots);
  ``arg`` is ``src.size(dim)`` for empty segments.  Its backward routes the
  gradient to the arg-max element only.

Nothing in the product package imports this module: it exists so that
``oracle/gen_golden.py`` and the CPU baseline can execute the reference's own
Python files verbatim.  Parity is therefore "unpinned" at this boundary (the
reference ships no test that fixes torch_scatter's results); see DESIGN.md.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def _broadcast(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src: torch.Tensor, index: torch.Tensor, dim: int, dim_size: Optional[int]):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = int(dim_size)
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None):
    index = _broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, dim, size):
        lowest = torch.finfo(src.dtype).min if src.dtype.is_floating_point else torch.iinfo(src.dtype).min
        out = torch.full(size, lowest, dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(dim, index, src, reduce="amax", include_self=True)
        # arg-max: first position (in edge order) that attains the maximum
        n = src.size(dim)
        pos_shape = [1] * src.dim()
        pos_shape[dim] = n
        pos = torch.arange(n, device=src.device).view(pos_shape).expand_as(src)
        hit = src == out.gather(dim, index)
        cand = torch.where(hit, pos, torch.full_like(pos, n))
        arg = torch.full(size, n, dtype=torch.long, device=src.device)
        arg = arg.scatter_reduce(dim, index, cand, reduce="amin", include_self=True)
        out = out.masked_fill(arg == n, 0)
        ctx.save_for_backward(arg)
        ctx.dim = dim
        ctx.src_shape = list(src.shape)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        dim = ctx.dim
        shape = list(ctx.src_shape)
        shape[dim] += 1  # slot n swallows the empty segments
        grad_src = torch.zeros(shape, dtype=grad_out.dtype, device=grad_out.device)
        grad_src.scatter_(dim, arg, grad_out)
        grad_src = grad_src.narrow(dim, 0, shape[dim] - 1)
        return grad_src, None, None, None


def scatter_max(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    if out is not None:
        raise NotImplementedError("stand-in: out= is not used by the reference")
    index = _broadcast(index, src, dim)
    if dim < 0:
        dim = src.dim() + dim
    return _ScatterMax.apply(src, index, dim, _out_size(src, index, dim, dim_size))


__all__ = ["scatter_add", "scatter_sum", "scatter_max"]
