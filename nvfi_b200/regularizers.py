"""Fused plane regularisers: density_L1, TV_loss_density and
TV_loss_app of models/tensorf_keyframe.py:188-231 with TVLoss of utils/tensorf_utils.py:139-158,
each plane in ONE streaming CUDA pass that yields the loss term and its gradient
(csrc/regularizers.cu) instead of ~10 elementwise torch kernels and their autograd twins."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _PlaneReg(torch.autograd.Function):
    """sum over planes of one regulariser term; spec[i] = (kind, arg, scale):
    ("tv", time_plane, scale) or ("l1", offset, scale)."""

    @staticmethod
    def forward(ctx, spec, *planes):
        lib = L.load()
        dev = planes[0].device
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        need = any(ctx.needs_input_grad[1:])
        grads: List[torch.Tensor] = []
        for (kind, arg, scale), p in zip(spec, planes):
            x = p.detach()
            if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
                raise RuntimeError("nvfi_b200: plane regularisers need contiguous float32 CUDA planes")
            g = torch.empty_like(x) if need else None
            gp = g.data_ptr() if g is not None else None
            if kind == "tv":
                if x.dim() != 4 or x.shape[0] != 1:
                    raise RuntimeError("nvfi_b200: TV loss expects a (1, C, H, W) plane")
                _, C_, H_, W_ = x.shape
                rc = lib.nvfi_tv_loss(x.data_ptr(), C_, H_, W_, int(bool(arg)), float(scale), acc.data_ptr(), gp,
                                      _stream())
            else:
                rc = lib.nvfi_l1_loss(x.data_ptr(), x.numel(), float(arg), float(scale), acc.data_ptr(), gp,
                                      _stream())
            L.check(rc, "plane regulariser")
            grads.append(g)
        ctx.grads = grads
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, go):
        out = [None]
        for need, g in zip(ctx.needs_input_grad[1:], ctx.grads):
            out.append(go * g if (need and g is not None) else None)
        return tuple(out)


def density_l1(space: Sequence[torch.Tensor], time: Sequence[torch.Tensor]) -> torch.Tensor:
    """models/tensorf_keyframe.py:188-203: sum_k mean|P^s_k| + mean|1 - P^t_k|."""
    spec: List[Tuple[str, float, float]] = []
    planes: List[torch.Tensor] = []
    for s, t in zip(space, time):
        if s.shape[1] == 0:
            continue
        spec += [("l1", 0.0, 1.0), ("l1", 1.0, 1.0)]
        planes += [s, t]
    return _PlaneReg.apply(spec, *planes)


def tv_loss(space: Sequence[torch.Tensor], time: Sequence[torch.Tensor], weight: float,
            with_time: bool) -> torch.Tensor:
    """models/tensorf_keyframe.py:205-231: sum_k reg(P^s_k) 1e-2 (+ reg(P^t_k, t=True) 1e-2)."""
    spec: List[Tuple[str, float, float]] = []
    planes: List[torch.Tensor] = []
    for s, t in zip(space, time):
        spec.append(("tv", 0, weight * 1e-2))
        planes.append(s)
        if with_time:
            spec.append(("tv", 1, weight * 1e-2))
            planes.append(t)
    return _PlaneReg.apply(spec, *planes)
