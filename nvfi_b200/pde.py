"""PDE (divergence + transport) loss of the velocity field: host side of ``NVFi.get_vel_loss``
(reference: models/nvfi.py:42-84).

The reference builds the Jacobian with ``functorch.vmap(jacrev(...))`` and lets autograd
differentiate through it.  Here the occupancy filter uses the advection + density kernels
and the loss, its Jacobian and the gradients with respect to both weight nets come from
``nvfi_pde_loss`` (forward-mode tile GEMMs + a hand-written reverse pass, csrc/pde.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L
from . import engine


def _pde_params(field) -> List[torch.Tensor]:
    ps = []
    for net in (field.vel_net.weight_net, field.vel_net.a_weight_net):
        for w, b in engine.vel_linears(net):
            ps += [w, b]
    return ps


def occupancy_filter(field, points_n: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """The no-grad occupancy test of models/nvfi.py:50-64: bool mask over the points."""
    with torch.no_grad():
        K = int(field.num_keyframes)
        tsf = field.tmax / (K - 1)
        base = torch.round((t / tsf).clamp(0.0, K - 1)) * tsf
        prev = engine.integrate_pos(field.binding, points_n, t, base, group=True)   # t ~ U(0, 1) per point
        xyzt = torch.cat([prev, field.normalize_time_coord(base)], dim=-1)
        sigma = engine.density_sigma(field.binding, xyzt)
        alpha = 1 - torch.exp(-sigma * 0.01 * 25)
        return alpha >= field.alphaMask_thres


class _PdeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, field, xyzt, want_grad, *params):
        lib = L.load()
        b = field.binding
        s = b.sync()
        dev = xyzt.device
        n = xyzt.shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        va = engine.velocity(b, xyzt, True)
        sums = torch.zeros(2, device=dev, dtype=torch.float64)
        g = L.NvfiPdeGrads()
        keep = []
        gw = {"vel": [], "acc": []}
        gb = {"vel": [], "acc": []}
        if want_grad:
            for name, packed, dst_w, dst_b in (("vel", b.vel, g.g_vel_w, g.g_vel_b),
                                               ("acc", b.acc, g.g_acc_w, g.g_acc_b)):
                for j in range(L.VEL_LAYERS):
                    w_ = torch.zeros_like(packed[j].wt)
                    b_ = torch.zeros_like(packed[j].bias)
                    gw[name].append(w_)
                    gb[name].append(b_)
                    dst_w[j], dst_b[j] = w_.data_ptr(), b_.data_ptr()
            ga = torch.empty(n, 3, **f32)
            g.g_acc_pts = ga.data_ptr()
            keep.append(ga)
        ws_bytes = int(lib.nvfi_backward_workspace_bytes())
        ws = torch.empty(ws_bytes // 4, **f32)
        g.workspace, g.workspace_bytes = ws.data_ptr(), ws_bytes
        cnt = torch.empty(16, device=dev, dtype=torch.int32)
        L.check(lib.nvfi_pde_loss(C.byref(s), xyzt.data_ptr(), va.data_ptr(), n, sums.data_ptr(),
                                  C.byref(g), 1 if want_grad else 0, cnt.data_ptr(), engine._stream()),
                "pde_loss")
        loss = (5.0 * sums[0] / n + 0.1 * sums[1] / (3.0 * n)).to(torch.float32)
        grads: List[Optional[torch.Tensor]] = []
        if want_grad:
            for name, packed in (("vel", b.vel), ("acc", b.acc)):
                for j in range(L.VEL_LAYERS):
                    w_, b_ = packed[j].unpack_grad(gw[name][j], gb[name][j], True)
                    grads += [w_, b_]
        ctx.grads = grads
        ctx.needs = [p.requires_grad for p in params]
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        if not ctx.grads:
            raise RuntimeError("nvfi_b200: PDE loss was evaluated without gradients")
        out = [g_loss * gr if need else None for gr, need in zip(ctx.grads, ctx.needs)]
        return (None, None, None, *out)


def pde_loss_from_points(field, xyzt: torch.Tensor) -> torch.Tensor:
    """models/nvfi.py:69-84 on already filtered points (P,4): normalised xyz, raw t."""
    xyzt = xyzt.detach().reshape(-1, 4).contiguous().float()
    params = _pde_params(field)
    want = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _PdeFn.apply(field, xyzt, want, *params)


def vel_loss(field, n_pts: int, points: Optional[torch.Tensor] = None, t: Optional[torch.Tensor] = None):
    """NVFi.get_vel_loss (models/nvfi.py:42-84).  Returns python 0.0 when no point is occupied
    (the reference's "nothing to do" signal, :66-67)."""
    if not field.use_vel:
        raise RuntimeError("nvfi_b200: get_vel_loss needs use_vel")
    dev = field.aabb.device
    if points is None:
        lo, hi = field.aabb
        points = field.normalize_coord(torch.rand(int(n_pts), 3, device=dev) * (hi - lo) + lo)
    if t is None:
        t = torch.rand(points.shape[0], 1, device=dev)
    points = points.detach().to(dev).reshape(-1, 3).float()
    t = t.detach().to(dev).reshape(-1, 1).float()
    keep = occupancy_filter(field, points, t)
    xyzt = torch.cat([points, t], dim=-1)[keep]
    if xyzt.shape[0] == 0:
        return 0.
    return pde_loss_from_points(field, xyzt)
