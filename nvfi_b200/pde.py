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
        nets = (("vel", b.vel, g.g_vel_w, g.g_vel_b), ("acc", b.acc, g.g_acc_w, g.g_acc_b))
        outs = {}
        if want_grad:
            # ONE zero-filled buffer for the 24 packed accumulators, ONE for the gradients in nn.Linear layout
            al = lambda x: (x + 31) // 32 * 32
            sizes = []
            for name, packed, _, _ in nets:
                for j in range(L.VEL_LAYERS):
                    sizes.append((packed[j].wt.numel(), (packed[j].out_dim, packed[j].in_dim)))
                    sizes.append((packed[j].bias.numel(), (packed[j].out_dim,)))
            import math
            acc = torch.zeros(sum(al(n_) for n_, _ in sizes), **f32)
            outbuf = torch.empty(sum(al(math.prod(sh)) for _, sh in sizes), **f32)
            P = L.NvfiPdeParamGrads()
            po = oo = 0
            k = 0
            for name, packed, dst_w, dst_b in nets:
                pw, pb = (P.vel_w, P.vel_b) if name == "vel" else (P.acc_w, P.acc_b)
                outs[name] = []
                for j in range(L.VEL_LAYERS):
                    for dst, pdst in ((dst_w, pw), (dst_b, pb)):
                        n_, sh = sizes[k]
                        k += 1
                        dst[j] = acc[po:po + n_].data_ptr()
                        o_ = outbuf[oo:oo + math.prod(sh)].view(sh)
                        pdst[j] = o_.data_ptr()
                        outs[name].append(o_)
                        po += al(n_)
                        oo += al(math.prod(sh))
            ga = torch.empty(n, 3, **f32)
            g.g_acc_pts = ga.data_ptr()
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
            L.check(lib.nvfi_unpack_pde_grads(C.byref(s), C.byref(g), C.byref(P), engine._stream()), "unpack_pde_grads")
            grads = outs["vel"] + outs["acc"]
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
