"""Autograd wiring of the fused render: one ``torch.autograd.Function`` whose forward is
``nvfi_render_forward`` and whose backward is ``nvfi_render_backward``.

Gradients are produced for exactly the tensors the reference's autograd reaches from a
render: the 12 factor planes, ``basis_mat``, the render
MLP and — through the sample coordinates — ``vel_net.weight_net``.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import engine


def _diff_params(field) -> List[torch.Tensor]:
    ps = list(field.density_plane_space) + list(field.density_plane_time)
    ps += list(field.app_plane_space) + list(field.app_plane_time)
    ps.append(field.basis_mat.weight)
    if field.shadingMode == "MLP_PE":
        mlp = field.renderModule.mlp
        for i in (0, 2, 4):
            ps += [mlp[i].weight, mlp[i].bias]
    if field.use_vel:
        for w, b in engine.vel_linears(field.vel_net.weight_net):
            ps += [w, b]
    return ps


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, field, call, *params):
        out = engine.render_forward(field.binding, call["o"], call["d"], call["t"],
                                    white_bg=call["white_bg"], training=call["training"],
                                    jitter=call["jitter"], chunk_bg=call["chunk_bg"],
                                    transfer_vel=call["transfer_vel"], ray_chunk=call["ray_chunk"],
                                    save_sigma=True)
        ctx.field = field
        ctx.out = out
        ctx.n_params = len(params)
        ctx.versions = [p._version for p in params]
        ctx.params = params
        ctx.mark_non_differentiable(out.mask_map)
        return out.rgb_map, out.depth_map, out.acc_map, out.weights, out.mask_map

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc, g_w, _g_mask):
        if any(p._version != v for p, v in zip(ctx.params, ctx.versions)):
            raise RuntimeError("nvfi_b200: a parameter was modified between render forward and backward")
        grads = engine.render_backward(ctx.field.binding, ctx.out, g_rgb, g_depth, g_acc, g_w,
                                       [p.requires_grad for p in ctx.params])
        ctx.out = None
        return (None, None, *grads)


def render_with_grad(field, t, ray_o, ray_d, white_bg, training, jitter, chunk_bg, transfer_vel,
                     ray_chunk):
    call = dict(o=ray_o, d=ray_d, t=t, white_bg=white_bg, training=training, jitter=jitter,
                chunk_bg=chunk_bg, transfer_vel=transfer_vel, ray_chunk=ray_chunk)
    params = _diff_params(field)
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _RenderFn.apply(field, call, *params)
    out = engine.render_forward(field.binding, ray_o, ray_d, t, white_bg=white_bg, training=training,
                                jitter=jitter, chunk_bg=chunk_bg, transfer_vel=transfer_vel,
                                ray_chunk=ray_chunk)
    return out.rgb_map, out.depth_map, out.acc_map, out.weights, out.mask_map
