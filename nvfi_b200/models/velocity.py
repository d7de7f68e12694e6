"""Velocity-field parameter containers (reference: models/velocity_field.py:14-98,
models/base_network.py:20-54).

These modules own the parameters under the reference's ``state_dict`` names
(``weight_net.1.weight``, ``weight_net.3.0.weight`` ... ``a_weight_net.7.0.bias``,
``weight_net.0.frequency_bands``).  Evaluation runs in CUDA (``nvfi_velocity``); the
differentiable uses of the velocity field are the fused render backward and the PDE
loss, so calling these modules with autograd enabled on inputs is not supported.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine


def N_to_reso(n_voxels, bbox):
    """Grid resolution for a voxel budget (models/velocity_field.py:14-18)."""
    lo, hi = bbox
    extent = hi - lo
    cell = (extent.prod() / n_voxels).pow(1 / len(lo))
    return (extent / cell).long().tolist()


class PositionEncoder(nn.Module):
    """[x, sin(f x), cos(f x) for f in 2^0..2^(L-1)] (models/base_network.py:20-54).
    Only a buffer holder here: the encoding is computed inside the velocity kernels."""

    def __init__(self, encode_dim, log_sampling=True):
        super().__init__()
        self.encode_dim = encode_dim
        if log_sampling:
            bands = 2.0 ** torch.linspace(0.0, encode_dim - 1, encode_dim, dtype=torch.float32)
        else:
            bands = torch.linspace(2.0 ** 0.0, 2.0 ** (encode_dim - 1), encode_dim, dtype=torch.float32)
        self.register_buffer("frequency_bands", bands)

    def forward(self, x):
        parts = [x]
        for f in self.frequency_bands:
            parts += [torch.sin(x * f), torch.cos(x * f)]
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


def _weight_net(act):
    net = nn.Sequential(PositionEncoder(3), nn.Linear(28, 128), act())
    for _ in range(4):
        net.append(nn.Sequential(nn.Linear(128, 128), act()))
    net.append(nn.Sequential(nn.Linear(128, 6)))
    return net


class VelBasis(nn.Module):
    """6 rigid-motion basis weights from a SiLU MLP, plus the ReLU twin for the
    acceleration (models/velocity_field.py:54-98)."""

    def __init__(self):
        super().__init__()
        self.weight_net = _weight_net(nn.SiLU)
        self.a_weight_net = _weight_net(nn.ReLU)
        self._binding = None   # set by the owning field

    def _eval(self, xt, full):
        if self._binding is None:
            raise RuntimeError("VelBasis is evaluated through its owning field (no binding set)")
        if torch.is_grad_enabled() and xt.requires_grad:
            raise RuntimeError("nvfi_b200: autograd through a bare VelBasis call is not supported; "
                               "use the fused render / get_vel_loss paths")
        shape = xt.shape[:-1]
        return engine.velocity(self._binding, xt, full).reshape(*shape, 6 if full else 3)

    def forward(self, xt):
        return self._eval(xt, True)

    def get_vel(self, xt):
        return self._eval(xt, True)[..., :3]


class _GatedVelocity(nn.Module):
    def __init__(self, vel_net):
        super().__init__()
        self.vel_net = vel_net

    def forward(self, xt):
        b = self.vel_net._binding
        if b is None:
            raise RuntimeError("velocity gate is evaluated through its owning field")
        return engine.velocity(b, xt, False).reshape(*xt.shape[:-1], 3)


class VelocityAABB(_GatedVelocity):
    """Zero velocity outside |x| <= 1 - eps (models/velocity_field.py:21-33)."""

    def __init__(self, vel_net, eps=-0.03):
        super().__init__(vel_net)
        self.eps = eps

    def gate(self):
        return "aabb", [engine._f32(-1 + self.eps)] * 3, [engine._f32(1 - self.eps)] * 3


class VelocityAABBSur(_GatedVelocity):
    """Zero velocity outside the normalised 'surround' box (models/velocity_field.py:36-51)."""

    def __init__(self, vel_net, aabb, surround):
        super().__init__(vel_net)
        self.aabb = aabb
        self.surround = surround
        self.bounds = (surround - aabb[0]) * 2 / (aabb[1] - aabb[0]) - 1

    def gate(self):
        b = self.bounds.detach().float().cpu()
        return "sur", [float(v) for v in b[0]], [float(v) for v in b[1]]
