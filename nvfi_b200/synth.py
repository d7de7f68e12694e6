"""Synthetic scenes for parity tests and the bench (no dataset ships with the task).

Implements the recipe of SURVEY.md section 8(d): an occupied cube in the middle of the
box (so valid / app masks are non-trivial), time planes around 1, default-initialised
MLPs with an O(1) velocity bias, and cameras on the Blender ``pose_spherical`` circle
(datasets/load_blender.py:62-67 in the reference).  Parameter tensors are returned under
the reference's ``state_dict`` key names (relative to the field module, i.e. without
the leading ``nvfi.``), so the same dict loads into the reference model, into
``nvfi_b200.models`` and into the test oracle.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch

MAT_MODE_SPACE = ((0, 1), (0, 2), (1, 2))
MAT_MODE_TIME = ((2, 3), (1, 3), (0, 3))


def n_to_reso(n_voxels: float, aabb: torch.Tensor) -> List[int]:
    """Grid resolution for a voxel budget (utils/tensorf_utils.py:53-57)."""
    size = aabb[1] - aabb[0]
    voxel = (size.prod() / n_voxels).pow(1 / 3)
    return (size / voxel).long().tolist()


def aabb_from_cfg(cfg) -> torch.Tensor:
    """train_nvfi.py:63-64."""
    bx, by, bz = [torch.tensor(cfg.nvfi[k], dtype=torch.float32) for k in ("bbox_x", "bbox_y", "bbox_z")]
    return torch.stack([bx, by, bz], dim=-1)


def _linear_init(gen: torch.Generator, out_f: int, in_f: int):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)))."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def synth_state(cfg, grid_size: Sequence[int], num_keyframes: int, seed: int = 233,
                app_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters, keyed like the reference field's state_dict."""
    gen = torch.Generator().manual_seed(seed)
    nv = cfg.nvfi
    G = list(grid_size)
    K = num_keyframes
    sd: Dict[str, torch.Tensor] = {}

    def cube_plane(R, H, W):
        u = torch.linspace(-1, 1, W)[None, :].expand(H, W)
        v = torch.linspace(-1, 1, H)[:, None].expand(H, W)
        inside = ((u.abs() < 0.5) & (v.abs() < 0.5)).float()
        return (0.05 * torch.rand(1, R, H, W, generator=gen) + 0.1 + 0.9 * inside[None, None])

    for k in range(3):
        m0, m1 = MAT_MODE_SPACE[k]
        n0, _ = MAT_MODE_TIME[k]
        Rd, Ra = nv.density_n_comp[k], nv.appearance_n_comp[k]
        sd[f"density_plane_space.{k}"] = cube_plane(Rd, G[m1], G[m0])
        sd[f"density_plane_time.{k}"] = 1 + 0.1 * torch.randn(1, Rd, K, G[n0], generator=gen)
        sd[f"app_plane_space.{k}"] = app_scale * (0.2 + 0.8 * torch.rand(1, Ra, G[m1], G[m0], generator=gen))
        sd[f"app_plane_time.{k}"] = 1 + 0.1 * torch.randn(1, Ra, K, G[n0], generator=gen)

    sd["basis_mat.weight"] = _linear_init(gen, nv.app_dim, nv.appearance_n_comp[0])[0]
    sd["basis_mat_density.weight"] = _linear_init(gen, 1, nv.density_n_comp[0])[0]

    if nv.shadingMode == "MLP_PE":
        in_c = (3 + 2 * nv.view_pe * 3) + (3 + 2 * nv.pos_pe * 3) + nv.app_dim
        dims = [(nv.featureC, in_c), (nv.featureC, nv.featureC), (3, nv.featureC)]
        for idx, (o, i) in zip((0, 2, 4), dims):
            w, b = _linear_init(gen, o, i)
            if idx == 4:
                b = torch.zeros_like(b)       # models/tensorf_base.py:86
            sd[f"renderModule.mlp.{idx}.weight"] = w
            sd[f"renderModule.mlp.{idx}.bias"] = b

    vel_dims = [(128, 28)] + [(128, 128)] * 4 + [(6, 128)]
    vel_keys = ["1", "3.0", "4.0", "5.0", "6.0", "7.0"]   # models/velocity_field.py:60-67
    for net in ("weight_net", "a_weight_net"):
        for key, (o, i) in zip(vel_keys, vel_dims):
            w, b = _linear_init(gen, o, i)
            if net == "weight_net" and key == "7.0":
                b = torch.tensor([0.5, -0.3, 0.2, 1.0, -0.8, 0.6])
            sd[f"vel_net.{net}.{key}.weight"] = w
            sd[f"vel_net.{net}.{key}.bias"] = b
    return sd


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> torch.Tensor:
    """Camera-to-world pose on a sphere around the origin, Blender convention
    (behaviour of datasets/load_blender.py:62-67), written in closed form."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    trans = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=np.float64)
    rphi = np.array([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0],
                     [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1]], dtype=np.float64)
    rth = np.array([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0],
                    [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]], dtype=np.float64)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    return torch.tensor(flip @ rth @ rphi @ trans, dtype=torch.float32)


BLENDER_CAMERA_ANGLE_X = 0.6911112070083618


def blender_focal(width: int, camera_angle_x: float = BLENDER_CAMERA_ANGLE_X) -> float:
    """datasets/load_blender.py: focal = .5 W / tan(.5 camera_angle_x)."""
    return 0.5 * width / math.tan(0.5 * camera_angle_x)


def pinhole_rays(pose: torch.Tensor, H: int, W: int, focal: float, x0: int = 0, y0: int = 0,
                 h: int | None = None, w: int | None = None):
    """Host-side pinhole rays for a crop [y0:y0+h, x0:x0+w] of an HxW camera
    (same convention as models/camera.py:112-133).  Used to build synthetic inputs."""
    h = H if h is None else h
    w = W if w is None else w
    X, Y = torch.meshgrid(torch.arange(x0, x0 + w, dtype=torch.float32),
                          torch.arange(y0, y0 + h, dtype=torch.float32), indexing="xy")
    dirs = torch.stack([(X - W * 0.5) / focal, -(Y - H * 0.5) / focal, -torch.ones_like(X)], -1)
    d = torch.sum(dirs[..., None, :] * pose[:3, :3], dim=-1)
    o = pose[:3, -1].expand(d.shape)
    return o.contiguous(), d.contiguous()
