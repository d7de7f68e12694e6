"""Build an oracle ``Scene`` from a config + a reference-keyed state dict.

TEST INFRASTRUCTURE (see nvfi_oracle.py header).  The construction-time decisions of the
reference that matter on the hot path are restated here:
  * velocity gate choice and bounds: models/tensorf_keyframe.py:96-107,
    models/velocity_field.py:38-44;
  * attribute plumbing of TensorBase.__init__: models/tensorf_base.py:134-183.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from .nvfi_oracle import Scene

VEL_KEYS = ["1", "3.0", "4.0", "5.0", "6.0", "7.0"]


def scene_from_state(cfg, grid_size: Sequence[int], num_keyframes: int,
                     sd: Dict[str, torch.Tensor], *, alpha_volume: Optional[torch.Tensor] = None,
                     mask_field=None, prefix: str = "", requires_grad: bool = False) -> Scene:
    nv = cfg.nvfi
    aabb = torch.stack([torch.tensor(nv[k], dtype=torch.float32)
                        for k in ("bbox_x", "bbox_y", "bbox_z")], dim=-1)

    def g(key):
        t = sd[prefix + key].detach().clone().float()
        if requires_grad:
            t.requires_grad_(True)
        return t

    sc = Scene(
        aabb=aabb, grid_size=list(grid_size), num_keyframes=num_keyframes, tmax=float(nv.tmax),
        near=float(cfg.dataset.near), far=float(cfg.dataset.far), step_ratio=float(nv.step_ratio),
        max_n_samples=int(nv.max_n_samples), density_shift=float(nv.density_shift),
        distance_scale=float(nv.distance_scale), alpha_mask_thres=float(nv.alphaMask_thres),
        ray_march_weight_thres=float(nv.rayMarch_weight_thres), fea2dense_act=nv.fea2denseAct,
        shading_mode=nv.shadingMode, pos_pe=int(nv.pos_pe), view_pe=int(nv.view_pe),
        use_vel=bool(nv.use_vel),
    )
    sc.density_plane_space = [g(f"density_plane_space.{k}") for k in range(3)]
    sc.density_plane_time = [g(f"density_plane_time.{k}") for k in range(3)]
    sc.app_plane_space = [g(f"app_plane_space.{k}") for k in range(3)]
    sc.app_plane_time = [g(f"app_plane_time.{k}") for k in range(3)]
    sc.basis_mat = g("basis_mat.weight")
    if nv.shadingMode == "MLP_PE":
        sc.render_mlp = [(g(f"renderModule.mlp.{i}.weight"), g(f"renderModule.mlp.{i}.bias"))
                         for i in (0, 2, 4)]
    sc.vel_net = [(g(f"vel_net.weight_net.{k}.weight"), g(f"vel_net.weight_net.{k}.bias"))
                  for k in VEL_KEYS]
    sc.acc_net = [(g(f"vel_net.a_weight_net.{k}.weight"), g(f"vel_net.a_weight_net.{k}.bias"))
                  for k in VEL_KEYS]
    if "sur_x" in nv:
        sur = torch.stack([torch.tensor(nv[k], dtype=torch.float32)
                           for k in ("sur_x", "sur_y", "sur_z")], dim=-1)
        sc.vel_gate = "sur"
        sc.vel_bounds = (sur - aabb[0]) * 2 / (aabb[1] - aabb[0]) - 1
    else:
        sc.vel_gate = "aabb"
        sc.vel_eps = float(nv.eps) if "eps" in nv else 0.03
    sc.alpha_volume = alpha_volume
    sc.mask_field = mask_field
    return sc
