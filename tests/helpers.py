"""Shared test helpers: golden-fixture loading and comparison utilities."""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
import torch

from nvfi_b200.configs import AttrDict
from oracle import nvfi_oracle as O
from oracle.scene_io import scene_from_state

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_SCENES = ("bat_small", "chess_small", "sh_small")


class Golden:
    def __init__(self, name: str):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))
        self.cfg = AttrDict(self.meta["cfg"])
        self.grid = self.meta["grid"]
        self.K = self.meta["K"]
        self.cfg.nvfi.num_keyframes = self.K
        self.ray_chunk = self.meta["ray_chunk"]
        self.sd: Dict[str, torch.Tensor] = {
            k[3:]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith("sd/")}

    def t(self, key) -> torch.Tensor:
        return torch.from_numpy(self.z[key])

    def has(self, key) -> bool:
        return key in self.z.files

    def keys(self, prefix):
        return [k for k in self.z.files if k.startswith(prefix)]

    def rays(self):
        return self.t("rays_o"), self.t("rays_d")

    def scene(self, requires_grad=False, alpha=False, mask_field=False) -> O.Scene:
        av = None
        if alpha:
            av = self.t("alpha/volume").float()
        mf = None
        if mask_field:
            mf = []
            i = 0
            while self.has(f"maskfield/{i}/weight"):
                mf.append((self.t(f"maskfield/{i}/weight"), self.t(f"maskfield/{i}/bias")))
                i += 1
        return scene_from_state(self.cfg, self.grid, self.K, self.sd, alpha_volume=av,
                                mask_field=mf, requires_grad=requires_grad)

    def case(self, name) -> Dict[str, torch.Tensor]:
        pre = f"case/{name}/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def loss_weights(self):
        return {k: self.t(f"lossw/{k}") for k in ("wr", "wd", "wa", "ww")}


def scalar_loss(out, lw):
    rgb, depth, acc, w = out[:4]
    return (rgb * lw["wr"]).sum() + (depth * lw["wd"]).sum() + (acc * lw["wa"]).sum() + (w * lw["ww"]).sum()


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor)  — element-wise 'relative FP32' with an absolute floor for values near 0
    (rgb / acc / weights live in [0, 1]: for them this is the largest ABSOLUTE error).  Used together with
    ``norm_rel_err`` (per-tensor relative L2 norm, no floor) wherever the 1e-4 gate is applied: see
    ``assert_close``."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    if a.numel() == 0:
        return 0.0
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def assert_close(a, b, tol=1e-4, what=""):
    """The parity gate for value tensors: BOTH the floored element-wise error and the per-tensor relative
    L2-norm error must be below ``tol``.  Returns (max_abs_floored, rel_norm) for reporting."""
    e_max, e_norm = rel_err(a, b), norm_rel_err(a, b)
    assert e_max < tol and e_norm < tol, f"{what}: max |d|/max(|ref|,1) = {e_max:.3e}, ||d||/||ref|| = {e_norm:.3e}"
    return e_max, e_norm


def norm_rel_err(a, b):
    """||a-b|| / max(||b||, tiny): for gradient tensors."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# reference parameter name (relative to the field) -> accessor on an oracle Scene
def oracle_param_map(sc: O.Scene) -> Dict[str, torch.Tensor]:
    m = {}
    for k in range(3):
        m[f"density_plane_space.{k}"] = sc.density_plane_space[k]
        m[f"density_plane_time.{k}"] = sc.density_plane_time[k]
        m[f"app_plane_space.{k}"] = sc.app_plane_space[k]
        m[f"app_plane_time.{k}"] = sc.app_plane_time[k]
    m["basis_mat.weight"] = sc.basis_mat
    if sc.render_mlp is not None:
        for i, (w, b) in zip((0, 2, 4), sc.render_mlp):
            m[f"renderModule.mlp.{i}.weight"] = w
            m[f"renderModule.mlp.{i}.bias"] = b
    keys = ["1", "3.0", "4.0", "5.0", "6.0", "7.0"]
    for net, layers in (("weight_net", sc.vel_net), ("a_weight_net", sc.acc_net)):
        for k, (w, b) in zip(keys, layers):
            m[f"vel_net.{net}.{k}.weight"] = w
            m[f"vel_net.{net}.{k}.bias"] = b
    return m


# ------------------------------------------------------------------------------------------
# CUDA model construction from a golden fixture (used by the -m gpu tests)
# ------------------------------------------------------------------------------------------
def build_model(g: "Golden", device="cuda", alpha=False, mask_field=False, requires_grad=False):
    """nvfi_b200.models.NVFi carrying the fixture's parameters."""
    from nvfi_b200 import models as M
    from nvfi_b200.synth import aabb_from_cfg

    nv = M.NVFi(g.cfg, device, aabb_from_cfg(g.cfg), list(g.grid), [g.cfg.dataset.near, g.cfg.dataset.far])
    nv = nv.to(device)
    missing, unexpected = nv.load_state_dict({"nvfi." + k: v for k, v in g.sd.items()}, strict=False)
    assert not unexpected, unexpected
    bad = [m for m in missing if "frequency_bands" not in m and ".vel.vel_net." not in m and m != "nvfi.aabb"]
    assert not bad, bad
    f = nv.nvfi
    if alpha:
        vol = g.t("alpha/volume").float().to(device)
        f.alphaMask = M.AlphaGridMask(device, f.aabb, vol)
    if mask_field:
        mf = M.MaskField(n_layer=4, n_dim=128, input_dim=3, skips=[], mask_dim=3, mask_act="softmax")
        lins = list(mf.point_fc) + [mf.mask_fc]
        for i, lin in enumerate(lins):
            lin.weight.data.copy_(g.t(f"maskfield/{i}/weight"))
            lin.bias.data.copy_(g.t(f"maskfield/{i}/bias"))
        f.mask_field = mf.to(device)
    nv.requires_grad_(requires_grad)
    return nv
