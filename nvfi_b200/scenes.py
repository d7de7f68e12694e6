"""Build ready-to-render models for the BASELINE.json workloads from synthetic parameters
(SURVEY.md section 8d).  Shared by bench.py, __graft_entry__.smoke() and the tests."""
from __future__ import annotations

import torch

from . import configs, synth


def build_scene(name: str = "bat", grid=(199, 199, 199), device="cuda", seed: int = 233, **overrides):
    """Returns (cfg, nvfi_b200.models.NVFi on `device`, state dict (CPU tensors))."""
    from . import models as M

    cfg = configs.get_config(name, **overrides)
    aabb = synth.aabb_from_cfg(cfg)
    K = int(cfg.nvfi.num_keyframes)
    sd = synth.synth_state(cfg, list(grid), K, seed=seed)
    gen_state = torch.get_rng_state()
    nv = M.NVFi(cfg, device, aabb, list(grid), [cfg.dataset.near, cfg.dataset.far]).to(device)
    torch.set_rng_state(gen_state)
    missing, unexpected = nv.load_state_dict({"nvfi." + k: v for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    return cfg, nv, sd


def frame_rays(H: int = 800, W: int = 800, theta: float = 30.0, phi: float = -30.0, radius: float = 4.0,
               crop=None, z_shift: float = 0.0):
    """Pinhole rays of the synthetic camera (host tensors, (H*W,3) each)."""
    pose = synth.pose_spherical(theta, phi, radius)
    pose[2, 3] += z_shift
    focal = synth.blender_focal(W)
    if crop is None:
        o, d = synth.pinhole_rays(pose, H, W, focal)
    else:
        x0, y0, w, h = crop
        o, d = synth.pinhole_rays(pose, H, W, focal, x0=x0, y0=y0, h=h, w=w)
    return o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
