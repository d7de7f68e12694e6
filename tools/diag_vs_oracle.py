"""Gradients of a small train step at an extrapolated time (10 RK2 steps) against the oracle's
autograd, in both arithmetic modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import engine
from nvfi_b200.scenes import build_scene, frame_rays
from oracle import nvfi_oracle as O
from oracle.scene_io import scene_from_state

grid = (48, 48, 48)
cfg, nv, sd = build_scene("bat", grid=grid, device="cuda:0", max_n_samples=64)
f = nv.nvfi
nv.requires_grad_(True)
f.train()
o, d = frame_rays(800, 800, crop=(368, 368, 64, 64))
gen = torch.Generator().manual_seed(0)
n = 1024
sel = torch.randperm(o.shape[0], generator=gen)[:n]
oo, dd = o[sel].contiguous(), d[sel].contiguous()
jit = torch.rand(n, 1, generator=gen)
tgt = torch.rand(n, 3, generator=gen)
for t in (0.33, 1.0):
    scg = scene_from_state(cfg, list(grid), int(cfg.nvfi.num_keyframes), sd, requires_grad=True)
    ref = O.render_chunk(scg, t, oo, dd, white_bg=True, training=True, jitter=jit)
    torch.nn.functional.mse_loss(ref[0], tgt).backward()
    refs = {"density_plane_space.0": scg.density_plane_space[0], "app_plane_space.0": scg.app_plane_space[0],
            "app_plane_time.1": scg.app_plane_time[1], "basis_mat.weight": scg.basis_mat,
            "renderModule.mlp.0.weight": scg.render_mlp[0][0], "renderModule.mlp.2.weight": scg.render_mlp[1][0],
            "renderModule.mlp.4.weight": scg.render_mlp[2][0], "vel_net.weight_net.4.0.weight": scg.vel_net[2][0]}
    for mode in ("tf32x3", "simt"):
        prev = engine.set_mlp_mode(mode)
        nv.zero_grad(set_to_none=True)
        rgb, *_ = f.render_rays(t, oo.cuda(), dd.cuda(), white_bg=True, ray_chunk=2048, jitter=jit)
        torch.nn.functional.mse_loss(rgb, tgt.cuda()).backward()
        engine.set_mlp_mode(prev)
        got = dict(nv.named_parameters())
        errs = {k: float((got["nvfi." + k].grad.cpu() - r.grad).norm() / r.grad.norm().clamp_min(1e-30))
                for k, r in refs.items()}
        print(f"t={t} {mode:7s}", {k.split('.')[0] + k[-9:]: f"{v:.1e}" for k, v in errs.items()})
