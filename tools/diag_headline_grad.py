"""Diagnostic: gradient parity of one headline-size chunk against the FP32 and the FLOAT64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200.scenes import build_scene, frame_rays
from oracle import nvfi_oracle as O
from oracle.scene_io import scene_from_state
from tests.helpers import norm_rel_err, oracle_param_map

GRID = (199, 199, 199)
where = sys.argv[1] if len(sys.argv) > 1 else "silhouette"
s0 = {"top": 0, "silhouette": 230 * 800, "centre": 400 * 800 + 296}[where]
cfg, nv, sd = build_scene("bat", grid=GRID, step_ratio=1.79)
o, d = frame_rays(800, 800, theta=30.0)
oo, dd = o[s0:s0 + 2048].contiguous(), d[s0:s0 + 2048].contiguous()
gen = torch.Generator().manual_seed(11)
jit = torch.rand(2048, 1, generator=gen)
target = torch.rand(2048, 3, generator=gen)
nv.requires_grad_(True)
f = nv.nvfi
f.train()
rgb, depth, acc, w, _ = f.render_rays(0.33, oo.cuda(), dd.cuda(), white_bg=True, ray_chunk=2048, jitter=jit)
torch.nn.functional.mse_loss(rgb, target.cuda()).backward()
K = int(cfg.nvfi.num_keyframes)
mine = (w.detach().cpu() > 1e-4)


def run(dtype):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        sc = scene_from_state(cfg, list(GRID), K, sd, requires_grad=True)
        if dtype == torch.float64:
            def cast(x):
                if isinstance(x, torch.Tensor):
                    return x.detach().double().requires_grad_(True)
                if isinstance(x, (list, tuple)):
                    return type(x)(cast(y) for y in x)
                return x
            for name in ("density_plane_space", "density_plane_time", "app_plane_space", "app_plane_time",
                         "basis_mat", "render_mlp", "vel_net", "acc_net"):
                setattr(sc, name, cast(getattr(sc, name)))
            sc.aabb = sc.aabb.double()
        r = O.render_chunk(sc, 0.33, oo.to(dtype), dd.to(dtype), white_bg=True, training=True, jitter=jit.to(dtype),
                           app_mask_override=mine)
        torch.nn.functional.mse_loss(r[0], target.to(dtype)).backward()
        return {k: p.grad.detach().double() for k, p in oracle_param_map(sc).items() if p.grad is not None}, r
    finally:
        torch.set_default_dtype(prev)


g32, r32 = run(torch.float32)
g64, r64 = run(torch.float64)
params = dict(f.named_parameters())
print(f"{where}: app samples {int(mine.sum())}, rgb err vs f32 {float((rgb.detach().cpu() - r32[0].detach()).abs().max()):.2e}")
print(f"{'tensor':42s} {'|g|':>10s} {'cuda-f32':>10s} {'cuda-f64':>10s} {'f32-f64':>10s}")
for k in g32:
    if "a_weight_net" in k or params[k].grad is None or float(g32[k].abs().max()) == 0:
        continue
    c = params[k].grad.detach().double().cpu()
    print(f"{k:42s} {float(g64[k].norm()):10.3e} {norm_rel_err(c, g32[k]):10.2e} {norm_rel_err(c, g64[k]):10.2e} "
          f"{norm_rel_err(g32[k], g64[k]):10.2e}")
