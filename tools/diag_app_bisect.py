"""Diagnostic: find the rays of the headline 'silhouette' chunk whose appearance gradients differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200.scenes import build_scene, frame_rays
from nvfi_b200 import engine
from oracle import nvfi_oracle as O
from oracle.scene_io import scene_from_state
from tests.helpers import norm_rel_err

GRID = (199, 199, 199)
s0 = 230 * 800
cfg, nv, sd = build_scene("bat", grid=GRID, step_ratio=1.79)
o, d = frame_rays(800, 800, theta=30.0)
K = int(cfg.nvfi.num_keyframes)
gen = torch.Generator().manual_seed(11)
jit_all = torch.rand(2048, 1, generator=gen)
wr_all = torch.randn(2048, 3, generator=gen)
nv.requires_grad_(True)
f = nv.nvfi
f.train()
KEY = "renderModule.mlp.0.weight"


def err(sel):
    oo, dd = o[s0:s0 + 2048][sel].contiguous(), d[s0:s0 + 2048][sel].contiguous()
    jit, wr = jit_all[sel], wr_all[sel]
    nv.zero_grad(set_to_none=True)
    rgb, depth, acc, w, _ = f.render_rays(0.33, oo.cuda(), dd.cuda(), white_bg=True, ray_chunk=2048, jitter=jit)
    (rgb * wr.cuda()).sum().backward()
    mine = (w.detach().cpu() > 1e-4)
    sc = scene_from_state(cfg, list(GRID), K, sd, requires_grad=True)
    r = O.render_chunk(sc, 0.33, oo, dd, white_bg=True, training=True, jitter=jit, app_mask_override=mine, return_aux=True)
    (r[0] * wr).sum().backward()
    g = dict(f.named_parameters())[KEY].grad.cpu()
    gr = sc.render_mlp[0][0].grad
    return norm_rel_err(g, gr), float((g - gr).norm()), float(gr.norm()), (rgb, w, r, mine)


idx = torch.arange(2048)
e, dn, rn, _ = err(idx)
print(f"all: rel {e:.2e} |d| {dn:.3e} |ref| {rn:.3e}")
lo, hi = 0, 2048
while hi - lo > 1:
    mid = (lo + hi) // 2
    ea = err(idx[lo:mid])
    eb = err(idx[mid:hi])
    print(f"[{lo},{mid}) |d| {ea[1]:.3e} rel {ea[0]:.2e}   [{mid},{hi}) |d| {eb[1]:.3e} rel {eb[0]:.2e}")
    if ea[1] >= eb[1]:
        hi = mid
    else:
        lo = mid
ray = lo
e, dn, rn, (rgb, w, r, mine) = err(idx[ray:ray + 1])
print(f"worst ray {ray}: rel {e:.2e}; rgb cuda {rgb.detach().cpu().tolist()} ref {r[0].detach().tolist()}")
aux = r[5]
ws = w.detach().cpu()[0]
sel = torch.nonzero(mine[0]).reshape(-1)
print("app samples", sel.tolist())
print("w cuda", ws[sel].tolist())
print("w ref ", r[3].detach()[0][sel].tolist())
print("x_adv ref", aux["xyz_adv"][0][sel].tolist())
out = engine.render_forward(f.binding, o[s0 + ray:s0 + ray + 1].cuda(), d[s0 + ray:s0 + ray + 1].cuda(), 0.33, white_bg=True,
                            training=True, jitter=jit_all[ray:ray + 1], ray_chunk=2048)
print("x_adv cuda", out.x_adv[0][sel.cuda()].cpu().tolist())
print("rgb_s cuda", out.rgb[0][sel.cuda()].cpu().tolist())
print("rgb_s ref ", aux["rgb"][0][sel].tolist())
