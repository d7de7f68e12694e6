"""Diagnostic: gradients of the tensor-core and the SIMT paths, run to run and against each other."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import engine
from nvfi_b200.scenes import build_scene, frame_rays

cfg, nv, _ = build_scene("bat", grid=(199, 199, 199), step_ratio=1.79)
nv.requires_grad_(True)
f = nv.nvfi
f.train()
o, d = frame_rays(800, 800)
r0 = 398
oo, dd = o[r0 * 800:(r0 + 4) * 800].contiguous().cuda(), d[r0 * 800:(r0 + 4) * 800].contiguous().cuda()
n = oo.shape[0]
gen = torch.Generator().manual_seed(9)
jit = torch.rand(n, 1, generator=gen)
tgt = torch.rand(n, 3, generator=gen).cuda()


def nrel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run(t, mode):
    prev = engine.set_mlp_mode(mode)
    try:
        nv.zero_grad(set_to_none=True)
        out = engine.render_forward(f.binding, oo, dd, t, white_bg=True, training=True, jitter=jit.cuda(),
                                    ray_chunk=2048, want_stats=True)
        stats = out.stats.tolist()
        xa, w = out.x_adv.clone(), out.weights.clone()
        rgb, *_ = f.render_rays(t, oo, dd, white_bg=True, ray_chunk=2048, jitter=jit)
        torch.nn.functional.mse_loss(rgb, tgt).backward()
        g = {k: p.grad.detach().clone() for k, p in nv.named_parameters() if p.grad is not None}
    finally:
        engine.set_mlp_mode(prev)
    return g, xa, w, stats, rgb.detach().clone()


for t in (0.33, 1.0):
    a1, xa1, w1, s1, rgb1 = run(t, "tf32x3")
    a2, xa2, w2, s2, rgb2 = run(t, "tf32x3")
    b1, xb1, wb1, sb1, rgbb = run(t, "simt")
    b2, *_ = run(t, "simt")
    valid = w1 > 0
    print(f"t={t}: stats tc {s1} simt {sb1}; |x_adv tc-simt| max {float((xa1 - xb1)[wb1 > 1e-6].abs().max()):.2e}; "
          f"weights max diff {float((w1 - wb1).abs().max()):.2e}; rgb max diff {float((rgb1 - rgbb).abs().max()):.2e}")
    groups = {"dens": "density_plane", "app": "app_plane", "basis": "basis_mat", "mlp0": "mlp.0.w", "mlp2": "mlp.2.w",
              "mlp4": "mlp.4.w", "vel": "weight_net"}
    for gname, key in groups.items():
        ks = [k for k in a1 if key in k]
        print(f"   {gname:6s} tc/tc {max(nrel(a1[k], a2[k]) for k in ks):.1e}  simt/simt {max(nrel(b1[k], b2[k]) for k in ks):.1e}"
              f"  tc/simt {max(nrel(a1[k], b1[k]) for k in ks):.1e}")
