"""Diagnostic: velocity-net gradients of the tensor-core backward vs the FP32 SIMT backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import Golden, build_model, scalar_loss
from nvfi_b200 import engine

g = Golden(sys.argv[1] if len(sys.argv) > 1 else "bat_small")
case = g.case("train0")
o, d = g.rays()
lw = {k: v.cuda() for k, v in g.loss_weights().items()}
res = {}
for mode in ("simt", "tf32x3"):
    engine.set_mlp_mode(mode)
    model = build_model(g, requires_grad=True)
    f = model.nvfi; f.train()
    bg = torch.from_numpy(case["random_bg"].astype(np.uint8)) if len(case["random_bg"]) else None
    out = f.render_rays(float(case["t"]), o.cuda(), d.cuda(), white_bg=bool(g.cfg.dataset.white_background),
                        ray_chunk=g.ray_chunk, jitter=torch.from_numpy(case["jitter"]), chunk_bg=bg)
    scalar_loss(out, lw).backward()
    torch.cuda.synchronize()
    res[mode] = {k: v.grad.detach().double().cpu() for k, v in f.named_parameters() if v.grad is not None and "vel_net.weight_net" in k}
for k in res["simt"]:
    a, b = res["simt"][k], res["tf32x3"][k]
    cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-300))
    print(f"{k:34s} |simt| {float(a.norm()):.4e} |tc| {float(b.norm()):.4e} cos {cos:+.6f} relerr {float((a-b).norm()/a.norm()):.3e}")
k = "vel_net.weight_net.4.0.weight"
a, b = res["simt"][k], res["tf32x3"][k]
print("simt[0,:6]", a[0, :6].tolist()); print("tc  [0,:6]", b[0, :6].tolist())
print("simt[:6,0]", a[:6, 0].tolist()); print("tc  [:6,0]", b[:6, 0].tolist())
# is tc a permutation / transpose of simt?
print("relerr vs transpose", float((a.t() - b).norm() / a.norm()))
