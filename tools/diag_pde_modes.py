"""get_vel_loss gradients of every arithmetic mode against the CPU oracle (functorch Jacobian + autograd) on a
2 048-point subsample of the fallingball scene: per-tensor relative-norm errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import pde, engine
from nvfi_b200.scenes import build_scene
from oracle import nvfi_oracle as O
from oracle.scene_io import scene_from_state
from tests.helpers import norm_rel_err, oracle_param_map

cfg, nv, sd = build_scene("fallingball", grid=(199, 199, 199))
f = nv.nvfi
P = 262144
gen = torch.Generator().manual_seed(5)
pts = (torch.rand(P, 3, generator=gen) * 2 - 1).cuda()
t = torch.rand(P, 1, generator=gen).cuda()
keep = pde.occupancy_filter(f, pts, t)
n_occ = int(keep.sum())
idx = torch.nonzero(keep).reshape(-1)[torch.randperm(n_occ, generator=gen)[:2048].cuda()]
xyzt = torch.cat([pts[idx], t[idx]], -1)
sc = scene_from_state(cfg, [199, 199, 199], int(cfg.nvfi.num_keyframes), sd, requires_grad=True)
ref = O.pde_loss_from_points(sc, xyzt.cpu())
ref.backward()
# float64 oracle
prev = torch.get_default_dtype()
torch.set_default_dtype(torch.float64)
sc64 = scene_from_state(cfg, [199, 199, 199], int(cfg.nvfi.num_keyframes), sd, requires_grad=True)
for name in ("vel_net", "acc_net"):
    v = getattr(sc64, name)
    setattr(sc64, name, [(w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)) for w, b in v])
ref64 = O.pde_loss_from_points(sc64, xyzt.cpu().double())
ref64.backward()
torch.set_default_dtype(prev)
g64 = {}
for net, key in (("vel_net", "vel_net.weight_net"), ("acc_net", "vel_net.a_weight_net")):
    names = ["1", "3.0", "4.0", "5.0", "6.0", "7.0"]
    for (w, b), nm in zip(getattr(sc64, net), names):
        g64[f"{key}.{nm}.weight"], g64[f"{key}.{nm}.bias"] = w.grad, b.grad
nv.requires_grad_(True)
pm = oracle_param_map(sc)
params = dict(f.named_parameters())
print(f"loss ref32 {ref.item():.8f} ref64 {ref64.item():.8f}")
for mode in ("simt", "f16x3"):
    engine.set_mlp_mode(mode)
    nv.zero_grad(set_to_none=True)
    loss = pde.pde_loss_from_points(f, xyzt)
    loss.backward()
    print(f"== {mode}: loss {loss.item():.8f}")
    for name, p in pm.items():
        if "vel_net" in name and p.grad is not None:
            g = params[name].grad.cpu()
            e32 = norm_rel_err(g, p.grad)
            e64 = norm_rel_err(g.double(), g64[name]) if name in g64 else float("nan")
            r64 = norm_rel_err(p.grad.double(), g64[name]) if name in g64 else float("nan")
            print(f"  {name:40s} vs oracle32 {e32:.2e}  vs oracle64 {e64:.2e}   (oracle32 vs 64 {r64:.2e})")
