"""Diagnostic: compare backward intermediates (dL/dsigma, dL/dx_adv) and plane gradients of the
CUDA path with the oracle in float32 and float64 on a golden scene."""
import sys, os, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import Golden, build_model, oracle_param_map, norm_rel_err, scalar_loss
from oracle import nvfi_oracle as O
from nvfi_b200 import engine

name = sys.argv[1] if len(sys.argv) > 1 else "chess_small"
keys = tuple(sys.argv[2].split(",")) if len(sys.argv) > 2 else ("wr",)
g = Golden(name)
case = g.case(sys.argv[3] if len(sys.argv) > 3 else "train0")
o, d = g.rays()
lw = g.loss_weights()
white = bool(g.cfg.dataset.white_background)

FWD = {}


def run_oracle(dtype):
    torch.set_default_dtype(dtype)
    sc = g.scene()
    for f in dataclasses.fields(sc):
        v = getattr(sc, f.name)
        if torch.is_tensor(v): setattr(sc, f.name, v.to(dtype))
        elif isinstance(v, list) and v and torch.is_tensor(v[0]): setattr(sc, f.name, [x.to(dtype).requires_grad_(True) for x in v])
        elif isinstance(v, list) and v and isinstance(v[0], tuple): setattr(sc, f.name, [(w.to(dtype).requires_grad_(True), b.to(dtype).requires_grad_(True)) for w, b in v])
    sc.basis_mat = sc.basis_mat.requires_grad_(True)
    assert o.shape[0] <= g.ray_chunk * 100
    outs, auxs = [], []
    n = o.shape[0]; ch = g.ray_chunk
    loss = 0
    rb = list(case["random_bg"]) if len(case["random_bg"]) else None
    for c in range((n + ch - 1) // ch):
        sl = slice(c * ch, (c + 1) * ch)
        r = O.render_chunk(sc, float(case["t"]), o[sl].to(dtype), d[sl].to(dtype), white_bg=white, training=True,
                           jitter=torch.from_numpy(case["jitter"])[sl].to(dtype), random_bg=bool(rb[c]) if rb else False,
                           return_aux=True)
        aux = r[5]; aux["sigma"].retain_grad(); aux["rgb"].retain_grad() if aux["rgb"].requires_grad else None; aux["xyz_adv"].retain_grad() if (aux["xyz_adv"].requires_grad and not aux["xyz_adv"].is_leaf) else None
        auxs.append(aux)
        l = 0
        for key, idx in (("wr", 0), ("wd", 1), ("wa", 2), ("ww", 3)):
            if key in keys:
                l = l + (r[idx] * lw[key][sl].to(dtype)).sum()
        loss = loss + l
    loss.backward()
    pm = oracle_param_map(sc)
    gs = torch.cat([a["sigma"].grad for a in auxs], 0)
    gx = torch.cat([a["xyz_adv"].grad if (a["xyz_adv"].requires_grad and a["xyz_adv"].grad is not None) else torch.zeros_like(a["xyz_adv"]) for a in auxs], 0)
    valid = torch.cat([a["valid"] for a in auxs], 0)
    torch.set_default_dtype(torch.float32)
    FWD[str(dtype)] = dict(rgb=torch.cat([a["rgb"].detach() for a in auxs], 0).double(), x=torch.cat([a["xyz_adv"].detach() for a in auxs], 0).double(),
                           sigma=torch.cat([a["sigma"].detach() for a in auxs], 0).double())
    return {k: v.grad.double() for k, v in pm.items() if v.grad is not None}, gs.double(), gx.double(), valid

g32, gs32, gx32, valid = run_oracle(torch.float32)
g64, gs64, gx64, _ = run_oracle(torch.float64)

engine.DEBUG_KEEP = {}
model = build_model(g, requires_grad=True)
f = model.nvfi; f.train()
bg = torch.from_numpy(case["random_bg"].astype(np.uint8)) if len(case["random_bg"]) else None
out = f.render_rays(float(case["t"]), o.cuda(), d.cuda(), white_bg=white, ray_chunk=g.ray_chunk,
                    jitter=torch.from_numpy(case["jitter"]), chunk_bg=bg)
lwc = {k: (v.cuda() if k in keys else torch.zeros_like(v).cuda()) for k, v in lw.items()}
scalar_loss(out, lwc).backward()
run1 = {k: v.grad.detach().clone() for k, v in f.named_parameters() if v.grad is not None}
# run-to-run noise of the atomics-based accumulation: same inputs, second backward
model.zero_grad(set_to_none=True)
out2 = f.render_rays(float(case["t"]), o.cuda(), d.cuda(), white_bg=white, ray_chunk=g.ray_chunk,
                     jitter=torch.from_numpy(case["jitter"]), chunk_bg=bg)
scalar_loss(out2, lwc).backward()
for k, v in f.named_parameters():
    if v.grad is not None and k in run1 and ("density_plane" in k or "weight_net.1.weight" in k):
        print(f"run-to-run {k:32s} {norm_rel_err(v.grad.cpu(), run1[k].cpu()):.2e}")
dbg = engine.DEBUG_KEEP
fo = engine.DEBUG_KEEP["fwd"]
x64, x32 = FWD["torch.float64"]["x"], FWD["torch.float32"]["x"]
s64, s32 = FWD["torch.float64"]["sigma"], FWD["torch.float32"]["sigma"]
xm, sm_ = fo.x_adv.cpu().double(), fo.sigma.cpu().double()
print("x_adv max abs err  mine-f64 %.3e  o32-f64 %.3e" % (float((xm[valid] - x64[valid]).abs().max()), float((x32[valid] - x64[valid]).abs().max())))
print("sigma max abs err  mine-f64 %.3e  o32-f64 %.3e ; norm-rel mine %.2e o32 %.2e" % (
    float((sm_[valid] - s64[valid]).abs().max()), float((s32[valid] - s64[valid]).abs().max()),
    norm_rel_err(sm_[valid], s64[valid]), norm_rel_err(s32[valid], s64[valid])))
gs = dbg["g_sigma"].cpu().double(); gx = dbg["g_x_adv"].cpu().double()
v = valid
print("dL/dsigma  mine vs f64: %.2e   oracle32 vs f64: %.2e" % (norm_rel_err(gs[v], gs64[v]), norm_rel_err(gs32[v], gs64[v])))
print("dL/dx_adv  mine vs f64: %.2e   oracle32 vs f64: %.2e" % (norm_rel_err(gx[v], gx64[v]), norm_rel_err(gx32[v], gx64[v])))
params = dict(f.named_parameters())
for k in g64:
    if g64[k].abs().max() == 0 or "a_weight" in k: continue
    if params[k].grad is None: continue
    mine = params[k].grad.cpu().double()
    print(f"{k:36s} mine-vs-f64 {norm_rel_err(mine, g64[k]):.2e}  o32-vs-f64 {norm_rel_err(g32[k], g64[k]):.2e}  |g| {float(g64[k].norm()):.3e}")
# structure of the worst plane
k = "density_plane_time.1"
e = (params[k].grad.cpu().double() - g64[k]).abs()
idx = torch.topk(e.reshape(-1), 5).indices
print("worst entries of", k, [(int(i), float(e.reshape(-1)[i]), float(g64[k].reshape(-1)[i])) for i in idx])

# ---- does the density-plane error come from the scatter or from dL/dsigma?  Push MY dL/dsigma and
# the oracle32's through an exact (float64) density backward and compare with the f64 gradient.
torch.set_default_dtype(torch.float64)
sc64 = g.scene()
for fld in dataclasses.fields(sc64):
    v_ = getattr(sc64, fld.name)
    if torch.is_tensor(v_): setattr(sc64, fld.name, v_.double())
    elif isinstance(v_, list) and v_ and torch.is_tensor(v_[0]): setattr(sc64, fld.name, [x.double().requires_grad_(True) for x in v_])
    elif isinstance(v_, list) and v_ and isinstance(v_[0], tuple): setattr(sc64, fld.name, [(w.double(), b.double()) for w, b in v_])
tt = torch.full((1, 1), float(case["t"]))
base = O.keyframe_snap(sc64, tt)
tnb = float(O.normalize_time_coord(sc64, base))
vmask = valid
xv = x64[vmask]
xyzt = torch.cat([xv, torch.full((xv.shape[0], 1), tnb)], -1)
sig = O.feature2density(sc64, O.density_feature(sc64, xyzt))
planes = list(sc64.density_plane_space) + list(sc64.density_plane_time)
names = [f"density_plane_space.{k}" for k in range(3)] + [f"density_plane_time.{k}" for k in range(3)]
for label, gsv in (("mine", gs), ("oracle32", gs32), ("f64", gs64)):
    grads = torch.autograd.grad(sig, planes, grad_outputs=gsv[vmask].double(), retain_graph=True)
    print("exact density backward of dL/dsigma from", label, {n[14:]: f"{norm_rel_err(gr, g64[n]):.2e}" for n, gr in zip(names, grads)})
    if label == "mine":
        print("   vs MY plane grads      ", {n[14:]: f"{norm_rel_err(params[n].grad.cpu().double(), gr):.2e}" for n, gr in zip(names, grads)})
torch.set_default_dtype(torch.float32)

# ---- structure of the dL/dsigma error: by sample class and per-ray correlation
dm, do = (gs - gs64), (gs32 - gs64)
N, S = gs64.shape
first_app = torch.full((N,), S, dtype=torch.long)
thr = 1e-4
wm = fo.weights.cpu().double()
is_app = wm > thr
idx = torch.arange(S)[None, :].expand(N, S)
first_app = torch.where(is_app, idx, torch.full_like(idx, S)).amin(1)
before = (idx < first_app[:, None]) & valid
after = (idx > first_app[:, None]) & valid & ~is_app
for name, msk in (("before first app sample", before), ("app samples", is_app & valid), ("behind, non-app", after)):
    if msk.any():
        nr = gs64[msk].norm()
        print(f"{name:26s} n={int(msk.sum()):7d} |g|={float(nr):.3e}  mine {float(dm[msk].norm()/nr):.2e}  o32 {float(do[msk].norm()/nr):.2e}"
              f"   slope mine {float((dm[msk]*gs64[msk]).sum()/(nr*nr)):+.2e}  o32 {float((do[msk]*gs64[msk]).sum()/(nr*nr)):+.2e}")
# per-ray relative scale error of the 'before' samples (R_i is shared by them)
num_m = (dm * gs64 * before).sum(1); num_o = (do * gs64 * before).sum(1); den = (gs64 * gs64 * before).sum(1)
ok = den > 0
rm, ro = (num_m[ok] / den[ok]), (num_o[ok] / den[ok])
if int(ok.sum()) > 0: print("per-ray scale error of dL/dsigma (before-surface samples): mine rms %.2e max %.2e ; o32 rms %.2e max %.2e" % (
    float(rm.pow(2).mean().sqrt()), float(rm.abs().max()), float(ro.pow(2).mean().sqrt()), float(ro.abs().max())))
# forward colour / weight agreement
rgb_m = fo.rgb.cpu().double()

r64, r32 = FWD["torch.float64"]["rgb"], FWD["torch.float32"]["rgb"]
am = is_app & valid
print("per-sample rgb (app samples): mine-vs-f64 max %.2e rms %.2e ; o32-vs-f64 max %.2e rms %.2e" % (
    float((rgb_m[am] - r64[am]).abs().max()), float((rgb_m[am] - r64[am]).pow(2).mean().sqrt()),
    float((r32[am] - r64[am]).abs().max()), float((r32[am] - r64[am]).pow(2).mean().sqrt())))
