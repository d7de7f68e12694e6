"""GPU parity of the hand-written backward pass: gradients of a random linear functional
of (rgb, depth, acc, weights) with respect to every parameter the reference's autograd
reaches, against the gradients the reference itself produced (golden fixtures) and
against the CPU oracle's autograd.  Tolerance: ||g - g_ref|| / ||g_ref|| < 1e-4 per tensor
(1e-4 relative FP32, north_star); outputs as in test_gpu_render."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN_SCENES, Golden, build_model, norm_rel_err, rel_err, scalar_loss

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def g(request):
    return Golden(request.param)


def _run_train(g, case, lw_keys=("wr", "wd", "wa", "ww")):
    model = build_model(g, requires_grad=True)
    f = model.nvfi
    f.train()
    o, d = g.rays()
    bg = torch.from_numpy(case["random_bg"].astype(np.uint8)) if len(case["random_bg"]) else None
    out = f.render_rays(float(case["t"]), o.cuda(), d.cuda(),
                        white_bg=bool(g.cfg.dataset.white_background), ray_chunk=g.ray_chunk,
                        jitter=torch.from_numpy(case["jitter"]), chunk_bg=bg)
    lw = {k: v.cuda() for k, v in g.loss_weights().items()}
    for k in lw:
        if k not in lw_keys:
            lw[k] = torch.zeros_like(lw[k])
    loss = scalar_loss(out, lw)
    loss.backward()
    return model, out, loss


@pytest.mark.parametrize("i", range(2))
def test_train_grads_vs_reference(g, i):
    case = g.case(f"train{i}")
    model, out, loss = _run_train(g, case)
    assert abs(loss.item() - float(case["loss"])) < 1e-4 * max(1.0, abs(float(case["loss"])))
    params = dict(model.nvfi.named_parameters())
    checked, errs = 0, {}
    for k, v in case.items():
        if k.startswith("grad/"):
            name = k[len("grad/"):]
            p = params[name]
            assert p.grad is not None, name
            errs[name] = norm_rel_err(p.grad.cpu(), v)
            checked += 1
        elif k.startswith("grad_sub/"):
            name = k[len("grad_sub/"):]
            p = params[name]
            assert p.grad is not None, name
            errs[name] = norm_rel_err(p.grad.cpu().reshape(-1)[::7], v)
            full = float(p.grad.double().norm())
            ref = float(case["grad_norm/" + name])
            assert abs(full - ref) < 1e-4 * max(ref, 1e-30), (name, full, ref)
            checked += 1
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    assert checked > 10
    # parameters the render never reaches keep grad None like the reference
    for name, p in params.items():
        if "a_weight_net" in name or "basis_mat_density" in name:
            assert p.grad is None, name
    if float(case["t"]) and any(k.startswith("grad/vel_net.weight_net") for k in case):
        assert params["vel_net.weight_net.1.weight"].grad is not None


def test_rgb_only_loss_matches_oracle_autograd(g):
    """The shipped trainer's loss touches rgb only (train_nvfi.py:162): g_depth / g_acc /
    g_weights are absent.  Reference values come from the oracle's autograd."""
    from oracle import nvfi_oracle as O
    from tests.helpers import oracle_param_map
    case = g.case("train0")
    model, out, loss = _run_train(g, case, lw_keys=("wr",))
    sc = g.scene(requires_grad=True)
    o, d = g.rays()
    ref = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                   white_bg=bool(g.cfg.dataset.white_background), training=True,
                   jitter=torch.from_numpy(case["jitter"]),
                   random_bg=list(case["random_bg"]) if len(case["random_bg"]) else None)
    lw = g.loss_weights()
    (ref[0] * lw["wr"]).sum().backward()
    pm = oracle_param_map(sc)
    params = dict(model.nvfi.named_parameters())
    errs = {}
    for name, p in pm.items():
        if p.grad is None or "a_weight_net" in name:
            continue
        if float(p.grad.abs().max()) == 0:
            continue
        errs[name] = norm_rel_err(params[name].grad.cpu(), p.grad)
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    assert len(errs) > 10


def test_two_renders_accumulate(g):
    """Two renders before one backward (the trainer's pattern, train_nvfi.py:158-204):
    gradients add up."""
    model = build_model(g, requires_grad=True)
    f = model.nvfi
    f.train()
    o, d = g.rays()
    c0, c1 = g.case("train0"), g.case("train1")
    white = bool(g.cfg.dataset.white_background)
    lw = {k: v.cuda() for k, v in g.loss_weights().items()}
    total = 0
    for c in (c0, c1):
        bg = torch.from_numpy(c["random_bg"].astype(np.uint8)) if len(c["random_bg"]) else None
        out = f.render_rays(float(c["t"]), o.cuda(), d.cuda(), white_bg=white, ray_chunk=g.ray_chunk,
                            jitter=torch.from_numpy(c["jitter"]), chunk_bg=bg)
        total = total + scalar_loss(out, lw)
    total.backward()
    params = dict(f.named_parameters())
    name = "density_plane_space.0"
    want = c0["grad_sub/" + name] + c1["grad_sub/" + name]
    assert norm_rel_err(params[name].grad.cpu().reshape(-1)[::7], want) < TOL
    name = "renderModule.mlp.0.weight" if g.cfg.nvfi.shadingMode == "MLP_PE" else "basis_mat.weight"
    want = c0["grad/" + name] + c1["grad/" + name]
    assert norm_rel_err(params[name].grad.cpu(), want) < TOL
