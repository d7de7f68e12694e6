"""The velocity-MLP GEMMs run either on the FP32 SIMT verification path or on the tcgen05
tensor cores: FP16 2-way operand split (hi + lo = 22 mantissa bits, 3 MMAs per GEMM; the default),
the round-1 3-term TF32 split, or a single TF32 pass.  Every mode is held against the same golden
vectors; the tolerance is the north-star 1e-4 for 'simt', 'f16x3' and 'tf32x3' and a stated, looser
bound for the single-pass mode."""
import pytest
import torch

from tests.helpers import GOLDEN_SCENES, Golden, build_model, rel_err

pytestmark = pytest.mark.gpu
TOL = {"simt": 1e-4, "f16x3": 1e-4, "tf32x3": 1e-4, "tf32": 2e-2}


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def g(request):
    return Golden(request.param)


@pytest.fixture(scope="module")
def model(g):
    return build_model(g)


@pytest.fixture(params=["simt", "f16x3", "tf32x3", "tf32"])
def mode(request):
    from nvfi_b200 import engine
    prev = engine.set_mlp_mode(request.param)
    yield request.param
    engine.set_mlp_mode(prev)


def test_default_mode_is_tensor_core():
    from nvfi_b200 import engine
    import os
    if "NVFI_MLP_MODE" not in os.environ:
        assert engine.get_mlp_mode() == "f16x3"


def test_velocity_and_advection(g, model, mode):
    f = model.nvfi
    f.eval()
    tol = TOL[mode]
    xyz, t, base = g.t("op/xyz").cuda(), g.t("op/t").cuda(), g.t("op/base").cuda()
    xt = torch.cat([xyz, t], -1)
    assert rel_err(f.vel_net(xt).cpu(), g.t("op/vfull")) < tol
    assert rel_err(f.vel(xt).cpu(), g.t("op/vgate")) < tol
    assert rel_err(f.integrate_pos(xyz, t, base).cpu(), g.t("op/adv")) < tol
    assert rel_err(f.integrate_pos(xyz, torch.zeros_like(t), t).cpu(), g.t("op/adv_fwd")) < tol


def test_velocity_ragged_sizes(g, model, mode):
    """Tile tails: n not a multiple of 128, n < 128, n == 0."""
    f = model.nvfi
    f.eval()
    xyz, t = g.t("op/xyz").cuda(), g.t("op/t").cuda()
    xt = torch.cat([xyz, t], -1)
    full = f.vel_net(xt)
    for n in (1, 127, 129, min(300, xt.shape[0])):
        part = f.vel_net(xt[:n].contiguous())
        assert torch.equal(part, full[:n]), n     # same tile arithmetic whatever the tail
    assert f.vel_net(xt[:0].contiguous()).shape == (0, 6)


@pytest.mark.parametrize("i", (1, 2, 4))
def test_eval_render_modes(g, model, mode, i):
    from nvfi_b200 import models as M
    case = g.case(f"eval{i}")
    o, d = g.rays()
    r = M.Renderer(model, 0, 0, g.ray_chunk)
    out = r.render(float(case["t"]), M.Ray(o.cuda(), d.cuda(), 0, 0),
                   white_background=bool(g.cfg.dataset.white_background), mode="test")
    n = case["rgb"].shape[0]
    tol = TOL[mode]
    assert rel_err(out[0].reshape(n, -1).cpu(), case["rgb"]) < tol
    assert rel_err(out[1].reshape(n).cpu(), case["depth"]) < tol
    assert rel_err(out[2].reshape(n).cpu(), case["acc"]) < tol
    assert rel_err(out[3].reshape(n, -1).cpu(), case["weights"]) < tol


def test_modes_agree_on_large_batch(model, g):
    """A batch spanning many CTAs and ring wrap-arounds: tensor-core result vs SIMT result."""
    from nvfi_b200 import engine
    f = model.nvfi
    f.eval()
    gen = torch.Generator().manual_seed(5)
    n = 128 * 700 + 37
    xt = torch.cat([torch.rand(n, 3, generator=gen) * 1.8 - 0.9, torch.rand(n, 1, generator=gen)], -1).cuda()
    prev = engine.set_mlp_mode("simt")
    try:
        ref = f.vel_net(xt)
        engine.set_mlp_mode("tf32x3")
        got = f.vel_net(xt)
        engine.set_mlp_mode("f16x3")
        got_h = f.vel_net(xt)
    finally:
        engine.set_mlp_mode(prev)
    assert rel_err(got.cpu(), ref.cpu()) < 2e-5
    assert rel_err(got_h.cpu(), ref.cpu()) < 2e-5


@pytest.mark.parametrize("i", range(2))
def test_train_grads_modes(g, mode, i):
    """Backward pass (RK2 adjoint through the velocity MLP) in every arithmetic mode against the
    gradients the reference itself produced."""
    import numpy as np
    from tests.helpers import norm_rel_err, scalar_loss
    case = g.case(f"train{i}")
    model = build_model(g, requires_grad=True)
    f = model.nvfi
    f.train()
    o, d = g.rays()
    bg = torch.from_numpy(case["random_bg"].astype(np.uint8)) if len(case["random_bg"]) else None
    out = f.render_rays(float(case["t"]), o.cuda(), d.cuda(), white_bg=bool(g.cfg.dataset.white_background),
                        ray_chunk=g.ray_chunk, jitter=torch.from_numpy(case["jitter"]), chunk_bg=bg)
    lw = {k: v.cuda() for k, v in g.loss_weights().items()}
    scalar_loss(out, lw).backward()
    params = dict(f.named_parameters())
    tol = TOL[mode]
    errs = {}
    for k, v in case.items():
        if k.startswith("grad/vel_net.weight_net"):
            name = k[len("grad/"):]
            assert params[name].grad is not None, name
            errs[name] = norm_rel_err(params[name].grad.cpu(), v)
    if float(case["t"]) and errs:
        bad = {k: v for k, v in errs.items() if not v < tol}
        assert not bad, bad
        assert len(errs) == 12
