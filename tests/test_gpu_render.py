"""GPU parity of the CUDA render path (through the C ABI) against (a) the golden vectors
produced by the reference itself and (b) the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): 1e-4 relative FP32 on values (relative to
max(|ref|, 1)); integer quantities (ray validity masks, appearance-mask membership up to
weights within 1 ulp of the threshold, ray directions) exact.
"""
import numpy as np
import pytest
import torch

from oracle import nvfi_oracle as O
from tests.helpers import GOLDEN_SCENES, Golden, build_model, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def g(request):
    return Golden(request.param)


@pytest.fixture(scope="module")
def model(g):
    return build_model(g)


def _renderer(model, g):
    from nvfi_b200 import models as M
    return M.Renderer(model, 0, 0, g.ray_chunk)


def _rays(g, dev="cuda"):
    from nvfi_b200 import models as M
    o, d = g.rays()
    return M.Ray(o.to(dev), d.to(dev), 0, 0)


def _check(out, case, tol=TOL):
    n = case["rgb"].shape[0]
    errs = dict(
        rgb=rel_err(out[0].reshape(n, -1).cpu(), case["rgb"]),
        depth=rel_err(out[1].reshape(n).cpu(), case["depth"]),
        acc=rel_err(out[2].reshape(n).cpu(), case["acc"]),
        weights=rel_err(out[3].reshape(n, -1).cpu(), case["weights"]),
        mask=rel_err(out[4].reshape(n, -1).cpu(), case["mask_map"]))
    assert all(v < tol for v in errs.values()), errs


def test_library_loads():
    from nvfi_b200 import _lib
    lib = _lib.load()
    assert lib.nvfi_abi_version() == _lib.ABI_VERSION


def test_raygen_exact(g):
    from nvfi_b200 import engine
    H, W, focal = g.z["cam"]
    ro, rd = engine.raygen(g.t("pose").cuda(), int(H), int(W), float(focal))
    assert torch.equal(ro.cpu(), g.t("rays_o"))
    assert torch.equal(rd.cpu(), g.t("rays_d"))
    ids = torch.tensor([0, 5, int(H * W) - 1, 17], dtype=torch.int64)
    ro2, rd2 = engine.raygen(g.t("pose").cuda(), int(H), int(W), float(focal), ids.cuda())
    assert torch.equal(rd2.cpu(), g.t("rays_d")[ids])


def test_field_ops(g, model):
    f = model.nvfi
    f.eval()
    xyz, t, base = g.t("op/xyz").cuda(), g.t("op/t").cuda(), g.t("op/base").cuda()
    adv = f.integrate_pos(xyz, t, base)
    assert rel_err(adv.cpu(), g.t("op/adv")) < TOL
    xyzt = torch.cat([g.t("op/adv").cuda(), f.normalize_time_coord(base)], -1)
    df = f.compute_densityfeature(xyzt)
    assert rel_err(df.cpu().reshape(-1), g.t("op/dfeat").reshape(-1)) < TOL
    assert rel_err(f.feature2density(df, {}).cpu().reshape(-1), g.t("op/sigma").reshape(-1)) < TOL
    assert rel_err(f.compute_appfeature(xyzt).cpu(), g.t("op/afeat")) < TOL
    xt = torch.cat([xyz, t], -1)
    assert rel_err(f.vel_net(xt).cpu(), g.t("op/vfull")) < TOL
    assert rel_err(f.vel(xt).cpu(), g.t("op/vgate")) < TOL
    fwd = f.integrate_pos(xyz, torch.zeros_like(t), t)
    assert rel_err(fwd.cpu(), g.t("op/adv_fwd")) < TOL


@pytest.mark.parametrize("i", range(5))
def test_eval_render(g, model, i):
    case = g.case(f"eval{i}")
    out = _renderer(model, g).render(float(case["t"]), _rays(g),
                                     white_background=bool(g.cfg.dataset.white_background), mode="test")
    _check(out, case)


def test_eval_masks_exact_vs_oracle(g, model):
    """valid mask and appearance-mask membership against the oracle on the same rays."""
    from nvfi_b200 import engine
    sc = g.scene()
    o, d = g.rays()
    case = g.case("eval1")
    t = float(case["t"])
    white = bool(g.cfg.dataset.white_background)
    model.nvfi.eval()
    with torch.no_grad():
        ref = O.render_chunk(sc, t, o, d, white_bg=white, training=False, return_aux=True)
    out = engine.render_forward(model.nvfi.binding, o.cuda(), d.cuda(), t, white_bg=white, training=False,
                                ray_chunk=o.shape[0], want_stats=True)
    aux = ref[5]
    assert torch.equal(out.valid.cpu().bool(), aux["valid"])
    thr = float(sc.ray_march_weight_thres)
    app = (out.weights > thr).cpu()
    flips = (app != aux["app_mask"])
    near_thr = (ref[3] - thr).abs() < 1e-6
    assert not bool((flips & ~near_thr).any())
    stats = out.stats.cpu()
    assert int(stats[0]) == int(aux["valid"].sum())
    assert int(stats[2]) == int(app.sum())
    v = aux["valid"].clone()
    if out.ray_term is not None:     # early ray termination: samples behind the end of a ray are not advected
        S = v.shape[1]
        v &= torch.arange(S)[None, :] < out.ray_term.cpu()[:, None]
        assert int(stats[1]) == int(v.sum())                 # advected = in-box samples in front of the termination
        # ... and everything that was skipped has zero weight in the reference as well
        assert float(ref[3][aux["valid"] & ~v].abs().max() if bool((aux["valid"] & ~v).any()) else 0.0) == 0.0
    assert rel_err(out.x_adv.cpu()[v], aux["xyz_adv"][v]) < TOL


def test_transfer(g, model):
    case = g.case("transfer")
    out = _renderer(model, g).render(float(case["t"]), _rays(g),
                                     white_background=bool(g.cfg.dataset.white_background), mode="test",
                                     transfer_vel=True)
    _check(out, case)


@pytest.mark.parametrize("name", ["eval_alpha", "eval_alpha_extrap"])
def test_eval_alpha_mask(g, name):
    m = build_model(g, alpha=True)
    case = g.case(name)
    out = _renderer(m, g).render(float(case["t"]), _rays(g),
                                 white_background=bool(g.cfg.dataset.white_background), mode="test")
    _check(out, case)


def test_transfer_mask_field(g):
    m = build_model(g, alpha=True, mask_field=True)
    case = g.case("transfer_mask")
    out = _renderer(m, g).render(float(case["t"]), _rays(g),
                                 white_background=bool(g.cfg.dataset.white_background), mode="test",
                                 transfer_vel=True)
    _check(out, case)
    assert float(out[4].abs().max()) > 0


@pytest.mark.parametrize("i", range(2))
def test_train_forward(g, model, i):
    """Training-mode forward (stratified jitter, per-chunk random background)."""
    case = g.case(f"train{i}")
    o, d = g.rays()
    model.nvfi.train()
    bg = torch.from_numpy(case["random_bg"].astype(np.uint8)) if len(case["random_bg"]) else None
    white = bool(g.cfg.dataset.white_background)
    with torch.no_grad():
        out = model.nvfi.render_rays(float(case["t"]), o.cuda(), d.cuda(), white_bg=white,
                                     ray_chunk=g.ray_chunk, jitter=torch.from_numpy(case["jitter"]),
                                     chunk_bg=bg)
    _check(out, case)


def test_train_rng_stream_matches_reference(g, model):
    """With the same torch seed the host draws (jitter per chunk, then background) are the
    ones the reference made (tests/golden/make_golden.py seeds spec.seed + 7 + i)."""
    case = g.case("train0")
    o, d = g.rays()
    model.nvfi.train()
    torch.manual_seed(g.meta["seed"] + 7)
    with torch.no_grad():
        out = model.nvfi.render_rays(float(case["t"]), o.cuda(), d.cuda(),
                                     white_bg=bool(g.cfg.dataset.white_background), ray_chunk=g.ray_chunk)
    _check(out, case)


def test_empty_and_ragged(g, model):
    from nvfi_b200 import models as M
    model.nvfi.eval()
    o, d = g.rays()
    r = M.Renderer(model, 0, 0, 7)      # ragged chunking: 400 rays in chunks of 7
    case = g.case("eval1")
    out = r.render(float(case["t"]), M.Ray(o.cuda(), d.cuda(), 0, 0),
                   white_background=bool(g.cfg.dataset.white_background), mode="test")
    _check(out, case)
    e = torch.zeros(0, 3, device="cuda")
    out = r.render(0.1, M.Ray(e, e, 0, 0), white_background=True, mode="test")
    assert out[0].shape == (0, 3) and out[3].shape[0] == 0
