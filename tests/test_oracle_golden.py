"""Pin the CPU oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only; runs under ``-m "not gpu"``."""
import numpy as np
import pytest
import torch

from oracle import nvfi_oracle as O
from tests.helpers import GOLDEN_SCENES, Golden, norm_rel_err, oracle_param_map, rel_err, scalar_loss

TOL = 1e-4          # north_star: within 1e-4 relative FP32
TOL_ORACLE = 2e-5   # the oracle uses the same library ops: expect (near-)identical results


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def g(request):
    return Golden(request.param)


def _check_outputs(out, case, tol=TOL_ORACLE):
    n = case["rgb"].shape[0]
    assert rel_err(out[0].reshape(n, -1), case["rgb"]) < tol
    assert rel_err(out[1].reshape(n), case["depth"]) < tol
    assert rel_err(out[2].reshape(n), case["acc"]) < tol
    assert rel_err(out[3].reshape(n, -1), case["weights"]) < tol
    assert rel_err(out[4].reshape(n, -1), case["mask_map"]) < tol


def test_step_size(g):
    sc = g.scene()
    step, n = O.step_size_and_nsamples(sc)
    assert n == g.meta["nSamples"]
    assert abs(float(step) - g.meta["stepSize"]) < 1e-9


@pytest.mark.parametrize("i", range(5))
def test_eval_render(g, i):
    sc = g.scene()
    case = g.case(f"eval{i}")
    o, d = g.rays()
    with torch.no_grad():
        out = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                       white_bg=bool(g.cfg.dataset.white_background), training=False)
    _check_outputs(out, case)


def test_eval_manual_bilerp_matches(g):
    """The restated grid_sample algorithm agrees with the library call the reference uses."""
    sc = g.scene()
    gen = torch.Generator().manual_seed(0)
    xyzt = torch.rand(500, 4, generator=gen) * 2.4 - 1.2
    a = O.density_feature(sc, xyzt, manual=False)
    b = O.density_feature(sc, xyzt, manual=True)
    assert rel_err(a, b) < 1e-5
    a = O.app_feature(sc, xyzt, manual=False)
    b = O.app_feature(sc, xyzt, manual=True)
    assert rel_err(a, b) < 1e-5


def test_transfer(g):
    sc = g.scene()
    case = g.case("transfer")
    o, d = g.rays()
    with torch.no_grad():
        out = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                       white_bg=bool(g.cfg.dataset.white_background), training=False, transfer_vel=True)
    _check_outputs(out, case)


def test_transfer_mask_field(g):
    sc = g.scene(alpha=True, mask_field=True)   # generated after updateAlphaMask: mask active
    case = g.case("transfer_mask")
    o, d = g.rays()
    with torch.no_grad():
        out = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                       white_bg=bool(g.cfg.dataset.white_background), training=False, transfer_vel=True)
    _check_outputs(out, case)
    assert float(out[4].abs().max()) > 0


@pytest.mark.parametrize("name", ["eval_alpha", "eval_alpha_extrap"])
def test_eval_alpha_mask(g, name):
    sc = g.scene(alpha=True)
    case = g.case(name)
    o, d = g.rays()
    with torch.no_grad():
        out = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                       white_bg=bool(g.cfg.dataset.white_background), training=False)
    _check_outputs(out, case)


@pytest.mark.parametrize("i", range(2))
def test_train_render_and_grads(g, i):
    sc = g.scene(requires_grad=True)
    case = g.case(f"train{i}")
    o, d = g.rays()
    out = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                   white_bg=bool(g.cfg.dataset.white_background), training=True,
                   jitter=torch.from_numpy(case["jitter"]),
                   random_bg=list(case["random_bg"]) if len(case["random_bg"]) else None)
    _check_outputs(out, case)
    loss = scalar_loss(out, g.loss_weights())
    assert abs(loss.item() - float(case["loss"])) < 1e-4 * max(1.0, abs(float(case["loss"])))
    loss.backward()
    pm = oracle_param_map(sc)
    checked = 0
    for k, v in case.items():
        if k.startswith("grad/"):
            p = pm[k[len("grad/"):]]
            assert p.grad is not None, k
            assert norm_rel_err(p.grad, v) < 1e-4, k
            checked += 1
        elif k.startswith("grad_sub/"):
            p = pm[k[len("grad_sub/"):]]
            assert norm_rel_err(p.grad.reshape(-1)[::7], v) < 1e-4, k
            checked += 1
    assert checked > 10


def test_ops(g):
    sc = g.scene()
    xyz, t, base = g.t("op/xyz"), g.t("op/t"), g.t("op/base")
    with torch.no_grad():
        assert torch.equal(O.keyframe_snap(sc, t), base)
        adv = O.integrate_pos(sc, xyz, t, base)
        assert rel_err(adv, g.t("op/adv")) < TOL_ORACLE
        xyzt = torch.cat([g.t("op/adv"), O.normalize_time_coord(sc, base)], -1)
        df = O.density_feature(sc, xyzt)
        assert rel_err(df[:, None], g.t("op/dfeat")) < TOL_ORACLE
        assert rel_err(O.feature2density(sc, df), g.t("op/sigma")) < TOL_ORACLE
        assert rel_err(O.app_feature(sc, xyzt), g.t("op/afeat")) < TOL_ORACLE
        xt = torch.cat([xyz, t], -1)
        assert rel_err(O.vel_full(sc, xt), g.t("op/vfull")) < TOL_ORACLE
        assert rel_err(O.gated_vel(sc, xt), g.t("op/vgate")) < TOL_ORACLE
        fwd = O.integrate_pos(sc, xyz, torch.zeros_like(t), t)
        assert rel_err(fwd, g.t("op/adv_fwd")) < TOL_ORACLE


def test_rk2_schedule_matches_tensor_loop(g):
    sc = g.scene()
    for t in (0.33, sc.tmax, sc.tmax + 0.25, 0.01):
        base = float(O.keyframe_snap(sc, torch.tensor([[t]], dtype=torch.float32)))
        sched = O.rk2_schedule(sc, t, base)
        x = torch.tensor([[0.1, -0.2, 0.3]])
        ref = O.integrate_pos(sc, x, torch.tensor([[t]], dtype=torch.float32), torch.tensor([[base]]))
        y = x.clone()
        for dt, tc in sched:
            dt_t = torch.tensor([[dt]], dtype=torch.float32)
            tc_t = torch.tensor([[tc]], dtype=torch.float32)
            v0 = O.gated_vel(sc, torch.cat([y, tc_t], -1))
            mid = y - 0.5 * dt_t * v0
            ynew = y - dt_t * O.gated_vel(sc, torch.cat([mid, tc_t - 0.5 * dt_t], -1))
            if sc.vel_gate == "sur" and bool(O.gate_outside(sc, ynew)):
                ynew = y
            y = ynew
        assert torch.allclose(y, ref, atol=1e-7)


def test_pde_loss(g):
    sc = g.scene(requires_grad=True)
    pu, t = g.t("pde/points_u"), g.t("pde/t")
    pts = pu * (sc.aabb[1] - sc.aabb[0]) + sc.aabb[0]
    loss = O.vel_loss(sc, O.normalize_coord(sc, pts), t)
    if bool(g.z["pde/empty"]):
        assert loss == 0.0
        return
    assert abs(loss.item() - float(g.z["pde/loss"])) < 2e-5 * max(1.0, abs(float(g.z["pde/loss"])))
    loss.backward()
    pm = oracle_param_map(sc)
    n = 0
    for k in g.keys("pde/grad/"):
        p = pm[k[len("pde/grad/"):]]
        assert norm_rel_err(p.grad, g.z[k]) < 1e-4, k
        n += 1
    assert n == 24      # both nets: 6 layers x (W,b) x 2


def test_raygen(g):
    H, W, focal = g.z["cam"]
    o, d = O.raygen(g.t("pose"), int(H), int(W), float(focal))
    assert torch.equal(o.reshape(-1, 3), g.t("rays_o"))
    assert torch.equal(d.reshape(-1, 3), g.t("rays_d"))
