"""Parity at BASELINE.json's full sizes (configs[1]: bat.yaml, 199^3 grid, K = 16, 192 samples per
ray, 800x800 frame) through size-independent properties — the CPU oracle cannot render these
sizes in seconds, so the checks are identities of the algorithm itself
(models/tensorf_keyframe.py:613-755, models/tensorf_model_utils.py:186-197):

 * rays are independent: a frame rendered as one call equals the same rays rendered in shards,
   bit for bit (this is what the multi-GPU ray sharding relies on);
 * acc = sum(weights), depth = sum(w z) + (1 - acc) far with z recomputed from the sampler
   formula, weights in [0, 1], T monotone: alpha-composite identities on the returned weights;
 * a key-frame time renders without advection (x_adv == sample position);
 * the backward pass is linear in the upstream gradient (additivity over two losses, scaling);
 * the tensor-core backward agrees with the FP32 SIMT backward (an independent implementation,
   itself held against the reference's autograd on the golden scenes) on tens of thousands of
   rays, within the 1e-4 gate.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 800
T_RENDER = 0.33
CHUNK = 2048


@pytest.fixture(scope="module")
def scene():
    from nvfi_b200.scenes import build_scene, frame_rays
    cfg, nv, _ = build_scene("bat", grid=(199, 199, 199), step_ratio=1.79)
    assert nv.nvfi.nSamples == 192
    o, d = frame_rays(H, W)
    return cfg, nv, o, d


def _band(o, d, rows, r0=None):
    r0 = (H - rows) // 2 if r0 is None else r0
    sl = slice(r0 * W, (r0 + rows) * W)
    return o[sl].contiguous().cuda(), d[sl].contiguous().cuda()


def test_full_frame_eval_identities(scene):
    cfg, nv, o, d = scene
    f = nv.nvfi
    f.eval()
    oo, dd = o.cuda(), d.cuda()
    with torch.no_grad():
        rgb, depth, acc, w, _ = f.render_rays(T_RENDER, oo, dd, white_bg=True, ray_chunk=CHUNK)
    n = oo.shape[0]
    assert rgb.shape == (n, 3) and w.shape == (n, 192)
    assert torch.isfinite(rgb).all() and torch.isfinite(w).all()
    assert float(w.min()) >= 0.0 and float(w.max()) <= 1.0
    assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0
    # acc = sum(weights) (models/tensorf_keyframe.py:737)
    assert float((w.sum(-1) - acc).abs().max()) < 2e-5
    # depth = sum(w z) + (1 - acc) far, z = near + step (i)  (camera inside-test true for this rig)
    near, far = float(cfg.dataset.near), float(cfg.dataset.far)
    z = near + float(f.stepSize) * torch.arange(192, device="cuda", dtype=torch.float32)
    ref_depth = (w * z[None, :]).sum(-1) + (1.0 - acc) * far
    assert float(((depth - ref_depth).abs() / ref_depth.abs().clamp_min(1.0)).max()) < 2e-5
    # something was rendered: the cube covers the middle of the frame
    assert 0.2 < float(acc.mean()) < 0.95


def test_shards_equal_whole_bitwise(scene):
    """Ray independence: contiguous chunk-aligned shards == the whole call."""
    cfg, nv, o, d = scene
    f = nv.nvfi
    f.eval()
    oo, dd = _band(o, d, 128)          # 102 400 rays = 50 chunks
    n = oo.shape[0]
    with torch.no_grad():
        whole = f.render_rays(T_RENDER, oo, dd, white_bg=True, ray_chunk=CHUNK)
        cut = (n // CHUNK // 3) * CHUNK
        a = f.render_rays(T_RENDER, oo[:cut], dd[:cut], white_bg=True, ray_chunk=CHUNK)
        b = f.render_rays(T_RENDER, oo[cut:], dd[cut:], white_bg=True, ray_chunk=CHUNK)
    for k in range(4):
        assert torch.equal(whole[k], torch.cat([a[k], b[k]], 0)), k


def test_train_jitter_shards_equal_whole_bitwise(scene):
    cfg, nv, o, d = scene
    f = nv.nvfi
    f.train()
    oo, dd = _band(o, d, 64)
    n = oo.shape[0]
    jit = torch.rand(n, 1, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        whole = f.render_rays(T_RENDER, oo, dd, white_bg=True, ray_chunk=CHUNK, jitter=jit)
        cut = 7 * CHUNK
        a = f.render_rays(T_RENDER, oo[:cut], dd[:cut], white_bg=True, ray_chunk=CHUNK, jitter=jit[:cut])
        b = f.render_rays(T_RENDER, oo[cut:], dd[cut:], white_bg=True, ray_chunk=CHUNK, jitter=jit[cut:])
    for k in range(4):
        assert torch.equal(whole[k], torch.cat([a[k], b[k]], 0)), k


def test_keyframe_time_needs_no_advection(scene):
    """t on a key frame: `isclose(t, base)` -> no sample is advected (models/tensorf_keyframe.py:683-692)."""
    from nvfi_b200 import engine
    cfg, nv, o, d = scene
    f = nv.nvfi
    f.eval()
    oo, dd = _band(o, d, 32)
    out = engine.render_forward(f.binding, oo, dd, 0.25, white_bg=True, training=False, jitter=None,
                                ray_chunk=CHUNK, want_stats=True)
    n_valid, n_adv = int(out.stats[0]), int(out.stats[1])
    assert n_valid > 0 and n_adv == 0
    prev = engine.set_early_termination(False)      # every in-box sample is advected, as in the reference
    try:
        out2 = engine.render_forward(f.binding, oo, dd, T_RENDER, white_bg=True, training=False, jitter=None,
                                     ray_chunk=CHUNK, want_stats=True)
    finally:
        engine.set_early_termination(prev)
    assert int(out2.stats[0]) == n_valid and int(out2.stats[1]) == n_valid
    out3 = engine.render_forward(f.binding, oo, dd, T_RENDER, white_bg=True, training=False, jitter=None,
                                 ray_chunk=CHUNK, want_stats=True)     # with early ray termination: fewer, same mask
    assert int(out3.stats[0]) == n_valid and 0 < int(out3.stats[1]) < n_valid
    assert torch.equal(out3.valid, out2.valid) and torch.equal(out3.weights, out2.weights)


def _grads(nv, loss_fn, oo, dd, jit):
    f = nv.nvfi
    nv.zero_grad(set_to_none=True)
    rgb, depth, acc, w, _ = f.render_rays(T_RENDER, oo, dd, white_bg=True, ray_chunk=CHUNK, jitter=jit)
    loss_fn(rgb, depth, acc, w).backward()
    return {k: p.grad.detach().clone() for k, p in nv.named_parameters() if p.grad is not None}


def _nrel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_backward_is_linear_in_the_upstream_gradient(scene):
    cfg, nv, o, d = scene
    nv.requires_grad_(True)
    nv.nvfi.train()
    oo, dd = _band(o, d, 24)
    n = oo.shape[0]
    gen = torch.Generator().manual_seed(11)
    jit = torch.rand(n, 1, generator=gen)
    ta = torch.rand(n, 3, generator=gen).cuda()
    tb = torch.rand(n, device="cuda")
    la = lambda rgb, depth, acc, w: torch.nn.functional.mse_loss(rgb, ta)
    lb = lambda rgb, depth, acc, w: 0.1 * ((depth - 4.0 * tb) ** 2).mean() + 0.3 * acc.mean()
    ga = _grads(nv, la, oo, dd, jit)
    gb = _grads(nv, lb, oo, dd, jit)
    gab = _grads(nv, lambda *a: la(*a) + lb(*a), oo, dd, jit)
    g2 = _grads(nv, lambda *a: 2.0 * la(*a), oo, dd, jit)
    assert set(ga) == set(gab) and len(ga) >= 30
    for k in gab:
        # three FP32 passes with different summation orders (atomic reductions): the 1e-4 gate
        assert _nrel(ga[k] + gb[k], gab[k]) < 1e-4, k
        # scaling by a power of two changes no rounding, only the reduction order
        assert _nrel(2.0 * ga[k], g2[k]) < 2e-5, k
    # gradient reach is the reference's
    assert not any("a_weight_net" in k or "basis_mat_density" in k for k in ga)
    assert any(k.startswith("nvfi.vel_net.weight_net") for k in ga)
    nv.requires_grad_(False)


def test_tensor_core_backward_matches_simt_backward_at_scale(scene):
    from nvfi_b200 import engine
    cfg, nv, o, d = scene
    nv.requires_grad_(True)
    nv.nvfi.train()
    oo, dd = _band(o, d, 40)           # 32 000 rays, ~3.8 M advected samples
    n = oo.shape[0]
    gen = torch.Generator().manual_seed(3)
    jit = torch.rand(n, 1, generator=gen)
    tgt = torch.rand(n, 3, generator=gen).cuda()
    loss = lambda rgb, depth, acc, w: torch.nn.functional.mse_loss(rgb, tgt)
    g_tc = _grads(nv, loss, oo, dd, jit)
    prev = engine.set_mlp_mode("simt")
    try:
        g_simt = _grads(nv, loss, oo, dd, jit)
    finally:
        engine.set_mlp_mode(prev)
    assert set(g_tc) == set(g_simt)
    worst = {k: _nrel(g_tc[k], g_simt[k]) for k in g_tc}
    bad = {k: v for k, v in worst.items() if not v < 1e-4}
    assert not bad, bad
    nv.requires_grad_(False)


def test_multi_step_backward_tc_matches_simt(scene):
    """t = 1.0 > tmax: 10 RK2 steps per sample (models/tensorf_keyframe.py:592-609) — the tile
    program of the tensor-core backward runs its forward sweep, two weight rings and 10 reverse
    steps; compared with the FP32 SIMT backward."""
    from nvfi_b200 import engine
    cfg, nv, o, d = scene
    nv.requires_grad_(True)
    nv.nvfi.train()
    oo, dd = _band(o, d, 4)            # 3 200 rays
    n = oo.shape[0]
    gen = torch.Generator().manual_seed(9)
    jit = torch.rand(n, 1, generator=gen)
    tgt = torch.rand(n, 3, generator=gen).cuda()

    def grads():
        nv.zero_grad(set_to_none=True)
        rgb, *_ = nv.nvfi.render_rays(1.0, oo, dd, white_bg=True, ray_chunk=CHUNK, jitter=jit)
        torch.nn.functional.mse_loss(rgb, tgt).backward()
        return {k: p.grad.detach().clone() for k, p in nv.named_parameters() if p.grad is not None}

    g_tc = grads()
    prev = engine.set_mlp_mode("simt")
    try:
        g_simt = grads()
    finally:
        engine.set_mlp_mode(prev)
    assert set(g_tc) == set(g_simt)
    errs = {k: _nrel(g_tc[k], g_simt[k]) for k in g_tc}
    # Ten RK2 steps amplify the last-bit differences of the two forward paths (x_adv agrees to
    # 2e-7).  Everything downstream of the ReLUs of MLPRender_PE (its first two layers, basis_mat,
    # the appearance planes) has a DIScontinuous gradient in x_adv — a hidden unit whose
    # pre-activation crosses 0 switches its whole gradient — so those tensors agree only to ~1e-3
    # (the reference's own FP32 autograd is off its FP64 value by the same 1e-3 there, see
    # test_extrapolated_time_gradients_vs_fp64_oracle); everything else meets the 1e-4 gate.
    kinked = ("app_plane", "basis_mat", "renderModule.mlp.0", "renderModule.mlp.2")
    bad = {k: v for k, v in errs.items() if not v < (5e-3 if any(s_ in k for s_ in kinked) else 1e-4)}
    assert not bad, bad
    nv.requires_grad_(False)


def test_extrapolated_time_gradients_vs_fp64_oracle():
    """Train step at t = 1.0 (10 RK2 steps) on a 48^3 scene against the oracle evaluated in FLOAT64.
    The FP32 autograd of the reference algorithm is itself 1e-3 away from this value for the tensors
    behind the ReLUs of MLPRender_PE (measured: app planes 1.0e-3, basis_mat 6e-4, vel_net 1.6e-3);
    the CUDA path has to meet the 1e-4 gate against the FP64 value."""
    from nvfi_b200.scenes import build_scene, frame_rays
    from oracle import nvfi_oracle as O
    from oracle.scene_io import scene_from_state
    grid = (48, 48, 48)
    cfg, nv, sd = build_scene("bat", grid=grid, max_n_samples=64)
    f = nv.nvfi
    nv.requires_grad_(True)
    f.train()
    o, d = frame_rays(800, 800, crop=(368, 368, 64, 64))
    gen = torch.Generator().manual_seed(0)
    n = 1024
    sel = torch.randperm(o.shape[0], generator=gen)[:n]
    oo, dd = o[sel].contiguous(), d[sel].contiguous()
    jit = torch.rand(n, 1, generator=gen)
    tgt = torch.rand(n, 3, generator=gen)
    rgb, *_ = f.render_rays(1.0, oo.cuda(), dd.cuda(), white_bg=True, ray_chunk=CHUNK, jitter=jit)
    torch.nn.functional.mse_loss(rgb, tgt.cuda()).backward()
    got = dict(nv.named_parameters())

    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        sc = scene_from_state(cfg, list(grid), int(cfg.nvfi.num_keyframes), sd, requires_grad=True)

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
        ref = O.render_chunk(sc, 1.0, oo.double(), dd.double(), white_bg=True, training=True, jitter=jit.double())
        torch.nn.functional.mse_loss(ref[0], tgt.double()).backward()
    finally:
        torch.set_default_dtype(prev)
    pairs = {"nvfi.density_plane_space.0": sc.density_plane_space[0], "nvfi.app_plane_space.0": sc.app_plane_space[0],
             "nvfi.app_plane_time.1": sc.app_plane_time[1], "nvfi.basis_mat.weight": sc.basis_mat,
             "nvfi.renderModule.mlp.0.weight": sc.render_mlp[0][0], "nvfi.renderModule.mlp.2.weight": sc.render_mlp[1][0],
             "nvfi.renderModule.mlp.4.weight": sc.render_mlp[2][0], "nvfi.vel_net.weight_net.1.weight": sc.vel_net[0][0],
             "nvfi.vel_net.weight_net.4.0.weight": sc.vel_net[2][0]}
    errs = {k: _nrel(got[k].grad.double().cpu(), r.grad) for k, r in pairs.items()}
    print(errs)
    assert float((rgb.detach().double().cpu() - ref[0].detach()).abs().max()) < 1e-4
    bad = {k: v for k, v in errs.items() if not v < 1e-4}
    assert not bad, bad


def test_early_termination_changes_nothing():
    """Early ray termination (NvfiRenderBuffers.ray_T, engine.set_early_termination): at the headline size the
    outputs of a train step and every gradient are the same with the samples behind a saturated surface
    evaluated (the reference's behaviour) or skipped; only the amount of work differs."""
    from nvfi_b200 import engine
    from nvfi_b200.scenes import build_scene, frame_rays
    cfg, nv, _ = build_scene("bat", grid=(199, 199, 199), step_ratio=1.79)
    f = nv.nvfi
    o, d = frame_rays(800, 800, theta=30.0)
    sel = slice(300 * 800, 300 * 800 + 16384)
    oo, dd = o[sel].cuda(), d[sel].cuda()
    gen = torch.Generator().manual_seed(4)
    jit = torch.rand(oo.shape[0], 1, generator=gen)
    tgt = torch.rand(oo.shape[0], 3, generator=gen).cuda()
    nv.requires_grad_(True)
    f.train()
    res = {}
    for on in (False, True):
        prev = engine.set_early_termination(on)
        try:
            nv.zero_grad(set_to_none=True)
            out = engine.render_forward(f.binding, oo, dd, 0.33, white_bg=True, training=True, jitter=jit,
                                        ray_chunk=2048, want_stats=True)
            stats = out.stats.tolist()
            rgb, depth, acc, w, _ = f.render_rays(0.33, oo, dd, white_bg=True, ray_chunk=2048, jitter=jit)
            torch.nn.functional.mse_loss(rgb, tgt).backward()
            res[on] = (stats, rgb.detach(), depth.detach(), acc.detach(), w.detach(),
                       {k: p.grad.clone() for k, p in nv.named_parameters() if p.grad is not None})
        finally:
            engine.set_early_termination(prev)
    s_off, s_on = res[False][0], res[True][0]
    assert s_off[0] == s_on[0] and s_off[1] == s_off[0]        # same in-box samples; all advected without termination
    assert s_on[1] < 0.9 * s_off[1]                             # the cube saturates: a good part of the work is skipped
    print(f"[early termination] advected samples {s_on[1]} of {s_off[1]} ({s_on[1] / s_off[1]:.2f})")
    assert torch.equal(res[False][4], res[True][4])             # weights: bit-identical (same scan, carried exactly)
    for k in (1, 2, 3):     # rgb (through the background term 1 - acc) / depth / acc: the sums are formed wave by wave
        assert float((res[False][k] - res[True][k]).abs().max()) < 2e-6 * max(1.0, float(res[False][k].abs().max()))
    assert res[False][5].keys() == res[True][5].keys()
    for k, g0 in res[False][5].items():
        g1 = res[True][5][k]
        assert float((g0 - g1).norm()) <= 2e-6 * float(g0.norm()) + 1e-30, k
