"""CUDA path against the oracle AT the headline size (BASELINE.json configs[1]: bat.yaml, 199^3 grid,
K = 16, 192 samples per ray via step_ratio 1.79, the 800x800 synthetic camera).  The oracle renders
2 048-ray chunks of that scene in about a second each, so the parity gate of the small golden scenes
is applied here directly (not only the size-independent identities of test_gpu_fullsize.py):

  * three chunks of the frame — top (mostly empty rays), the row where the cube's silhouette starts,
    and the centre (every ray hits the cube) — outputs at t = 0.33 (one RK2 step), 0.75 (key frame,
    no advection) and 1.0 (extrapolation, 10 RK2 steps);
  * train gradients at t = 0.33 for every parameter the reference's autograd reaches.

Both error figures are reported and gated at 1e-4: the element-wise error with the floor max(|ref|, 1)
and the per-tensor relative L2 norm (tests/helpers.py:assert_close).  One documented exception: on
near-empty rays (top of the frame: acc ~ 5e-3) every alpha is 1 - exp(-x) with x ~ 4e-5, which FP32
evaluates to only ~1e-3 RELATIVE accuracy — the reference's own FP32 result is that far from the exact
value.  There the relative-norm gate is replaced by "not worse than the reference's FP32 arithmetic":
the CUDA result's distance to a FLOAT64 evaluation of the same algorithm must not exceed twice the FP32
oracle's own distance to it (the absolute gate still applies)."""
import pytest
import torch

from tests.helpers import assert_close, norm_rel_err, oracle_param_map

pytestmark = pytest.mark.gpu

GRID = (199, 199, 199)
H = W = 800
CHUNK = 2048
STARTS = {"top": 0, "silhouette": 230 * W, "centre": 400 * W + 37 * 8}


@pytest.fixture(scope="module")
def scene():
    from nvfi_b200.scenes import build_scene, frame_rays
    from oracle.scene_io import scene_from_state
    cfg, nv, sd = build_scene("bat", grid=GRID, step_ratio=1.79)
    assert nv.nvfi.nSamples == 192
    o, d = frame_rays(H, W, theta=30.0)
    K = int(cfg.nvfi.num_keyframes)
    return cfg, nv, sd, o, d, (lambda rg=False: scene_from_state(cfg, list(GRID), K, sd, requires_grad=rg))


def _oracle_f64(cfg, sd, t, oo, dd):
    """The oracle evaluated in float64 on the same parameters (the exact value of the algorithm)."""
    from oracle import nvfi_oracle as O
    from oracle.scene_io import scene_from_state
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        sc = scene_from_state(cfg, list(GRID), int(cfg.nvfi.num_keyframes), sd)

        def cast(x):
            if isinstance(x, torch.Tensor):
                return x.detach().double()
            if isinstance(x, (list, tuple)):
                return type(x)(cast(y) for y in x)
            return x
        for name in ("density_plane_space", "density_plane_time", "app_plane_space", "app_plane_time",
                     "basis_mat", "render_mlp", "vel_net", "acc_net"):
            setattr(sc, name, cast(getattr(sc, name)))
        sc.aabb = sc.aabb.double()
        with torch.no_grad():
            return O.render_chunk(sc, t, oo.double(), dd.double(), white_bg=True, training=False)
    finally:
        torch.set_default_dtype(prev)


def _gate(got, ref, what, f64=None):
    """(max floored, rel norm) of CUDA vs the FP32 oracle; see the module docstring for the exception: a
    figure above 1e-4 is accepted only if the CUDA result is not further from the FLOAT64 evaluation than
    twice the FP32 oracle's own distance (+2e-5 absolute slack on the floored figure)."""
    from tests.helpers import rel_err
    e_max, e_norm = rel_err(got, ref), norm_rel_err(got, ref)
    if e_max >= 1e-4 or e_norm >= 1e-4:
        exact = f64()
        m_max, t_max = rel_err(got.double(), exact), rel_err(ref.double(), exact)
        m_nrm, t_nrm = norm_rel_err(got.double(), exact), norm_rel_err(ref.double(), exact)
        msg = (f"{what}: CUDA vs FP32 oracle max {e_max:.2e} norm {e_norm:.2e}; vs float64: CUDA max {m_max:.2e} "
               f"norm {m_nrm:.2e}, FP32 oracle max {t_max:.2e} norm {t_nrm:.2e}")
        print("[headline parity] FP32 noise floor of the reference arithmetic: " + msg)
        assert e_max < 1e-4 or m_max <= 2.0 * t_max + 2e-5, msg
        assert e_norm < 1e-4 or m_nrm <= 2.0 * t_nrm + 1e-5, msg
    return e_max, e_norm


@pytest.mark.parametrize("where", list(STARTS))
@pytest.mark.parametrize("t", [0.33, 0.75, 1.0])
def test_eval_outputs_vs_oracle(scene, where, t):
    from oracle import nvfi_oracle as O
    cfg, nv, sd, o, d, make_sc = scene
    s0 = STARTS[where]
    oo, dd = o[s0:s0 + CHUNK].contiguous(), d[s0:s0 + CHUNK].contiguous()
    f = nv.nvfi
    f.eval()
    with torch.no_grad():
        got = f.render_rays(t, oo.cuda(), dd.cuda(), white_bg=True, ray_chunk=CHUNK)
        ref = O.render_chunk(make_sc(), t, oo, dd, white_bg=True, training=False, return_aux=True)
    rep, cache = {}, {}

    def f64(k):
        if "r" not in cache:
            cache["r"] = _oracle_f64(cfg, sd, t, oo, dd)
        return cache["r"][k]
    for k, name in enumerate(("rgb", "depth", "acc", "weights")):
        rep[name] = _gate(got[k].cpu(), ref[k], f"{where} t={t} {name}", lambda k=k: f64(k))
    # integer quantity: the number of in-box samples (exact)
    from nvfi_b200 import engine
    out = engine.render_forward(f.binding, oo.cuda(), dd.cuda(), t, white_bg=True, training=False,
                                ray_chunk=CHUNK, want_stats=True)
    assert int(out.stats[0]) == int(ref[5]["valid"].sum())
    assert torch.equal(out.valid.bool().cpu(), ref[5]["valid"])
    print(f"[headline parity] {where} t={t}: " + ", ".join(f"{k} max {v[0]:.2e} norm {v[1]:.2e}" for k, v in rep.items()))


def _relu_kink_rays(sc, oo, dd, jit, eps=4e-6):
    """Rays with an appearance sample whose MLPRender_PE hidden pre-activation (models/tensorf_base.py:88-98)
    lies within `eps` of zero.  ReLU's derivative jumps there: last-bit differences of the summation order
    decide whether the unit passes its gradient, and ONE such sample moves the first-layer gradients of a
    2 048-ray chunk by ~3e-3 (measured: tools/diag_app_bisect.py found a pre-activation of -8.7e-7).  `eps`
    is the FP32 summation noise of a 110-term dot product of O(1) features (a few 1e-6).  Such rays are
    taken out of the gradient comparison on both sides; their count is reported and bounded (with 256
    hidden units and ~3 000 appearance samples per chunk about one ray in a hundred carries such a unit)."""
    from oracle import nvfi_oracle as O
    with torch.no_grad():
        r = O.render_chunk(sc, 0.33, oo, dd, white_bg=True, training=True, jitter=jit, return_aux=True)
        aux = r[5]
        m = aux["app_mask"]
        if not bool(m.any()):
            return torch.zeros(oo.shape[0], dtype=torch.bool)
        tn = O.normalize_time_coord(sc, O.keyframe_snap(sc, torch.tensor([[0.33]])))
        xyzt = torch.cat([aux["xyz_adv"][m], tn.expand(int(m.sum()), 1)], -1)
        pts = xyzt[:, :3]
        vd = dd.view(-1, 1, 3).expand(m.shape[0], m.shape[1], 3)[m]
        inp = torch.cat([O.app_feature(sc, xyzt), vd, pts, O.positional_encoding(pts, sc.pos_pe),
                         O.positional_encoding(vd, sc.view_pe)], -1)
        (w0, b0), (w1, b1), _ = sc.render_mlp
        h0 = inp @ w0.t() + b0
        h1 = torch.relu(h0) @ w1.t() + b1
        near = (h0.abs().amin(-1) < eps) | (h1.abs().amin(-1) < eps)
        ray_of = torch.nonzero(m)[:, 0]
        bad = torch.zeros(oo.shape[0], dtype=torch.bool)
        bad[ray_of[near]] = True
        return bad


@pytest.mark.parametrize("where", list(STARTS))
def test_train_gradients_vs_oracle(scene, where):
    from oracle import nvfi_oracle as O
    cfg, nv, sd, o, d, make_sc = scene
    s0 = STARTS[where]
    oo, dd = o[s0:s0 + CHUNK].contiguous(), d[s0:s0 + CHUNK].contiguous()
    gen = torch.Generator().manual_seed(11)
    jit = torch.rand(CHUNK, 1, generator=gen)
    target = torch.rand(CHUNK, 3, generator=gen)
    kinks = _relu_kink_rays(make_sc(), oo, dd, jit)
    n_kink = int(kinks.sum())
    assert n_kink <= CHUNK // 32, n_kink
    if n_kink:      # same chunk without those rays (the chunk-global inside test does not change: same camera)
        keep = ~kinks
        oo, dd, jit, target = oo[keep].contiguous(), dd[keep].contiguous(), jit[keep].contiguous(), target[keep].contiguous()
    # CUDA
    nv.requires_grad_(True)
    f = nv.nvfi
    f.train()
    nv.zero_grad(set_to_none=True)
    rgb, depth, acc, w, _ = f.render_rays(0.33, oo.cuda(), dd.cuda(), white_bg=True, ray_chunk=CHUNK, jitter=jit)
    loss = torch.nn.functional.mse_loss(rgb, target.cuda())
    loss.backward()
    # oracle autograd.  Appearance-mask membership (weight > 1e-4, models/tensorf_keyframe.py:719) is held
    # fixed to the CUDA path's: at this grid the weights carry ~1e-4 of FP32 noise on BOTH sides (see
    # test_eval_outputs_vs_oracle), so a few samples near the threshold land on different sides, and each
    # flipped sample moves the appearance gradients of a sparse chunk by ~1e-3 — a property of the
    # threshold, not of either implementation.  The flips are counted and reported.
    sc = make_sc(True)
    thr = float(sc.ray_march_weight_thres)
    with torch.no_grad():
        w_ref = O.render_chunk(make_sc(), 0.33, oo, dd, white_bg=True, training=True, jitter=jit)[3]
    mine = (w.detach().cpu() > thr)
    flips = int((mine != (w_ref > thr)).sum())
    near = ((w_ref - thr).abs() < 3e-4)
    assert not bool(((mine != (w_ref > thr)) & ~near).any())      # only samples within the noise of the threshold
    r = O.render_chunk(sc, 0.33, oo, dd, white_bg=True, training=True, jitter=jit, app_mask_override=mine)
    ref_loss = torch.nn.functional.mse_loss(r[0], target)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * max(1.0, abs(float(ref_loss)))
    assert_close(rgb.detach().cpu(), r[0].detach(), 1e-4, f"{where} train rgb")
    from tests.helpers import rel_err
    assert rel_err(w.detach().cpu(), r[3].detach()) < 5e-4     # surface samples: see test_eval_outputs_vs_oracle
    params = dict(f.named_parameters())
    errs = {}
    for name, p in oracle_param_map(sc).items():
        if p.grad is None or "a_weight_net" in name or float(p.grad.abs().max()) == 0.0:
            continue
        assert params[name].grad is not None, name
        errs[name] = norm_rel_err(params[name].grad.cpu(), p.grad)
    nv.requires_grad_(False)
    f.eval()
    if where == "top" and not errs:      # a chunk of empty rays has no gradient at all on either side
        return
    bad = {k: v for k, v in errs.items() if not v < 1e-4}
    assert not bad, bad
    print(f"[headline parity] {where} gradients: {len(errs)} tensors, worst ||d||/||ref|| = {max(errs.values()):.2e}; "
          f"appearance-mask flips {flips} of {int(mine.sum())}; rays at a ReLU kink left out: {n_kink}")
    if where == "centre":
        assert len(errs) >= 25      # 12 planes + basis_mat + 6 render MLP + 12 velocity net
