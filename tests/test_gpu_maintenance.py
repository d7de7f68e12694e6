"""GPU parity of the 'next' rows of SURVEY.md section 8(f) that sit directly on the hot-path kernels:
compute_alpha / getDenseAlpha / updateAlphaMask (models/tensorf_keyframe.py:379-405, 461-537) and
the parameter-replacing maintenance ops upsample_volume_grid / shrink (:328-376, 408-458), which
must invalidate the packed device layouts."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nvfi_oracle as O
from oracle.scene_io import scene_from_state
from tests.helpers import GOLDEN_SCENES, Golden, build_model, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def g(request):
    return Golden(request.param)


def test_compute_alpha_vs_oracle(g):
    sc = g.scene()
    model = build_model(g)
    f = model.nvfi
    f.eval()
    gen = torch.Generator().manual_seed(2)
    n = 3000
    lo, hi = sc.aabb
    xyz = lo + (hi - lo) * torch.rand(n, 3, generator=gen)
    t = torch.rand(n, 1, generator=gen) * 1.2          # includes extrapolated times
    locs = torch.cat([xyz, t], -1)
    step, _ = O.step_size_and_nsamples(sc)
    with torch.no_grad():
        ref = O.compute_alpha(sc, locs, step)
        ref_tr = O.compute_alpha(sc, locs, step, transfer=True)
    got = f.compute_alpha(locs.cuda(), f.stepSize)
    got_tr = f.compute_alpha(locs.cuda(), f.stepSize, transfer=True)
    assert rel_err(got.cpu().reshape(-1), ref.reshape(-1)) < TOL
    assert rel_err(got_tr.cpu().reshape(-1), ref_tr.reshape(-1)) < TOL


def test_update_alpha_mask_vs_oracle(g):
    """Dense alpha over 60 time steps -> max-pool -> threshold -> new aabb."""
    sc = g.scene()
    model = build_model(g)
    f = model.nvfi
    f.eval()
    gs = (20, 18, 16)
    new_aabb = f.updateAlphaMask(gs)
    vol = f.alphaMask.alpha_volume.cpu().reshape(gs[::-1])
    # oracle restatement of getDenseAlpha + updateAlphaMask on the same grid
    samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, gs[0]), torch.linspace(0, 1, gs[1]),
                                         torch.linspace(0, 1, gs[2]), indexing="ij"), -1)
    dense = sc.aabb[0] * (1 - samples) + sc.aabb[1] * samples
    step, _ = O.step_size_and_nsamples(sc)
    alpha = torch.zeros(gs)
    flat = dense.view(-1, 3)
    with torch.no_grad():
        for tt in np.linspace(0, 59, 60) / 60:
            locs = torch.cat([flat, torch.ones(flat.shape[0], 1) * tt], -1)
            alpha = torch.maximum(alpha, O.compute_alpha(sc, locs, step).view(gs))
    a = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
    a = F.max_pool3d(a, kernel_size=3, padding=1, stride=1).view(gs[::-1])
    ref = (a >= sc.alpha_mask_thres).float()
    near_thr = (a - sc.alpha_mask_thres).abs() < 1e-6
    assert not bool(((vol != ref) & ~near_thr).any())
    dxyz = dense.transpose(0, 2).contiguous()
    valid = dxyz[ref > 0.5]
    if valid.numel():
        want = torch.stack((valid.amin(0), valid.amax(0)))
        assert rel_err(new_aabb.cpu(), want) < 1e-5
    # the freshly built mask is picked up by the next eval render (sampler skip)
    o, d = g.rays()
    from nvfi_b200 import models as M
    out = M.Renderer(model, 0, 0, g.ray_chunk).render(0.2, M.Ray(o.cuda(), d.cuda(), 0, 0),
                                                      white_background=True, mode="test")
    sc2 = g.scene()
    sc2.alpha_volume = ref.view(1, 1, *ref.shape)
    with torch.no_grad():
        want_r = O.render(sc2, 0.2, o, d, ray_chunk=g.ray_chunk, white_bg=True, training=False)
    assert rel_err(out[0].reshape(-1, 3).cpu(), want_r[0]) < TOL
    assert rel_err(out[3].reshape(o.shape[0], -1).cpu(), want_r[3]) < TOL


def test_upsample_repacks_and_matches_oracle(g):
    """upsample_volume_grid replaces every plane Parameter: the packed layouts must follow."""
    model = build_model(g)
    f = model.nvfi
    f.eval()
    o, d = g.rays()
    from nvfi_b200 import models as M
    r = M.Renderer(model, 0, 0, g.ray_chunk)
    rays = M.Ray(o.cuda(), d.cuda(), 0, 0)
    before = r.render(0.2, rays, white_background=True, mode="test")[0].clone()
    new_res = [int(x) + 9 for x in g.grid]
    f.upsample_volume_grid(new_res, g.K)
    after = r.render(0.2, rays, white_background=True, mode="test")
    sd = {k: v.detach().cpu() for k, v in f.state_dict().items()}
    sc = scene_from_state(g.cfg, new_res, g.K, sd)
    with torch.no_grad():
        want = O.render(sc, 0.2, o, d, ray_chunk=g.ray_chunk, white_bg=True, training=False)
    assert f.nSamples == O.step_size_and_nsamples(sc)[1]
    assert rel_err(after[0].reshape(-1, 3).cpu(), want[0]) < TOL
    assert rel_err(after[3].reshape(o.shape[0], -1).cpu(), want[3]) < TOL
    assert float((after[0] - before.reshape(after[0].shape)).abs().max()) > 0   # it did change something


def test_shrink_matches_oracle(g):
    model = build_model(g)
    f = model.nvfi
    f.eval()
    new_aabb = f.updateAlphaMask(tuple(int(x) for x in g.grid))
    f.shrink(new_aabb)
    o, d = g.rays()
    from nvfi_b200 import models as M
    out = M.Renderer(model, 0, 0, g.ray_chunk).render(0.2, M.Ray(o.cuda(), d.cuda(), 0, 0),
                                                      white_background=True, mode="test")
    sd = {k: v.detach().cpu() for k, v in f.state_dict().items()}
    grid = [int(x) for x in f.gridSize.tolist()]
    sc = scene_from_state(g.cfg, grid, g.K, sd)
    sc.aabb = f.aabb.detach().cpu()
    sc.alpha_volume = f.alphaMask.alpha_volume.detach().cpu().float()
    if sc.vel_gate == "sur":     # the Sur gate bounds were fixed at construction (reference behaviour)
        sc.vel_bounds = f.vel.bounds.detach().cpu()
    with torch.no_grad():
        want = O.render(sc, 0.2, o, d, ray_chunk=g.ray_chunk, white_bg=True, training=False)
    assert rel_err(out[0].reshape(-1, 3).cpu(), want[0]) < TOL
    assert rel_err(out[3].reshape(o.shape[0], -1).cpu(), want[3]) < TOL
