"""nvfi_b200.seg_loss (CUDA kNN + the reference's tensor algebra) against the golden vectors of the
reference's module and against the oracle at the trainer's scale."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = {k: torch.from_numpy(v) for k, v in
     np.load(os.path.join(os.path.dirname(__file__), "golden", "segloss_small.npz")).items()}


def _close(a, b, tol=1e-5):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


def test_knn_exact():
    from nvfi_b200 import seg_loss as S
    pc = G["pc"].cuda()
    for K in (1, 2, 4, 8, 16):
        dist, idx, _ = S.knn_points(pc, pc, K)
        if K == 8:
            assert torch.equal(idx.cpu(), G["knn8_idx"])
            assert torch.equal(dist.cpu(), G["knn8_dist"])      # same (a-b)^2 summation order: bit-exact
        assert torch.equal(idx[:, :, 0].cpu(), torch.arange(pc.shape[1]).expand(pc.shape[0], -1))   # itself first
        assert bool((dist[:, :, 1:] >= dist[:, :, :-1]).all())
    with pytest.raises(RuntimeError):
        S.knn_points(pc, pc, 3)          # unsupported K fails loudly
    # ragged: fewer points than K -> -1 / inf padding
    d, i, _ = S.knn_points(pc[:, :5], pc[:, :3], 4)
    assert bool((i[:, :, 3] == -1).all()) and bool(torch.isinf(d[:, :, 3]).all())


def test_losses_match_reference_goldens():
    from nvfi_b200 import seg_loss as S
    pc, flow, mask = G["pc"].cuda(), G["flow"].cuda(), G["mask"].cuda()
    R, t = S.fit_motion_svd_batch(pc, pc + flow, mask[..., 0])
    _close(R, G["fit_R"], 2e-5); _close(t, G["fit_t"], 2e-5)
    R, t = S.fit_motion_svd_batch(pc, pc + flow)
    _close(R, G["fit_R_nomask"], 2e-5); _close(t, G["fit_t_nomask"], 2e-5)
    mk = mask.clone().requires_grad_(True)
    loss, ptf = S.dynamic_loss(pc, mk, flow)
    loss.backward()
    _close(loss, G["dynamic_loss"], 2e-5); _close(ptf, G["dynamic_pc"], 2e-5)
    _close(mk.grad, G["dynamic_grad_mask"], 2e-5)
    for name, kw in (("smooth_k4", dict(k=4, radius=0.01)), ("smooth_k16", dict(k=16, radius=0.1)),
                     ("smooth_k8_l2", dict(k=8, radius=0.05, loss_norm=2))):
        mk = mask.clone().requires_grad_(True)
        ls = S.smooth_loss(pc, mk, **kw)
        ls.backward()
        _close(ls, G[name]); _close(mk.grad, G[name + "_grad_mask"])
    _close(S.entropy_loss(mask), G["entropy"]); _close(S.rank_loss(mask), G["rank"], 2e-5)


def test_trainer_scale_against_oracle():
    """train_segm.py:126-197 shape: one cloud of ~20 000 occupied points, 8 objects, k=4, radius 0.01."""
    from nvfi_b200 import seg_loss as S
    from oracle import seg_loss_oracle as O
    g = torch.Generator().manual_seed(2)
    N = 20000
    pc = torch.rand(1, N, 3, generator=g) * 2 - 1
    mask = torch.softmax(2 * torch.randn(1, N, 8, generator=g), -1)
    dist, idx, _ = S.knn_points(pc.cuda(), pc.cuda(), 4)
    # oracle in slabs of 2 000 queries (the full N x N matrix is 1.6 GB)
    for s in range(0, N, 2000):
        d_ref, i_ref = O.knn_points(pc[:, s:s + 2000], pc, 4)
        assert torch.equal(idx[:, s:s + 2000].cpu(), i_ref)
        assert torch.equal(dist[:, s:s + 2000].cpu(), d_ref)
    flow = 0.05 * torch.randn(1, N, 3, generator=g)
    mk = mask.cuda().requires_grad_(True)
    loss, _ = S.dynamic_loss(pc.cuda(), mk, flow.cuda())     # the reference formula would need a 25 GB diag_embed here
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(mk.grad).all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        S.knn_points(pc.cuda(), pc.cuda(), 4)
    e1.record()
    torch.cuda.synchronize()
    print(f"[seg_loss] knn_points N={N} K=4: {e0.elapsed_time(e1) / 5:.3f} ms")
