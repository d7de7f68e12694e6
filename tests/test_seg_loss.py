"""CPU: the seg-loss oracle against the golden vectors made from the reference's own module
(tests/golden/make_golden_segloss.py), SURVEY.md section 8 row f4."""
import os

import numpy as np
import torch

from oracle import seg_loss_oracle as O

G = {k: torch.from_numpy(v) for k, v in
     np.load(os.path.join(os.path.dirname(__file__), "golden", "segloss_small.npz")).items()}


def _close(a, b, tol=1e-5):
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


def test_fit_and_dynamic_loss_match_reference():
    pc, flow, mask = G["pc"], G["flow"], G["mask"]
    R, t = O.fit_motion_svd_batch(pc, pc + flow, mask[..., 0])
    _close(R, G["fit_R"]); _close(t, G["fit_t"])
    R, t = O.fit_motion_svd_batch(pc, pc + flow)
    _close(R, G["fit_R_nomask"]); _close(t, G["fit_t_nomask"])
    mk = mask.clone().requires_grad_(True)
    loss, ptf = O.dynamic_loss(pc, mk, flow)
    loss.backward()
    _close(loss.detach(), G["dynamic_loss"]); _close(ptf.detach(), G["dynamic_pc"])
    _close(mk.grad, G["dynamic_grad_mask"])


def test_smooth_entropy_rank_match_reference():
    pc, mask = G["pc"], G["mask"]
    for name, kw in (("smooth_k4", dict(k=4, radius=0.01)), ("smooth_k16", dict(k=16, radius=0.1)),
                     ("smooth_k8_l2", dict(k=8, radius=0.05, loss_norm=2))):
        mk = mask.clone().requires_grad_(True)
        ls = O.smooth_loss(pc, mk, **kw)
        ls.backward()
        _close(ls.detach(), G[name]); _close(mk.grad, G[name + "_grad_mask"])
    _close(O.entropy_loss(mask), G["entropy"]); _close(O.rank_loss(mask), G["rank"])


def test_rigid_motion_is_recovered():
    """Known answer: a pure rotation + translation of the whole cloud."""
    g = torch.Generator().manual_seed(1)
    pc = torch.rand(1, 50, 3, generator=g)
    a = 0.7
    R = torch.tensor([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=torch.float32)
    t = torch.tensor([0.1, -0.2, 0.3])
    Rf, tf = O.fit_motion_svd_batch(pc, pc @ R.t() + t)
    _close(Rf[0], R, 1e-5); _close(tf[0], t, 1e-5)
