"""Fused plane regularisers (csrc/regularizers.cu) against the
reference's torch expressions: TVLoss.forward (utils/tensorf_utils.py:139-158) as used by
TV_loss_density / TV_loss_app (models/tensorf_keyframe.py:205-231) and density_L1 (:188-203),
loss value and gradient, on ragged shapes and through the field methods the training loop calls
(train_nvfi.py:210-224)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class RefTVLoss(torch.nn.Module):
    """Restatement of utils/tensorf_utils.py:139-158 (the object train_nvfi.py passes as `reg`)."""

    def __init__(self, TVLoss_weight=1.0):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x, t=False):
        b, h_x, w_x = x.size(0), x.size(2), x.size(3)
        count_h = x[:, :, 1:, :].numel() // b
        count_w = x[:, :, :, 1:].numel() // b
        h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :h_x - 1, :], 2).sum() * (3 if t else 1)
        w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :w_x - 1], 2).sum()
        return self.TVLoss_weight * 2 * (h_tv / count_h + w_tv / count_w) / b


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(1, 24, 199, 199), (1, 24, 16, 199), (1, 48, 64, 80), (1, 3, 2, 2),
                                   (1, 5, 9, 33), (1, 7, 65, 31)])
@pytest.mark.parametrize("t", [False, True])
def test_tv_plane_matches_torch(shape, t):
    from nvfi_b200 import regularizers as R
    torch.manual_seed(0)
    x = (torch.rand(shape, device="cuda") - 0.3).requires_grad_(True)
    ref = RefTVLoss(0.7)(x.double(), t=t) * 1e-2
    gref, = torch.autograd.grad(ref, x)
    x2 = x.detach().clone().requires_grad_(True)
    got = R._PlaneReg.apply([("tv", int(t), 0.7 * 1e-2)], x2)
    got.backward()
    assert abs(float(got.detach()) - float(ref.detach())) <= 2e-6 * abs(float(ref.detach()))
    assert _rel(x2.grad.double(), gref.double()) < 2e-6


@pytest.mark.parametrize("n", [4, 7, 24 * 199 * 199, 1001])
@pytest.mark.parametrize("off", [0.0, 1.0])
def test_l1_plane_matches_torch(n, off):
    from nvfi_b200 import regularizers as R
    torch.manual_seed(1)
    x = (torch.rand(n, device="cuda") * 2 - 0.5)
    x[::5] = off                                  # the kink: torch.abs has gradient 0 there
    x.requires_grad_(True)
    ref = torch.mean(torch.abs(x.double() - off))
    gref, = torch.autograd.grad(ref, x)
    x2 = x.detach().clone().requires_grad_(True)
    got = R._PlaneReg.apply([("l1", off, 1.0)], x2)
    (3.0 * got).backward()
    assert abs(float(got.detach()) - float(ref.detach())) <= 2e-6 * abs(float(ref.detach()))
    assert _rel(x2.grad.double(), 3.0 * gref.double()) < 1e-6


def test_field_regularisers_match_reference_expressions():
    """The calls of train_nvfi.py:210-224 on the bench-size field (199^3, K = 16)."""
    from nvfi_b200.scenes import build_scene
    cfg, nv, _ = build_scene("bat", grid=(199, 199, 199))
    nv.requires_grad_(True)
    f = nv.nvfi
    reg = RefTVLoss(1.0)
    total = f.density_L1() * 8e-4 + f.TV_loss_density(reg) * 1.0 + f.TV_loss_app(reg) * 1.0
    nv.zero_grad(set_to_none=True)
    total.backward()
    got = {k: p.grad.clone() for k, p in f.named_parameters() if p.grad is not None}
    # reference expressions (models/tensorf_keyframe.py:188-231) in torch, float64 accumulation
    ref_total = 0
    for k in range(3):
        ds, dt, as_ = f.density_plane_space[k], f.density_plane_time[k], f.app_plane_space[k]
        ref_total = ref_total + (torch.mean(torch.abs(ds.double())) + torch.mean(torch.abs(1 - dt.double()))) * 8e-4
        ref_total = ref_total + (reg(ds.double()) + reg(dt.double(), t=True)) * 1e-2 + reg(as_.double()) * 1e-2
    nv.zero_grad(set_to_none=True)
    ref_total.backward()
    assert abs(float(total) - float(ref_total)) <= 5e-6 * abs(float(ref_total))
    names = [k for k, p in f.named_parameters() if p.grad is not None]
    assert sorted(names) == sorted(got) and len(names) == 9     # 3 density space + 3 density time + 3 app space
    for k, p in f.named_parameters():
        if p.grad is not None:
            assert _rel(got[k].double(), p.grad.double()) < 5e-6, k


def test_generic_callable_reg_still_works():
    """A `reg` that is not a TVLoss (no TVLoss_weight) is applied exactly as the reference applies it."""
    from nvfi_b200.scenes import build_scene
    cfg, nv, _ = build_scene("bat", grid=(32, 32, 32))
    f = nv.nvfi
    calls = []

    def reg(x, t=False):
        calls.append((tuple(x.shape), t))
        return x.abs().mean()

    out = f.TV_loss_density(reg)
    assert len(calls) == 6 and sum(1 for _, t in calls if t) == 3
    assert torch.is_tensor(out)
