"""Host side of the plane regularisers (models/tensorf_keyframe.py:188-231): on CPU tensors the
mirror keeps the reference's torch expressions (the fused CUDA kernels need CUDA planes and refuse
anything else), and a `reg` callable without TVLoss_weight is applied as the reference applies it."""
import pytest
import torch

from tests.test_gpu_regularizers import RefTVLoss


@pytest.fixture(scope="module")
def field():
    from nvfi_b200.scenes import build_scene
    cfg, nv, _ = build_scene("bat", grid=(20, 18, 16), device="cpu")
    return nv.nvfi


def test_cpu_planes_use_the_reference_expressions(field):
    f = field
    reg = RefTVLoss(0.5)
    l1 = f.density_L1()
    ref = sum(torch.mean(torch.abs(f.density_plane_space[k])) + torch.mean(torch.abs(1 - f.density_plane_time[k]))
              for k in range(3))
    assert torch.allclose(l1, ref)
    tv = f.TV_loss_density(reg)
    ref_tv = sum(reg(f.density_plane_space[k]) * 1e-2 + reg(f.density_plane_time[k], t=True) * 1e-2 for k in range(3))
    assert torch.allclose(tv, ref_tv)
    tva = f.TV_loss_app(reg)
    assert torch.allclose(tva, sum(reg(f.app_plane_space[k]) * 1e-2 for k in range(3)))


def test_fused_path_refuses_host_tensors():
    from nvfi_b200 import regularizers as R
    with pytest.raises(RuntimeError):
        R._PlaneReg.apply([("l1", 0.0, 1.0)], torch.rand(16))
