"""NVFi: the module the drivers hold (reference: models/nvfi.py:17-84)."""
import torch
import torch.nn as nn

from .field import TensorVMKeyframeTimeKplane

_MODELS = {"TensorVMKeyframeTimeKplane": TensorVMKeyframeTimeKplane}


class NVFi(nn.Module):
    def __init__(self, config, device, aabb, res_cur, near_far):
        super().__init__()
        self.config = config.nvfi
        name = self.config.model_name
        if name not in _MODELS:
            raise NotImplementedError(f"nvfi_b200: model {name!r} (every shipped config uses "
                                      "TensorVMKeyframeTimeKplane)")
        self.nvfi = _MODELS[name](aabb, res_cur, device, near_far=near_far, cfg=config.nvfi)

    def render_ray(self, t, ray_o, ray_d, white_bg=True, ndc_ray=False):
        return self.nvfi(t, ray_o, ray_d, white_bg, ndc_ray)

    def render_ray_transfer(self, t, ray_o, ray_d, white_bg=True, ndc_ray=False):
        return self.nvfi(t, ray_o, ray_d, white_bg, ndc_ray, transfer_vel=True)

    def update_nvfi_kwargs(self, kwargs):
        """Pokes values straight into the field's __dict__ like the reference
        (models/nvfi.py:33-35)."""
        for k, v in kwargs.items():
            self.nvfi.__dict__[k] = v

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001, lr_init_velocity=0.001):
        return self.nvfi.get_optparam_groups(lr_init_spatialxyz, lr_init_network)

    def get_vel_loss(self, n_pts=32768., points=None, t=None):
        """Divergence + transport PDE loss on occupied random points (models/nvfi.py:42-84).
        ``points`` (normalised, (P,3)) and ``t`` ((P,1)) may be passed in for reproducible
        tests; otherwise they are drawn on the device like the reference."""
        from ..pde import vel_loss
        return vel_loss(self.nvfi, int(n_pts), points, t)
