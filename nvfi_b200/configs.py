"""Scene configurations for the five BASELINE.json configs.

The reference keeps these as YAML files consumed through its ``CfgNode``
(utils/cfgnode.py:36-141, config/InDoorObj/bat.yaml:1-155,
config/InDoorSeg/chessboard.yaml).  The config system itself is out of scope
("reused as-is"): anything that offers attribute access to the same
key names works with this package — the reference's own ``CfgNode(yaml)`` included.
This module only provides (a) ``AttrDict``, a minimal attribute-dict so tests and the
bench do not need the reference checkout, and (b) the hot-path-relevant values of the
shipped configs as plain dicts.
"""
from __future__ import annotations

import copy
from typing import Any, Dict


class AttrDict(dict):
    """dict with recursive attribute access (``cfg.nvfi.tmax``) and ``in`` support."""

    def __init__(self, d: Dict[str, Any] | None = None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


_NVFI_COMMON = dict(
    state_res=64,
    model_name="TensorVMKeyframeTimeKplane",
    N_voxel_init=262144,
    N_voxel_final=8000000,
    upsamp_list=[2000, 4000, 6000, 8000, 10000],
    update_AlphaMask_list=[],
    density_n_comp=[24, 24, 24],
    appearance_n_comp=[48, 48, 48],
    app_dim=32,
    densityMode="Density",
    shadingMode="MLP_PE",
    alphaMask_thres=0.0001,
    rayMarch_weight_thres=0.0001,
    pos_pe=6,
    view_pe=6,
    fea_pe=6,
    featureC=128,
    step_ratio=0.5,
    fea2denseAct="softplus",
    max_n_samples=1024,
    tmax=0.75,
    dt=0.02,
    use_vel=True,
)

_EXPERIMENT_COMMON = dict(
    randomseed=233, device="cuda", lr_grid=0.02, lr_vel=1.0e-3, lr_net=1.0e-3,
    lr_decay_iters=-1, lr_decay_target_ratio=0.1, lr_upsample_reset=1, train_iters=30000,
    L1_weight_inital=8.0e-4, L1_weight_reset=4.0e-4, TV_weight_density=1.0, TV_weight_app=1.0,
    vel_reg_weight=1, vel_reg_n_pts=262144,
)

_SEGMENTATION = dict(n_object=8, n_iters=1000, smooth_iter=500, lrate=0.005, lrate_decay=1.0,
                     lrate_decay_step=1000, save_freq=100, loss_smooth_w=0.1, alpha_scale=10,
                     n_sample_res=64, min_t=0.5)


def _indoor_obj(name: str, train_iters: int = 30000) -> Dict[str, Any]:
    """config/InDoorObj/{bat,fallingball,fan}.yaml differ only in name/basedir/train_iters
    (SURVEY.md Appendix D)."""
    return dict(
        experiment=dict(_EXPERIMENT_COMMON, train_iters=train_iters),
        dataset=dict(type="blender", basedir=f"datasets/InDoorObj/data/{name}", half_res=True,
                     test_skip=1, near=1.0, far=8.0, white_background=True),
        renderer=dict(n_rays=2048, batch_size=131072, test_batch_size=640000, distance_scale=25,
                      tensorf_sample=True, ndc=False),
        nvfi=dict(_NVFI_COMMON, bbox_x=[-2, 2], bbox_y=[-2, 2], bbox_z=[-2, 2],
                  density_shift=-10, distance_scale=25, num_keyframes=16, num_keyframes_end=16),
        segmentation=dict(_SEGMENTATION),
    )


def _chessboard() -> Dict[str, Any]:
    """config/InDoorSeg/chessboard.yaml."""
    return dict(
        experiment=dict(_EXPERIMENT_COMMON, vel_reg_n_pts=131072),
        dataset=dict(type="blender", basedir="datasets/InDoorSeg/data/chessboard", half_res=False,
                     test_skip=1, near=0.8, far=8.1, white_background=False),
        renderer=dict(n_rays=2048, batch_size=131072, test_batch_size=640000, distance_scale=25,
                      tensorf_sample=True, ndc=False),
        nvfi=dict(_NVFI_COMMON, bbox_x=[-3.03, 3.03], bbox_y=[-3.03, 3.03], bbox_z=[-0.03, 6.03],
                  sur_x=[-2.5, 2.5], sur_y=[-2.5, 2.5], sur_z=[0.02, 5.95],
                  density_shift=-5, distance_scale=10, num_keyframes=4, num_keyframes_end=4),
        segmentation=dict(_SEGMENTATION),
    )


_CONFIGS = {
    "bat": lambda: _indoor_obj("bat"),
    "fallingball": lambda: _indoor_obj("fallingball"),
    "fan": lambda: _indoor_obj("fan", train_iters=50000),
    "chessboard": _chessboard,
}


def get_config(name: str, **nvfi_overrides) -> AttrDict:
    """Return a fresh config; ``nvfi_overrides`` patch ``cfg.nvfi`` (e.g. max_n_samples=192)."""
    cfg = AttrDict(_CONFIGS[name]())
    for k, v in nvfi_overrides.items():
        cfg.nvfi[k] = v
    return cfg


def config_names():
    return list(_CONFIGS)
