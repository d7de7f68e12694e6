"""ctypes binding of ``libnvfi_b200.so`` (the C ABI declared in include/nvfi_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails the
caller gets a RuntimeError.  The structures below mirror the C structs field by field.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# NVFI_LIB_PATH: development only (A/B timing of library variants, tools/build_variant.sh)
LIB_PATH = os.environ.get("NVFI_LIB_PATH") or os.path.join(_HERE, "libnvfi_b200.so")

ABI_VERSION = 12
VEL_LAYERS = 6
MAX_MASK_LAYERS = 8

ACT_SOFTPLUS, ACT_RELU, ACT_RELU_ABS = 0, 1, 2
MLP_DEFAULT, MLP_FP32_SIMT, MLP_TF32X3, MLP_TF32, MLP_F16X3 = 0, 1, 2, 3, 4
SHADING_MLP_PE, SHADING_SH = 0, 1
GATE_AABB, GATE_SUR = 0, 1

c_float_p = C.POINTER(C.c_float)
F3 = C.c_float * 3
I3 = C.c_int32 * 3
P3 = C.c_void_p * 3


class NvfiLinear(C.Structure):
    _fields_ = [("wt", C.c_void_p), ("bias", C.c_void_p), ("w_rows", C.c_void_p), ("umma", C.c_void_p),
                ("ummaT", C.c_void_p), ("himg", C.c_void_p), ("himgT", C.c_void_p), ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("k_pad", C.c_int32), ("n_pad", C.c_int32),
                ("umma_rows", C.c_int32), ("ummaT_rows", C.c_int32)]


class NvfiField(C.Structure):
    _fields_ = [
        ("aabb_min", F3), ("aabb_max", F3), ("inv_aabb", F3), ("grid", I3),
        ("num_keyframes", C.c_int32), ("tmax", C.c_float), ("time_scale", C.c_float),
        ("dt_max", C.c_float), ("near", C.c_float), ("far", C.c_float),
        ("step_size", C.c_float), ("n_samples", C.c_int32),
        ("density_shift", C.c_float), ("distance_scale", C.c_float), ("weight_thres", C.c_float),
        ("fea2dense_act", C.c_int32), ("shading_mode", C.c_int32), ("pos_pe", C.c_int32),
        ("view_pe", C.c_int32), ("rd", C.c_int32), ("ra", C.c_int32), ("app_dim", C.c_int32),
        ("dplane_space", P3), ("dplane_time", P3), ("aplane_space", P3), ("aplane_time", P3),
        ("basis_mat", NvfiLinear), ("render_mlp", NvfiLinear * 3),
        ("use_vel", C.c_int32),
        ("vel_net", NvfiLinear * VEL_LAYERS), ("acc_net", NvfiLinear * VEL_LAYERS),
        ("vel_gate", C.c_int32), ("gate_lo", F3), ("gate_hi", F3),
        ("alpha_volume", C.c_void_p), ("alpha_grid", I3),
        ("mask_layers", C.c_int32), ("mask_dim", C.c_int32),
        ("mask_net", NvfiLinear * MAX_MASK_LAYERS),
        ("mlp_mode", C.c_int32),
    ]


class NvfiRenderArgs(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int64), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p),
        ("jitter", C.c_void_p), ("ray_chunk", C.c_int32), ("chunk_bg", C.c_void_p),
        ("white_bg", C.c_int32), ("training", C.c_int32), ("t", C.c_float),
        ("base_time", C.c_float), ("t_norm_base", C.c_float), ("advect", C.c_int32),
    ]


class NvfiRenderBuffers(C.Structure):
    _fields_ = [
        ("rgb_map", C.c_void_p), ("depth_map", C.c_void_p), ("acc_map", C.c_void_p),
        ("weights", C.c_void_p), ("mask_map", C.c_void_p), ("x_adv", C.c_void_p),
        ("valid", C.c_void_p), ("rgb", C.c_void_p), ("sigma", C.c_void_p),
        ("chunk_inside", C.c_void_p), ("counters", C.c_void_p), ("stats", C.c_void_p),
        ("x_mid", C.c_void_p), ("ray_T", C.c_void_p), ("ray_term", C.c_void_p),
    ]


class NvfiRenderGrads(C.Structure):
    _fields_ = [
        ("g_rgb", C.c_void_p), ("g_depth", C.c_void_p), ("g_acc", C.c_void_p),
        ("g_weights", C.c_void_p),
        ("g_dplane_space", P3), ("g_dplane_time", P3), ("g_aplane_space", P3),
        ("g_aplane_time", P3), ("g_basis_mat", C.c_void_p),
        ("g_render_w", P3), ("g_render_b", P3),
        ("g_vel_w", C.c_void_p * VEL_LAYERS), ("g_vel_b", C.c_void_p * VEL_LAYERS),
        ("g_x_adv", C.c_void_p), ("g_sigma", C.c_void_p), ("g_rgb_eff", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class NvfiParamGrads(C.Structure):
    _fields_ = [
        ("dplane_space", P3), ("dplane_time", P3), ("aplane_space", P3), ("aplane_time", P3),
        ("basis_mat", C.c_void_p), ("render_w", P3), ("render_b", P3),
        ("vel_w", C.c_void_p * VEL_LAYERS), ("vel_b", C.c_void_p * VEL_LAYERS),
    ]


class NvfiPdeGrads(C.Structure):
    _fields_ = [
        ("g_vel_w", C.c_void_p * VEL_LAYERS), ("g_vel_b", C.c_void_p * VEL_LAYERS),
        ("g_acc_w", C.c_void_p * VEL_LAYERS), ("g_acc_b", C.c_void_p * VEL_LAYERS),
        ("g_acc_pts", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class NvfiPdeParamGrads(C.Structure):
    _fields_ = [
        ("vel_w", C.c_void_p * VEL_LAYERS), ("vel_b", C.c_void_p * VEL_LAYERS),
        ("acc_w", C.c_void_p * VEL_LAYERS), ("acc_b", C.c_void_p * VEL_LAYERS),
    ]


class NvfiProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_double), ("launches", C.c_int64)]


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol include/nvfi_b200.h declares
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "nvfi_abi_version": (_i, []),
    "nvfi_backward_workspace_bytes": (_i64, []),
    "nvfi_launch_count": (_i64, []),
    "nvfi_profile_enable": (_i, [_i]),
    "nvfi_profile_read": (_i, [C.POINTER(NvfiProfileEntry), _i, _i]),
    "nvfi_pack_linear_umma": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "nvfi_pack_linear_h": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nvfi_pack_plane": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "nvfi_unpack_plane": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "nvfi_pack_linear": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nvfi_unpack_linear": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nvfi_raygen": (_i, [_vp, _i, _i, _f, _vp, _i64, _vp, _vp, _vp]),
    "nvfi_render_forward": (_i, [C.POINTER(NvfiField), C.POINTER(NvfiRenderArgs),
                                 C.POINTER(NvfiRenderBuffers), _vp]),
    "nvfi_render_backward": (_i, [C.POINTER(NvfiField), C.POINTER(NvfiRenderArgs),
                                  C.POINTER(NvfiRenderBuffers), C.POINTER(NvfiRenderGrads), _vp]),
    "nvfi_unpack_render_grads": (_i, [C.POINTER(NvfiField), C.POINTER(NvfiRenderGrads),
                                      C.POINTER(NvfiParamGrads), _vp]),
    "nvfi_unpack_pde_grads": (_i, [C.POINTER(NvfiField), C.POINTER(NvfiPdeGrads), C.POINTER(NvfiPdeParamGrads), _vp]),
    "nvfi_render_forward_host": (_i, [C.POINTER(NvfiField), C.POINTER(NvfiRenderArgs), _vp, _vp,
                                      _vp, _vp, _vp, _vp, C.POINTER(NvfiRenderBuffers), _vp, _vp,
                                      _vp, _vp]),
    "nvfi_integrate_pos": (_i, [C.POINTER(NvfiField), _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "nvfi_density_feature": (_i, [C.POINTER(NvfiField), _vp, _i64, _vp, _vp]),
    "nvfi_density_sigma": (_i, [C.POINTER(NvfiField), _vp, _i64, _vp, _vp]),
    "nvfi_feature2density": (_i, [C.POINTER(NvfiField), _vp, _i64, _vp, _vp]),
    "nvfi_app_feature": (_i, [C.POINTER(NvfiField), _vp, _i64, _vp, _vp, _vp]),
    "nvfi_velocity": (_i, [C.POINTER(NvfiField), _vp, _i64, _i, _vp, _vp, _vp]),
    "nvfi_debug_timeline": (_i, [_vp, _i]),
    "nvfi_knn_points": (_i, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "nvfi_tv_loss": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "nvfi_l1_loss": (_i, [_vp, _i64, _f, _f, _vp, _vp, _vp]),
    "nvfi_pde_loss": (_i, [C.POINTER(NvfiField), _vp, _vp, _i64, _vp, C.POINTER(NvfiPdeGrads), _i, _vp, _vp]),
}


def load(path: Optional[str] = None) -> C.CDLL:
    """Load the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"nvfi_b200: CUDA library {p} not found. Build it with `python -m nvfi_b200.build` "
            "(needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    v = lib.nvfi_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"nvfi_b200: ABI mismatch (library {v}, binding {ABI_VERSION}); rebuild")
    if path is None:
        _lib = lib
    return lib


def load_debug() -> C.CDLL:
    """Development probes (include/nvfi_b200_debug.h, csrc/debug/): a separate library, used by tests only."""
    p = os.path.join(_HERE, "libnvfi_b200_debug.so")
    if not os.path.exists(p):
        raise RuntimeError(f"nvfi_b200: {p} not found. Build it with `python -m nvfi_b200.build`.")
    lib = C.CDLL(p)
    lib.nvfi_debug_mma_mn.restype = _i
    lib.nvfi_debug_mma_mn.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _vp]
    return lib


def launch_count() -> int:
    return int(load().nvfi_launch_count())


def profile_enable(on: bool) -> None:
    check(load().nvfi_profile_enable(1 if on else 0), "profile_enable")


def profile_read(reset: bool = True) -> dict:
    """{kernel name: (total ms, launches)} of the launches recorded since the last reset."""
    buf = (NvfiProfileEntry * 64)()
    n = load().nvfi_profile_read(buf, 64, 1 if reset else 0)
    if n < 0:
        raise RuntimeError(f"nvfi_b200 profile_read: error {n}")
    return {buf[i].name.decode(): (float(buf[i].ms), int(buf[i].launches)) for i in range(n)}


_ERRORS = {-1: "NVFI_EINVAL (bad argument)", -2: "NVFI_EUNSUPPORTED (configuration outside the kernels)"}


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"nvfi_b200 {what}: {_ERRORS.get(rc, rc)}")
    raise RuntimeError(f"nvfi_b200 {what}: CUDA error {rc}")
