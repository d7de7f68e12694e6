"""Host-side engine: binds a field module's parameters to the C-ABI structs, keeps the
packed (channels-last / transposed) device copies fresh, and drives the CUDA library.

PyTorch is plumbing here: device memory, streams and autograd bookkeeping.  All hot-path
arithmetic happens in ``libnvfi_b200.so``; there is no eager fallback.

Reference behaviour mirrored (host-side scalar logic only):
  * keyframe snap / isclose / time normalisation: models/tensorf_keyframe.py:646-654, 683,
    501-506 — evaluated with torch CPU FP32 ops on a 1-element tensor so the result is
    bit-identical to the reference's per-sample tensors (a render call has ONE time);
  * stratified jitter and random-background draws come from the CPU generator in the same
    order as the reference's chunk loop (models/tensorf_base.py:302-306,
    models/tensorf_keyframe.py:740, models/renderer.py:29-42).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

MAT_MODE_SPACE = ((0, 1), (0, 2), (1, 2))
MAT_MODE_TIME = ((2, 3), (1, 3), (0, 3))


MLP_MODES = {"simt": L.MLP_FP32_SIMT, "tf32x3": L.MLP_TF32X3, "tf32": L.MLP_TF32, "f16x3": L.MLP_F16X3}


_mlp_mode = os.environ.get("NVFI_MLP_MODE", "f16x3")
if _mlp_mode not in MLP_MODES:
    raise RuntimeError(f"nvfi_b200: NVFI_MLP_MODE={_mlp_mode!r} (expected one of {sorted(MLP_MODES)})")


def set_mlp_mode(mode: str) -> str:
    """Arithmetic of the velocity-MLP GEMMs: 'f16x3' (tcgen05 FP16 tensor cores, 2-way operand split,
    FP32-grade; default), 'tf32x3' (round-1 path: 3-term TF32 split, activations in tensor memory),
    'tf32' (single TF32 pass) or 'simt' (FP32 FMA verification path).  Returns the previous mode.

    The library itself has no process-wide state: the mode travels with every call in
    ``NvfiField.mlp_mode`` (include/nvfi_b200.h).  This host-side default is what ``FieldBinding.sync``
    writes there unless the binding carries its own ``mlp_mode``."""
    global _mlp_mode
    if mode not in MLP_MODES:
        raise RuntimeError(f"nvfi_b200: bad mlp mode {mode!r}")
    prev, _mlp_mode = _mlp_mode, mode
    return prev


def get_mlp_mode() -> str:
    return _mlp_mode


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _f32(x) -> float:
    """Round a python number to FP32 the way torch does when it meets a float tensor."""
    return float(torch.tensor(float(x), dtype=torch.float32))


_EPOCH = 0          # bumped by FieldBinding.invalidate(): part of every cache key
_STRICT_SYNC = os.environ.get("NVFI_STRICT_SYNC", "0") not in ("", "0")   # debug aid: re-pack on every call


class _Tracked:
    """Remembers (data_ptr, version, shape) of source tensors to know when a packed copy is stale.

    Every in-place operation on a parameter bumps its ``_version`` (optimizer steps, ``copy_`` under
    ``no_grad``, ``load_state_dict``).  Writes through ``p.data`` do NOT (``p.data.clamp_()`` leaves
    ``p._version`` unchanged): after such a write call ``FieldBinding.invalidate()`` (or
    ``field.invalidate_packed()``), or run with NVFI_STRICT_SYNC=1, which re-packs on every call."""

    def __init__(self):
        self.key = None

    def stale(self, tensors: Sequence[Optional[torch.Tensor]]) -> bool:
        key = (_EPOCH,) + tuple((t.data_ptr(), t._version, tuple(t.shape)) if t is not None else None
                                for t in tensors)
        if key != self.key or _STRICT_SYNC:
            self.key = key
            return True
        return False


class PackedLinear:
    """nn.Linear weight (out,in)[, bias] -> zero-padded W^T (k_pad, n_pad) [+ (n_pad)]."""

    def __init__(self, out_dim: int, in_dim: int, device, hidden: bool, has_bias: bool = True,
                 diff: bool = False, umma: bool = False):
        self.out_dim, self.in_dim = out_dim, in_dim
        self.k_pad = _round_up(in_dim, 32)
        self.n_pad = 128 if hidden else _round_up(out_dim, 4)
        if hidden and out_dim != 128:
            raise RuntimeError(f"nvfi_b200: hidden width {out_dim} unsupported (kernels are built for 128)")
        self.wt = torch.zeros(self.k_pad, self.n_pad, device=device, dtype=torch.float32)
        self.bias = torch.zeros(self.n_pad, device=device, dtype=torch.float32) if has_bias else None
        # W zero-padded to (128, k_pad) for the input-gradient GEMM of the backward pass
        self.w_rows = (torch.zeros(128, self.k_pad, device=device, dtype=torch.float32)
                       if (hidden and diff) else None)
        # tensor-core image (hi/lo TF32 slabs per 32-wide K block, 128B-swizzled K-major)
        self.umma_rows = (128 if hidden else _round_up(out_dim, 16)) if umma else 0
        self.umma = (torch.zeros((self.k_pad // 32) * 2 * self.umma_rows * 32, device=device,
                                 dtype=torch.float32) if umma else None)
        # image of W^T (hidden layers that are differentiated): rows = input features
        self.ummaT_rows = (128 if in_dim == 128 else _round_up(in_dim, 32)) if (umma and diff and hidden) else 0
        self.ummaT = (torch.zeros((128 // 32) * 2 * self.ummaT_rows * 32, device=device, dtype=torch.float32)
                      if self.ummaT_rows else None)
        # FP16-split images (mlp_h.cuh): rows x K padded to 64, hi slab + lo slab per 64-wide K block
        self.h_rows = self.umma_rows
        self.h_kpad = _round_up(in_dim, 64)
        self.himg = (torch.zeros((self.h_kpad // 64) * self.h_rows * 256, device=device, dtype=torch.uint8)
                     if umma else None)
        # W^T image: rows = input features (128; 32 for the 28-wide first layer), K = outputs padded to 64
        self.hT_rows = (_round_up(in_dim, 32) if umma and diff else 0)
        self.hT_kpad = _round_up(out_dim, 64)
        self.himgT = (torch.zeros((self.hT_kpad // 64) * self.hT_rows * 256, device=device, dtype=torch.uint8)
                      if self.hT_rows else None)
        self.track = _Tracked()

    def sync(self, w: torch.Tensor, b: Optional[torch.Tensor]):
        if not self.track.stale((w, b)):
            return
        lib = L.load()
        wd = w.detach().contiguous()
        bd = b.detach().contiguous() if b is not None else None
        L.check(lib.nvfi_pack_linear(wd.data_ptr(), _ptr(bd), self.wt.data_ptr(), _ptr(self.bias),
                                     self.out_dim, self.in_dim, self.k_pad, self.n_pad, _stream()),
                "pack_linear")
        if self.w_rows is not None:
            self.w_rows[:self.out_dim, :self.in_dim].copy_(wd)
        if self.umma is not None:
            L.check(lib.nvfi_pack_linear_umma(wd.data_ptr(), self.umma.data_ptr(), self.out_dim, self.in_dim,
                                              self.umma_rows, self.k_pad, _stream()), "pack_linear_umma")
        if self.ummaT is not None:
            wT = wd.t().contiguous()      # (in, out): "out_dim" = input features, K = 128 outputs
            L.check(lib.nvfi_pack_linear_umma(wT.data_ptr(), self.ummaT.data_ptr(), self.in_dim, self.out_dim,
                                              self.ummaT_rows, 128, _stream()), "pack_linear_umma(T)")
        if self.himg is not None:
            L.check(lib.nvfi_pack_linear_h(wd.data_ptr(), self.himg.data_ptr(), self.out_dim, self.in_dim,
                                           self.h_rows, self.h_kpad, 0, _stream()), "pack_linear_h")
        if self.himgT is not None:
            L.check(lib.nvfi_pack_linear_h(wd.data_ptr(), self.himgT.data_ptr(), self.out_dim, self.in_dim,
                                           self.hT_rows, self.hT_kpad, 1, _stream()), "pack_linear_h(T)")

    def fill(self, s: L.NvfiLinear):
        s.wt = self.wt.data_ptr()
        s.bias = _ptr(self.bias)
        s.w_rows = _ptr(self.w_rows)
        s.umma = _ptr(self.umma)
        s.umma_rows = self.umma_rows
        s.ummaT = _ptr(self.ummaT)
        s.ummaT_rows = self.ummaT_rows
        s.himg = _ptr(self.himg)
        s.himgT = _ptr(self.himgT)
        s.in_dim, s.out_dim, s.k_pad, s.n_pad = self.in_dim, self.out_dim, self.k_pad, self.n_pad

    def unpack_grad(self, g_wt: torch.Tensor, g_b: Optional[torch.Tensor], want_bias: bool):
        lib = L.load()
        gw = torch.empty(self.out_dim, self.in_dim, device=g_wt.device, dtype=torch.float32)
        gb = torch.empty(self.out_dim, device=g_wt.device, dtype=torch.float32) if want_bias else None
        L.check(lib.nvfi_unpack_linear(g_wt.data_ptr(), _ptr(g_b), gw.data_ptr(), _ptr(gb),
                                       self.out_dim, self.in_dim, self.k_pad, self.n_pad, _stream()),
                "unpack_linear")
        return gw, gb


class PackedPlane:
    """(1, R, H, W) parameter -> packed (H, W, R)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.track = _Tracked()
        self.shape = None

    def sync(self, p: torch.Tensor):
        if not self.track.stale((p,)):
            return
        _, R, H, W = p.shape
        if self.buf is None or self.shape != (R, H, W):
            self.buf = torch.empty(H, W, R, device=p.device, dtype=torch.float32)
            self.shape = (R, H, W)
        src = p.detach().contiguous()
        L.check(L.load().nvfi_pack_plane(src.data_ptr(), self.buf.data_ptr(), R, H, W, _stream()),
                "pack_plane")

    def unpack_grad(self, g: torch.Tensor) -> torch.Tensor:
        R, H, W = self.shape
        out = torch.empty(1, R, H, W, device=g.device, dtype=torch.float32)
        L.check(L.load().nvfi_unpack_plane(g.data_ptr(), out.data_ptr(), R, H, W, _stream()),
                "unpack_plane")
        return out


def vel_linears(vel_net) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """(weight, bias) of the 6 Linear layers of VelBasis.weight_net / a_weight_net
    (models/velocity_field.py:58-67: index 1, then [3..7][0])."""
    seq = vel_net
    out = [(seq[1].weight, seq[1].bias)]
    for i in range(3, 8):
        out.append((seq[i][0].weight, seq[i][0].bias))
    return out


class FieldBinding:
    """Device-side view of one field module (``nvfi_b200.models.TensorVMKeyframeTimeKplane``
    or anything exposing the same attributes)."""

    def __init__(self, field):
        self.field = field
        self.s = L.NvfiField()
        self.planes = {k: [PackedPlane() for _ in range(3)]
                       for k in ("density_plane_space", "density_plane_time",
                                 "app_plane_space", "app_plane_time")}
        self.basis: Optional[PackedLinear] = None
        self.render: Optional[List[PackedLinear]] = None
        self.vel: Optional[List[PackedLinear]] = None
        self.acc: Optional[List[PackedLinear]] = None
        self.mask: Optional[List[PackedLinear]] = None
        self.mask_key = None
        self.alpha_track = _Tracked()
        self.alpha_u8: Optional[torch.Tensor] = None
        self.mlp_mode: Optional[str] = None      # None = the host-side default (set_mlp_mode)
        self.scalar_track = _Tracked()
        self.scalars = None

    # -- helpers --------------------------------------------------------------------------
    @staticmethod
    def invalidate():
        """Forget every packed copy (planes, MLP weights and their tensor-core images, the alpha volume):
        the next call re-packs from the parameters.  Needed only after writes that bypass autograd's
        version counter (``p.data.<op>_()``); the model calls it from ``upsample_volume_grid``,
        ``shrink``, ``updateAlphaMask`` and ``load_state_dict``."""
        global _EPOCH
        _EPOCH += 1

    def _device(self):
        return self.field.density_plane_space[0].device

    def sync(self) -> L.NvfiField:
        """Refresh packed buffers whose sources changed and refill the C struct."""
        f, s = self.field, self.s
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("nvfi_b200: the field must live on a CUDA device (no CPU fallback)")
        # box / grid scalars live in device tensors in the reference's module; they change only in
        # update_stepSize / shrink, so the device->host reads (3 stream syncs) are cached on the tensors'
        # identity instead of being paid by every render call
        step_t = f.stepSize if torch.is_tensor(f.stepSize) else None
        grid_t = f.gridSize if torch.is_tensor(f.gridSize) else None
        if self.scalar_track.stale((f.aabb, f.invaabbSize, step_t, grid_t)) or self.scalars is None:
            aabb = f.aabb.detach().float().cpu()
            inv = f.invaabbSize.detach().float().cpu()
            grid = [int(g) for g in (f.gridSize.tolist() if grid_t is not None else f.gridSize)]
            self.scalars = ([float(aabb[0, a]) for a in range(3)], [float(aabb[1, a]) for a in range(3)],
                            [float(inv[a]) for a in range(3)], grid,
                            float(step_t.detach().float().cpu()) if step_t is not None else None)
        lo_, hi_, inv_, grid, step_cached = self.scalars
        if grid_t is None:
            grid = [int(g) for g in f.gridSize]
        K = int(f.num_keyframes)
        for a in range(3):
            s.aabb_min[a] = lo_[a]
            s.aabb_max[a] = hi_[a]
            s.inv_aabb[a] = inv_[a]
            s.grid[a] = grid[a]
        s.num_keyframes = K
        s.tmax = _f32(f.tmax)
        s.time_scale = _f32(f.tmax / (K - 1) if K > 1 else 1)
        s.dt_max = _f32(0.5 * f.tmax / (K - 1) if K > 1 else 1)
        s.near, s.far = _f32(f.near_far[0]), _f32(f.near_far[1])
        s.step_size = step_cached if step_t is not None else _f32(f.stepSize)
        s.n_samples = int(f.nSamples)
        s.density_shift = _f32(f.density_shift)
        s.distance_scale = _f32(f.distance_scale)
        s.weight_thres = _f32(f.rayMarch_weight_thres)
        s.fea2dense_act = {"softplus": L.ACT_SOFTPLUS, "relu": L.ACT_RELU, "relu_abs": L.ACT_RELU_ABS}[f.fea2denseAct]
        if f.densityMode != "Density":
            raise RuntimeError(f"nvfi_b200: densityMode {f.densityMode!r} unsupported (configs use 'Density')")
        if f.shadingMode == "MLP_PE":
            s.shading_mode = L.SHADING_MLP_PE
        elif f.shadingMode == "SH":
            s.shading_mode = L.SHADING_SH
        else:
            raise RuntimeError(f"nvfi_b200: shadingMode {f.shadingMode!r} unsupported (MLP_PE, SH)")
        s.pos_pe, s.view_pe = int(f.pos_pe), int(f.view_pe)
        s.rd = int(f.density_plane_space[0].shape[1])
        s.ra = int(f.app_plane_space[0].shape[1])
        s.app_dim = int(f.app_dim)
        for k in range(3):
            for name, dst in (("density_plane_space", s.dplane_space), ("density_plane_time", s.dplane_time),
                              ("app_plane_space", s.aplane_space), ("app_plane_time", s.aplane_time)):
                pp = self.planes[name][k]
                pp.sync(getattr(f, name)[k])
                dst[k] = pp.buf.data_ptr()
        # basis_mat
        if self.basis is None or self.basis.out_dim != s.app_dim or self.basis.in_dim != s.ra:
            self.basis = PackedLinear(s.app_dim, s.ra, dev, hidden=False, has_bias=False)
        self.basis.sync(f.basis_mat.weight, None)
        self.basis.fill(s.basis_mat)
        # render MLP
        if s.shading_mode == L.SHADING_MLP_PE:
            mlp = f.renderModule.mlp
            lins = [mlp[0], mlp[2], mlp[4]]
            if self.render is None or self.render[0].in_dim != lins[0].in_features:
                self.render = [PackedLinear(l.out_features, l.in_features, dev, hidden=(i < 2), diff=True)
                               for i, l in enumerate(lins)]
            for i, l in enumerate(lins):
                self.render[i].sync(l.weight, l.bias)
                self.render[i].fill(s.render_mlp[i])
        # velocity nets
        s.use_vel = 1 if f.use_vel else 0
        if f.use_vel:
            if self.vel is None:
                dims = [(128, 28)] + [(128, 128)] * 4 + [(6, 128)]
                self.vel = [PackedLinear(o, i, dev, hidden=(j < 5), diff=True, umma=True)
                            for j, (o, i) in enumerate(dims)]
                self.acc = [PackedLinear(o, i, dev, hidden=(j < 5), diff=True, umma=True)
                            for j, (o, i) in enumerate(dims)]
            for j, (w, b) in enumerate(vel_linears(f.vel_net.weight_net)):
                self.vel[j].sync(w, b)
                self.vel[j].fill(s.vel_net[j])
            for j, (w, b) in enumerate(vel_linears(f.vel_net.a_weight_net)):
                self.acc[j].sync(w, b)
                self.acc[j].fill(s.acc_net[j])
            lo, hi = f.vel_gate_bounds()
            s.vel_gate = L.GATE_SUR if f.vel_gate_kind() == "sur" else L.GATE_AABB
            for a in range(3):
                s.gate_lo[a], s.gate_hi[a] = lo[a], hi[a]
        # alpha mask
        am = getattr(f, "alphaMask", None)
        if am is not None:
            vol = am.alpha_volume
            if self.alpha_track.stale((vol,)):
                self.alpha_u8 = (vol.detach().reshape(vol.shape[-3:]) > 0).to(torch.uint8).contiguous()
            s.alpha_volume = self.alpha_u8.data_ptr()
            s.alpha_grid[0], s.alpha_grid[1], s.alpha_grid[2] = (int(vol.shape[-1]), int(vol.shape[-2]),
                                                                 int(vol.shape[-3]))
        else:
            s.alpha_volume = None
            self.alpha_track.key = None
        # mask field
        mf = getattr(f, "mask_field", None)
        if mf is not None:
            if getattr(mf, "point_embed", None) is not None or len(getattr(mf, "skips", [])) > 0 and any(
                    sk < len(mf.point_fc) for sk in mf.skips):
                raise RuntimeError("nvfi_b200: MaskField with point_embed / skips is unsupported")
            lins = list(mf.point_fc) + [mf.mask_fc]
            key = tuple((l.out_features, l.in_features) for l in lins)
            if self.mask is None or self.mask_key != key:
                self.mask = [PackedLinear(l.out_features, l.in_features, dev, hidden=(i < len(lins) - 1))
                             for i, l in enumerate(lins)]
                self.mask_key = key
            for i, l in enumerate(lins):
                self.mask[i].sync(l.weight, l.bias)
                self.mask[i].fill(s.mask_net[i])
            s.mask_layers = len(lins)
            s.mask_dim = int(mf.mask_dim)
        else:
            s.mask_layers = 0
            s.mask_dim = 3
        s.mlp_mode = MLP_MODES[self.mlp_mode or _mlp_mode]
        return s


# ------------------------------------------------------------------------------------------
# per-call scalar logic
# ------------------------------------------------------------------------------------------
def time_plan(field, t, transfer_vel: bool) -> Tuple[float, float, float, bool]:
    """(t, base_time, t_norm_of_eval_time, advect) with the reference's FP32 tensor semantics
    (models/tensorf_keyframe.py:646-654, 683-699, 501-506)."""
    tt = torch.ones(1, dtype=torch.float32) * (t.detach().cpu() if torch.is_tensor(t) else t)
    K = int(field.num_keyframes)
    tsf = field.tmax / (K - 1) if K > 1 else 1
    if transfer_vel:
        base = torch.zeros_like(tt)
    else:
        base = torch.round((tt / tsf).clamp(0.0, K - 1)) * tsf

    def norm_t(x):
        if K == 1 or field.tmax == 0:
            return x * 0
        return x * 2 / field.tmax - 1

    if field.use_vel:
        key = bool(torch.isclose(tt, base))
        return float(tt), float(base), float(norm_t(base)), not key
    return float(tt), float(base), float(norm_t(tt)), False


EARLY_TERMINATION = os.environ.get("NVFI_EARLY_TERMINATION", "1") not in ("", "0")


def set_early_termination(on: bool) -> bool:
    """Early ray termination of advecting renders (include/nvfi_b200.h, NvfiRenderBuffers.ray_T): samples
    behind the point where a ray's FP32 transmittance has underflowed to exactly 0 are not advected or
    gathered — they cannot change any output or gradient.  Off = every in-box sample is evaluated, as the
    reference does.  Returns the previous setting."""
    global EARLY_TERMINATION
    prev, EARLY_TERMINATION = EARLY_TERMINATION, bool(on)
    return prev


class RenderOutputs:
    __slots__ = ("rgb_map", "depth_map", "acc_map", "weights", "mask_map", "x_adv", "valid", "rgb",
                 "sigma", "chunk_inside", "counters", "stats", "args", "keep", "x_mid", "ray_T", "ray_term")


def render_forward(binding: FieldBinding, rays_o: torch.Tensor, rays_d: torch.Tensor, t, *,
                   white_bg: bool, training: bool, jitter: Optional[torch.Tensor] = None,
                   chunk_bg: Optional[torch.Tensor] = None, transfer_vel: bool = False,
                   ray_chunk: int = 2048, save_sigma: bool = False,
                   want_stats: bool = False) -> RenderOutputs:
    """One nvfi_render_forward call over all rays (all reference chunks at once)."""
    lib = L.load()
    s = binding.sync()
    dev = rays_o.device
    if dev.type != "cuda":
        raise RuntimeError("nvfi_b200: rays must be CUDA tensors (no CPU fallback)")
    rays_o = rays_o.detach().reshape(-1, 3).contiguous().float()
    rays_d = rays_d.detach().reshape(-1, 3).contiguous().float()
    n = rays_o.shape[0]
    S = int(s.n_samples)
    tt, base, tnb, advect = time_plan(binding.field, t, transfer_vel)
    n_chunks = max(1, (n + ray_chunk - 1) // ray_chunk)

    o = RenderOutputs()
    f32 = dict(device=dev, dtype=torch.float32)
    o.rgb_map = torch.empty(n, 3, **f32)
    o.depth_map = torch.empty(n, **f32)
    o.acc_map = torch.empty(n, **f32)
    o.weights = torch.empty(n, S, **f32)
    o.mask_map = torch.zeros(n, int(s.mask_dim) if s.mask_layers > 0 else 3, **f32)
    o.x_adv = torch.empty(n, S, 3, **f32)
    o.valid = torch.empty(n, S, device=dev, dtype=torch.uint8)
    o.rgb = torch.empty(n, S, 3, **f32)
    o.sigma = torch.empty(n, S, **f32) if save_sigma else None
    # RK2 midpoints for the backward pass (training renders that advect)
    o.x_mid = torch.empty(n, S, 3, **f32) if (save_sigma and advect) else None
    o.chunk_inside = torch.empty(n_chunks, device=dev, dtype=torch.uint8)
    o.counters = torch.empty(16, device=dev, dtype=torch.int32)
    o.stats = torch.empty(4, device=dev, dtype=torch.int64) if want_stats else None
    et = EARLY_TERMINATION and advect
    o.ray_T = torch.empty(n, **f32) if et else None
    o.ray_term = torch.empty(n, device=dev, dtype=torch.int32) if et else None

    a = L.NvfiRenderArgs()
    a.n_rays = n
    a.rays_o, a.rays_d = rays_o.data_ptr(), rays_d.data_ptr()
    if training:
        if jitter is None:
            raise RuntimeError("nvfi_b200: training render needs the per-ray jitter tensor")
        jitter = jitter.detach().reshape(-1).contiguous().float()
        if jitter.device != dev:
            jitter = jitter.to(dev, non_blocking=True)
        if jitter.numel() != n:
            raise RuntimeError("nvfi_b200: jitter must have one entry per ray")
        a.jitter = jitter.data_ptr()
    else:
        a.jitter = None
    a.ray_chunk = int(ray_chunk)
    if chunk_bg is not None:
        chunk_bg = chunk_bg.to(device=dev, dtype=torch.uint8).contiguous()
        if chunk_bg.numel() != n_chunks:
            raise RuntimeError("nvfi_b200: chunk_bg must have one entry per chunk")
        a.chunk_bg = chunk_bg.data_ptr()
    else:
        a.chunk_bg = None
    a.white_bg = 1 if white_bg else 0
    a.training = 1 if training else 0
    a.t, a.base_time, a.t_norm_base, a.advect = tt, base, tnb, 1 if advect else 0

    b = L.NvfiRenderBuffers()
    b.rgb_map, b.depth_map, b.acc_map = o.rgb_map.data_ptr(), o.depth_map.data_ptr(), o.acc_map.data_ptr()
    b.weights, b.mask_map = o.weights.data_ptr(), o.mask_map.data_ptr()
    b.x_adv, b.valid, b.rgb = o.x_adv.data_ptr(), o.valid.data_ptr(), o.rgb.data_ptr()
    b.sigma = _ptr(o.sigma)
    b.chunk_inside, b.counters, b.stats = o.chunk_inside.data_ptr(), o.counters.data_ptr(), _ptr(o.stats)
    b.x_mid = _ptr(o.x_mid)
    b.ray_T, b.ray_term = _ptr(o.ray_T), _ptr(o.ray_term)
    L.check(lib.nvfi_render_forward(C.byref(s), C.byref(a), C.byref(b), _stream()), "render_forward")
    o.args = (a, b)
    o.keep = (rays_o, rays_d, jitter, chunk_bg)
    return o


GRAD_FLAT_TAIL = 256
# (flat buffer, floats used) of the most recent render_backward: every gradient it returned is a view of
# this ONE buffer, so a data-parallel step can all-reduce it in place (sharding.allreduce_grads(flat=...))
LAST_GRAD_FLAT: Optional[Tuple[torch.Tensor, int]] = None


def last_grad_flat() -> Optional[Tuple[torch.Tensor, int]]:
    return LAST_GRAD_FLAT


DEBUG_KEEP: Optional[dict] = None   # tests may set this to a dict to receive backward intermediates
LAST_BWD_COUNTERS: Optional[torch.Tensor] = None   # int32[16] of the most recent render_backward


def render_backward(binding: FieldBinding, out: RenderOutputs, g_rgb, g_depth, g_acc, g_w,
                    needs: Sequence[bool]):
    """nvfi_render_backward + conversion of the packed gradients to the parameter layouts.
    Returns gradients in the order of ``autograd._diff_params``."""
    lib = L.load()
    s = binding.s          # parameters are unchanged since forward (checked by the caller)
    s.mlp_mode = MLP_MODES[binding.mlp_mode or _mlp_mode]
    a, b = out.args
    dev = out.weights.device
    n, S = out.weights.shape
    f32 = dict(device=dev, dtype=torch.float32)

    def cg(g, shape):
        if g is None:
            return None
        g = g.detach().to(**f32).expand(shape).contiguous()
        return g

    g_rgb, g_depth = cg(g_rgb, (n, 3)), cg(g_depth, (n,))
    g_acc, g_w = cg(g_acc, (n,)), cg(g_w, (n, S))
    d = L.NvfiRenderGrads()
    d.g_rgb, d.g_depth, d.g_acc, d.g_weights = _ptr(g_rgb), _ptr(g_depth), _ptr(g_acc), _ptr(g_w)
    mlp_mode = s.shading_mode == L.SHADING_MLP_PE
    advected = bool(a.advect) and bool(s.use_vel)
    # ONE zero-filled buffer for every packed gradient accumulator and ONE buffer for the gradients in the
    # parameter layouts (a memset and two batched launches instead of ~30 fills and 22 transposes)
    plane_names = ("density_plane_space", "density_plane_time", "app_plane_space", "app_plane_time")
    sizes = []          # (packed numel, out shape)
    for name in plane_names:
        for k in range(3):
            R, H, W = binding.planes[name][k].shape
            sizes.append((H * W * R, (1, R, H, W)))
    lin = [(binding.basis, False)]
    if mlp_mode:
        lin += [(binding.render[i], True) for i in range(3)]
    if s.use_vel:
        lin += [(binding.vel[j], True) for j in range(L.VEL_LAYERS)]
    for pl, has_b in lin:
        sizes.append((pl.wt.numel(), (pl.out_dim, pl.in_dim)))
        if has_b:
            sizes.append((pl.bias.numel(), (pl.out_dim,)))
    al = lambda x: (x + 31) // 32 * 32      # 128-byte alignment of every slice
    packed = torch.zeros(sum(al(n) for n, _ in sizes), **f32)
    used = sum(al(math.prod(sh)) for _, sh in sizes)
    # zero-filled (the alignment gaps travel through the all-reduce of sharding.allreduce_grads) with a
    # tail for that collective's flags and extras
    outbuf = torch.zeros(used + GRAD_FLAT_TAIL, **f32)
    pk, ov = [], []
    po = oo = 0
    for n_, sh in sizes:
        pk.append(packed[po:po + n_])
        po += al(n_)
        m_ = math.prod(sh)
        ov.append(outbuf[oo:oo + m_].view(sh))
        oo += al(m_)
    P = L.NvfiParamGrads()
    it = iter(range(len(sizes)))
    for dst, pdst in ((d.g_dplane_space, P.dplane_space), (d.g_dplane_time, P.dplane_time),
                      (d.g_aplane_space, P.aplane_space), (d.g_aplane_time, P.aplane_time)):
        for k in range(3):
            i = next(it)
            dst[k], pdst[k] = pk[i].data_ptr(), ov[i].data_ptr()
    grads_idx = list(range(12))
    i = next(it)
    d.g_basis_mat, P.basis_mat = pk[i].data_ptr(), ov[i].data_ptr()
    grads_idx.append(i)
    if mlp_mode:
        for r in range(3):
            iw, ib = next(it), next(it)
            d.g_render_w[r], d.g_render_b[r] = pk[iw].data_ptr(), pk[ib].data_ptr()
            P.render_w[r], P.render_b[r] = ov[iw].data_ptr(), ov[ib].data_ptr()
            grads_idx += [iw, ib]
    if s.use_vel:
        for j in range(L.VEL_LAYERS):
            iw, ib = next(it), next(it)
            d.g_vel_w[j], d.g_vel_b[j] = pk[iw].data_ptr(), pk[ib].data_ptr()
            if advected:
                P.vel_w[j], P.vel_b[j] = ov[iw].data_ptr(), ov[ib].data_ptr()
                grads_idx += [iw, ib]
            else:   # keyframe render: the velocity net is not on the graph (reference: grad None)
                grads_idx += [None, None]
    g_x = torch.empty(n, S, 3, **f32)
    g_sig = torch.empty(n, S, **f32)
    g_eff = torch.empty(n, 3, **f32)
    ws_bytes = int(lib.nvfi_backward_workspace_bytes())
    ws = torch.empty(ws_bytes // 4, **f32)
    d.g_x_adv, d.g_sigma, d.g_rgb_eff = g_x.data_ptr(), g_sig.data_ptr(), g_eff.data_ptr()
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws_bytes
    L.check(lib.nvfi_render_backward(C.byref(s), C.byref(a), C.byref(b), C.byref(d), _stream()),
            "render_backward")
    L.check(lib.nvfi_unpack_render_grads(C.byref(s), C.byref(d), C.byref(P), _stream()), "unpack_render_grads")
    global LAST_BWD_COUNTERS, LAST_GRAD_FLAT
    LAST_BWD_COUNTERS = out.counters
    LAST_GRAD_FLAT = (outbuf, used)
    if DEBUG_KEEP is not None:
        DEBUG_KEEP.update(g_sigma=g_sig, g_x_adv=g_x, g_rgb_eff=g_eff, fwd=out)
    grads: List[Optional[torch.Tensor]] = [None if i is None else ov[i] for i in grads_idx]
    assert len(grads) == len(needs), (len(grads), len(needs))
    return [g if need else None for g, need in zip(grads, needs)]


# ------------------------------------------------------------------------------------------
# field queries
# ------------------------------------------------------------------------------------------
def _counters(dev):
    return torch.empty(16, device=dev, dtype=torch.int32)


def integrate_pos(binding: FieldBinding, x: torch.Tensor, t: torch.Tensor, base: torch.Tensor,
                  group: Optional[bool] = None) -> torch.Tensor:
    """``group``: sort the points by RK2 step count before tiling (True), never (False), or decide by
    looking at the spread of the step counts (None; costs one device->host sync)."""
    s = binding.sync()
    x = x.detach().reshape(-1, 3).contiguous().float()
    n = x.shape[0]
    t = t.detach().reshape(-1).contiguous().float()
    base = base.detach().reshape(-1).contiguous().float()
    out = torch.empty_like(x)
    if n == 0:
        return out
    # A 128-point tile runs as many RK2 steps as its slowest point (models/tensorf_keyframe.py:592-609:
    # the loop ends when every offset is 0).  With per-point times (get_vel_loss draws t ~ U(0,1):
    # 1 step inside [0, tmax], up to 10 beyond) almost every random tile contains a 10-step point, so
    # the points are grouped by step count first; every point's arithmetic is unchanged.
    order = None
    if n >= 4 * 128 and group is not False:
        steps = torch.ceil((t - base).abs() / float(s.dt_max)).to(torch.int32)
        if group or int(steps.max()) > int(steps.min()) + 1:
            order = torch.argsort(steps)
            x, t, base = x[order].contiguous(), t[order].contiguous(), base[order].contiguous()
    cnt = _counters(x.device)
    res = torch.empty_like(x) if order is not None else out
    L.check(L.load().nvfi_integrate_pos(C.byref(s), x.data_ptr(), t.data_ptr(), base.data_ptr(), n,
                                        res.data_ptr(), cnt.data_ptr(), _stream()), "integrate_pos")
    if order is not None:
        out[order] = res
    return out


def density_feature(binding: FieldBinding, xyzt: torch.Tensor) -> torch.Tensor:
    s = binding.sync()
    xyzt = xyzt.detach().reshape(-1, 4).contiguous().float()
    n = xyzt.shape[0]
    out = torch.empty(n, 1, device=xyzt.device, dtype=torch.float32)
    if n == 0:
        return out
    L.check(L.load().nvfi_density_feature(C.byref(s), xyzt.data_ptr(), n, out.data_ptr(), _stream()),
            "density_feature")
    return out


def density_sigma(binding: FieldBinding, xyzt: torch.Tensor) -> torch.Tensor:
    s = binding.sync()
    xyzt = xyzt.detach().reshape(-1, 4).contiguous().float()
    n = xyzt.shape[0]
    out = torch.empty(n, device=xyzt.device, dtype=torch.float32)
    if n == 0:
        return out
    L.check(L.load().nvfi_density_sigma(C.byref(s), xyzt.data_ptr(), n, out.data_ptr(), _stream()),
            "density_sigma")
    return out


def feature2density(binding: FieldBinding, feat: torch.Tensor) -> torch.Tensor:
    s = binding.sync()
    f = feat.detach().contiguous().float()
    out = torch.empty_like(f)
    if f.numel() == 0:
        return out
    L.check(L.load().nvfi_feature2density(C.byref(s), f.data_ptr(), f.numel(), out.data_ptr(), _stream()),
            "feature2density")
    return out


def app_feature(binding: FieldBinding, xyzt: torch.Tensor) -> torch.Tensor:
    s = binding.sync()
    xyzt = xyzt.detach().reshape(-1, 4).contiguous().float()
    n = xyzt.shape[0]
    out = torch.empty(n, int(s.app_dim), device=xyzt.device, dtype=torch.float32)
    if n == 0:
        return out
    cnt = _counters(xyzt.device)
    L.check(L.load().nvfi_app_feature(C.byref(s), xyzt.data_ptr(), n, out.data_ptr(), cnt.data_ptr(),
                                      _stream()), "app_feature")
    return out


def velocity(binding: FieldBinding, xyzt: torch.Tensor, full: bool) -> torch.Tensor:
    s = binding.sync()
    xyzt = xyzt.detach().reshape(-1, 4).contiguous().float()
    n = xyzt.shape[0]
    out = torch.empty(n, 6 if full else 3, device=xyzt.device, dtype=torch.float32)
    if n == 0:
        return out
    cnt = _counters(xyzt.device)
    L.check(L.load().nvfi_velocity(C.byref(s), xyzt.data_ptr(), n, 1 if full else 0, out.data_ptr(),
                                   cnt.data_ptr(), _stream()), "velocity")
    return out


def raygen(pose: torch.Tensor, H: int, W: int, focal: float,
           pixel_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Camera.get_ray_bundle for selected flat pixel ids (or all pixels), on the device."""
    pose = pose.detach().contiguous().float()
    dev = pose.device
    if pixel_ids is not None:
        pixel_ids = pixel_ids.to(device=dev, dtype=torch.int64).contiguous()
        n = pixel_ids.numel()
    else:
        n = H * W
    ro = torch.empty(n, 3, device=dev, dtype=torch.float32)
    rd = torch.empty(n, 3, device=dev, dtype=torch.float32)
    L.check(L.load().nvfi_raygen(pose.data_ptr(), H, W, _f32(focal), _ptr(pixel_ids), n, ro.data_ptr(),
                                 rd.data_ptr(), _stream()), "raygen")
    return ro, rd
