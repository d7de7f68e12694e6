"""CPU oracle for the NVFi render + velocity-advection hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``nvfi_b200``) never imports this module and
fails loudly when its CUDA library is missing.

It is an independent functional restatement (torch, CPU, FP32) of the algorithm
in the reference repository, written from the behaviour spec in SURVEY.md
Appendix A.  Every function cites the reference ``file:line`` it follows.

Parity pinning: the reference has NO tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
generated in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them.  The bilinear gather's arithmetic
lives in a third-party dependency (``torch.nn.functional.grid_sample``, README pins
pytorch==1.12.1; torch 2.11 here): ``bilerp_manual`` restates its published
algorithm (bilinear, zero padding, align_corners=True) and is cross-checked against
the library call, which the fast path uses exactly as the reference's call sites do
(models/tensorf_keyframe.py:259-264, 300-305).

All tensors are torch CPU float32 unless noted; the module is autograd-friendly so
gradient parity is obtained with ``torch.autograd.grad`` on these functions.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

MAT_MODE_SPACE = ((0, 1), (0, 2), (1, 2))  # models/tensorf_keyframe.py:39
MAT_MODE_TIME = ((2, 3), (1, 3), (0, 3))   # models/tensorf_keyframe.py:40


# --------------------------------------------------------------------------------------
# Scene container (plain tensors + scalars; no nn.Module on purpose)
# --------------------------------------------------------------------------------------
@dataclass
class Scene:
    """All state the hot path reads.  Plane layout is the reference's NCHW."""

    aabb: torch.Tensor                      # (2,3)
    grid_size: Sequence[int]                # [Gx,Gy,Gz]
    num_keyframes: int
    tmax: float
    near: float
    far: float
    step_ratio: float = 0.5
    max_n_samples: int = 1024
    density_shift: float = -10.0
    distance_scale: float = 25.0
    alpha_mask_thres: float = 1e-4
    ray_march_weight_thres: float = 1e-4
    fea2dense_act: str = "softplus"
    shading_mode: str = "MLP_PE"            # or "SH"
    pos_pe: int = 6
    view_pe: int = 6
    # factor planes, lists of 3 tensors each
    density_plane_space: List[torch.Tensor] = field(default_factory=list)   # (1,Rd,G[m1],G[m0])
    density_plane_time: List[torch.Tensor] = field(default_factory=list)    # (1,Rd,K,G[n0])
    app_plane_space: List[torch.Tensor] = field(default_factory=list)
    app_plane_time: List[torch.Tensor] = field(default_factory=list)
    basis_mat: Optional[torch.Tensor] = None                                 # (app_dim, Ra)
    render_mlp: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None     # 3 x (W,b)
    vel_net: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None        # 6 x (W,b), SiLU
    acc_net: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None        # 6 x (W,b), ReLU
    vel_gate: str = "aabb"                  # "aabb" (VelocityAABB) | "sur" (VelocityAABBSur)
    vel_eps: float = 0.03
    vel_bounds: Optional[torch.Tensor] = None   # (2,3) normalised bounds for "sur"
    alpha_volume: Optional[torch.Tensor] = None  # (1,1,Gz,Gy,Gx) binary float
    mask_field: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None  # hidden layers + head
    use_vel: bool = True

    def parameters(self) -> List[torch.Tensor]:
        ps = list(self.density_plane_space) + list(self.density_plane_time)
        ps += list(self.app_plane_space) + list(self.app_plane_time)
        if self.basis_mat is not None:
            ps.append(self.basis_mat)
        for net in (self.render_mlp, self.vel_net):
            if net is not None:
                for w, b in net:
                    ps += [w, b]
        return ps


# --------------------------------------------------------------------------------------
# a8 / a9: step size, normalisation
# --------------------------------------------------------------------------------------
def step_size_and_nsamples(sc: Scene) -> Tuple[torch.Tensor, int]:
    """models/tensorf_base.py:214-227 (update_stepSize)."""
    aabb_size = sc.aabb[1] - sc.aabb[0]
    g = torch.tensor(list(sc.grid_size), dtype=torch.long)
    units = aabb_size / (g - 1)
    step = torch.mean(units) * sc.step_ratio
    diag = torch.sqrt(torch.sum(torch.square(aabb_size)))
    n = min(sc.max_n_samples, int((diag / step).item()) + 1)
    return step, n


def normalize_coord(sc: Scene, p: torch.Tensor) -> torch.Tensor:
    """models/tensorf_base.py:241-242."""
    inv = 2.0 / (sc.aabb[1] - sc.aabb[0])
    return (p - sc.aabb[0]) * inv - 1


def normalize_time_coord(sc: Scene, t: torch.Tensor) -> torch.Tensor:
    """models/tensorf_keyframe.py:501-506."""
    if sc.num_keyframes == 1 or sc.tmax == 0:
        return t * 0
    return t * 2 / sc.tmax - 1


def time_scale_factor(sc: Scene) -> float:
    """models/tensorf_keyframe.py:646."""
    return sc.tmax / (sc.num_keyframes - 1) if sc.num_keyframes > 1 else 1


def keyframe_snap(sc: Scene, t: torch.Tensor) -> torch.Tensor:
    """models/tensorf_keyframe.py:651-653: nearest keyframe time (round half to even)."""
    tsf = time_scale_factor(sc)
    return torch.round((t / tsf).clamp(0.0, sc.num_keyframes - 1)) * tsf


# --------------------------------------------------------------------------------------
# a1: pinhole rays
# --------------------------------------------------------------------------------------
def raygen(pose: torch.Tensor, H: int, W: int, focal: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """models/camera.py:112-133 (get_ray_bundle, non-NDC). Returns o,d of shape (H,W,3)."""
    X, Y = torch.meshgrid(torch.arange(W, dtype=pose.dtype), torch.arange(H, dtype=pose.dtype),
                          indexing="xy")
    dirs = torch.stack([(X - W * 0.5) / focal, -(Y - H * 0.5) / focal, -torch.ones_like(X)], -1)
    d = torch.sum(dirs[..., None, :] * pose[:3, :3], dim=-1)
    o = pose[:3, -1].expand(d.shape)
    return o, d


# --------------------------------------------------------------------------------------
# a7: stratified ray sampling
# --------------------------------------------------------------------------------------
def sample_ray(sc: Scene, o: torch.Tensor, d: torch.Tensor, jitter: Optional[torch.Tensor],
               n_samples: Optional[int] = None):
    """models/tensorf_base.py:290-314.

    ``jitter`` is the per-ray stratified offset u in [0,1) of shape (N,1) (training), or
    None (eval).  The reference draws it from the CPU generator (line 305); for parity it
    is an explicit input here.  Returns (pts (N,S,3), z (N,S), valid (N,S) bool).
    """
    step, n_default = step_size_and_nsamples(sc)
    S = n_samples if (n_samples is not None and n_samples > 0) else n_default
    near, far = sc.near, sc.far
    if ((sc.aabb[0] <= o) & (o <= sc.aabb[1])).any():   # chunk-global predicate (line 294)
        t_min = torch.ones_like(o[..., 0]) * near
    else:
        vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
        rate_a = (sc.aabb[1] - o) / vec
        rate_b = (sc.aabb[0] - o) / vec
        t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
    rng = torch.arange(S)[None].float()
    if jitter is not None:
        rng = rng.repeat(d.shape[-2], 1)
        rng = rng + jitter
    z = t_min[..., None] + step * rng
    pts = o[..., None, :] + d[..., None, :] * z[..., None]
    out = ((sc.aabb[0] > pts) | (pts > sc.aabb[1])).any(dim=-1)
    return pts, z, ~out


# --------------------------------------------------------------------------------------
# a12 / a13: velocity field
# --------------------------------------------------------------------------------------
def pos_encode_xt(xt: torch.Tensor, n_freq: int = 3) -> torch.Tensor:
    """models/base_network.py:42-54: [x, sin(x f), cos(x f) for f in 1,2,4] -> 28 dims."""
    out = [xt]
    for k in range(n_freq):
        f = float(2 ** k)
        out.append(torch.sin(xt * f))
        out.append(torch.cos(xt * f))
    return torch.cat(out, dim=-1)


def mlp_forward(layers, x: torch.Tensor, act) -> torch.Tensor:
    h = x
    for i, (w, b) in enumerate(layers):
        h = F.linear(h, w, b)
        if i < len(layers) - 1:
            h = act(h)
    return h


def vel_weights(sc: Scene, xt: torch.Tensor) -> torch.Tensor:
    """weight_net of models/velocity_field.py:58-63: 28->128 x5 (SiLU) -> 6."""
    return mlp_forward(sc.vel_net, pos_encode_xt(xt), F.silu)


def basis_velocity(w: torch.Tensor, xt: torch.Tensor) -> torch.Tensor:
    """models/velocity_field.py:77-98: v = sum_i w_i b_i(x), rigid-motion basis."""
    x, y, z = xt[..., 0], xt[..., 1], xt[..., 2]
    vx = w[..., 0] - w[..., 4] * z + w[..., 5] * y
    vy = w[..., 1] + w[..., 3] * z - w[..., 5] * x
    vz = w[..., 2] - w[..., 3] * y + w[..., 4] * x
    return torch.stack([vx, vy, vz], dim=-1)


def basis_acceleration(aw: torch.Tensor, xt: torch.Tensor) -> torch.Tensor:
    """models/velocity_field.py:69-75, 94-97: a = sum_i aw_i a_i(x)."""
    x, y, z = xt[..., 0], xt[..., 1], xt[..., 2]
    ax = aw[..., 0] - aw[..., 4] * x - aw[..., 5] * x
    ay = aw[..., 1] - aw[..., 3] * y - aw[..., 5] * y
    az = aw[..., 2] - aw[..., 3] * z - aw[..., 4] * z
    return torch.stack([ax, ay, az], dim=-1)


def get_vel(sc: Scene, xt: torch.Tensor) -> torch.Tensor:
    """VelBasis.get_vel, models/velocity_field.py:77-81."""
    return basis_velocity(vel_weights(sc, xt), xt)


def vel_full(sc: Scene, xt: torch.Tensor) -> torch.Tensor:
    """VelBasis.forward, models/velocity_field.py:69-75 -> (M,6) = [v, a]."""
    v = basis_velocity(vel_weights(sc, xt), xt)
    aw = mlp_forward(sc.acc_net, pos_encode_xt(xt), F.relu)
    return torch.cat([v, basis_acceleration(aw, xt)], dim=-1)


def gate_outside(sc: Scene, pts: torch.Tensor) -> torch.Tensor:
    """Out-of-bounds predicate of VelocityAABB (:31) / VelocityAABBSur (:49)."""
    if sc.vel_gate == "sur":
        return ((pts < sc.vel_bounds[0]) | (pts > sc.vel_bounds[1])).any(dim=-1)
    return ((pts < -1 + sc.vel_eps) | (pts > 1 - sc.vel_eps)).any(dim=-1)


def gated_vel(sc: Scene, xt: torch.Tensor) -> torch.Tensor:
    """models/velocity_field.py:28-33, 46-51: zero velocity outside the gate box."""
    outside = gate_outside(sc, xt[..., :3])
    vel = torch.zeros_like(xt[..., :3])
    inside = ~outside
    if inside.any():
        vel = vel.index_put((inside,), get_vel(sc, xt[inside]))
    return vel


def rk2_schedule(sc: Scene, t: float, base: float) -> List[Tuple[float, float]]:
    """FP32 (dt, t_curr) sequence of models/tensorf_keyframe.py:577-609 for a scalar time."""
    dt_max = torch.tensor(0.5 * sc.tmax / (sc.num_keyframes - 1) if sc.num_keyframes > 1 else 1.0,
                          dtype=torch.float32)
    off = torch.tensor(t, dtype=torch.float32) - torch.tensor(base, dtype=torch.float32)
    tc = torch.tensor(t, dtype=torch.float32)
    out = []
    while off.abs() > 0:
        dt = off.sign() * torch.minimum(off.abs(), dt_max)
        out.append((float(dt), float(tc)))
        off = off - dt
        tc = tc - dt
    return out


def integrate_pos(sc: Scene, pos: torch.Tensor, t: torch.Tensor, base: torch.Tensor) -> torch.Tensor:
    """models/tensorf_keyframe.py:575-611: RK2 (midpoint) backward advection to the keyframe.

    Functional (no in-place aliasing): returns the advected positions; inputs untouched.
    """
    dt_max = 0.5 * sc.tmax / (sc.num_keyframes - 1) if sc.num_keyframes > 1 else 1
    dt_max = torch.ones_like(t) * dt_max
    off = (t - base).clone()
    x = pos
    tc = t.clone()
    unfinished = (off.abs() > 0).squeeze(-1)
    while unfinished.any():
        idx = unfinished.nonzero(as_tuple=True)[0]
        o_u = off[idx]
        dt = o_u.sign() * torch.minimum(o_u.abs(), dt_max[idx])
        x_u, t_u = x[idx], tc[idx]
        v0 = gated_vel(sc, torch.cat([x_u, t_u], dim=-1))
        p_mid = x_u - 0.5 * dt * v0
        t_mid = t_u - 0.5 * dt
        x_new = x_u - dt * gated_vel(sc, torch.cat([p_mid, t_mid], dim=-1))
        if sc.vel_gate == "sur":   # revert samples that left the surround box (:603-605)
            left = gate_outside(sc, x_new)
            x_new = torch.where(left[:, None], x_u, x_new)
        x = x.index_put((idx,), x_new)
        off = off.index_put((idx,), o_u - dt)
        tc = tc.index_put((idx,), t_u - dt)
        unfinished = (off.abs() > 0).squeeze(-1)
    return x


# --------------------------------------------------------------------------------------
# a14 / a17: k-planes gather
# --------------------------------------------------------------------------------------
def bilerp_manual(plane: torch.Tensor, gx: torch.Tensor, gy: torch.Tensor) -> torch.Tensor:
    """Published algorithm of grid_sample(bilinear, zeros, align_corners=True) on one
    (1,R,H,W) plane at V points; gx indexes W, gy indexes H.  Returns (R,V).
    Restated from the ATen definition."""
    _, R, H, W = plane.shape
    fx = (gx + 1) / 2 * (W - 1)
    fy = (gy + 1) / 2 * (H - 1)
    x0 = torch.floor(fx)
    y0 = torch.floor(fy)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - fx) * (y1 - fy)
    w_ne = (fx - x0) * (y1 - fy)
    w_sw = (x1 - fx) * (fy - y0)
    w_se = (fx - x0) * (fy - y0)
    p = plane[0]
    out = torch.zeros(R, gx.shape[0], dtype=plane.dtype)
    for xi, yi, w in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        xc = xi.clamp(0, W - 1).long()
        yc = yi.clamp(0, H - 1).long()
        v = p[:, yc, xc]                       # (R,V)
        out = out + v * (w * ok.to(w.dtype))[None]
    return out


def _bilerp(plane: torch.Tensor, gx: torch.Tensor, gy: torch.Tensor, manual: bool) -> torch.Tensor:
    if manual:
        return bilerp_manual(plane, gx, gy)
    grid = torch.stack([gx, gy], dim=-1).view(1, -1, 1, 2)
    return F.grid_sample(plane, grid, align_corners=True).view(plane.shape[1], gx.shape[0])


def plane_features(space: Sequence[torch.Tensor], time: Sequence[torch.Tensor],
                   xyzt: torch.Tensor, manual: bool = False) -> torch.Tensor:
    """Hadamard product over the 3 space and 3 space-time planes, per component: (R,V).
    models/tensorf_keyframe.py:233-268 / 274-308."""
    prod = None
    for k in range(3):
        m0, m1 = MAT_MODE_SPACE[k]
        n0, n1 = MAT_MODE_TIME[k]
        bs = _bilerp(space[k], xyzt[:, m0], xyzt[:, m1], manual)
        bt = _bilerp(time[k], xyzt[:, n0], xyzt[:, n1], manual)
        prod = bs * bt if prod is None else prod * bs * bt
    return prod


def plane_features_ref_order(space, time, xyzt, manual=False):
    """Same as plane_features but with the reference's exact multiplication order
    (space factors multiplied together, time factors together, then the two; :266-272)."""
    ps, pt = 1.0, 1.0
    for k in range(3):
        m0, m1 = MAT_MODE_SPACE[k]
        n0, n1 = MAT_MODE_TIME[k]
        ps = ps * _bilerp(space[k], xyzt[:, m0], xyzt[:, m1], manual)
        pt = pt * _bilerp(time[k], xyzt[:, n0], xyzt[:, n1], manual)
    return ps * pt


def density_feature(sc: Scene, xyzt: torch.Tensor, manual: bool = False) -> torch.Tensor:
    """compute_densityfeature, densityMode == 'Density' (models/tensorf_keyframe.py:233-272)."""
    return plane_features_ref_order(sc.density_plane_space, sc.density_plane_time, xyzt, manual).sum(0)


def feature2density(sc: Scene, feat: torch.Tensor) -> torch.Tensor:
    """models/tensorf_keyframe.py:312-325."""
    if sc.fea2dense_act == "softplus":
        return F.softplus(feat + sc.density_shift)
    if sc.fea2dense_act == "relu":
        return F.relu(feat)
    return F.relu(torch.abs(feat))


def app_feature(sc: Scene, xyzt: torch.Tensor, manual: bool = False) -> torch.Tensor:
    """compute_appfeature (models/tensorf_keyframe.py:274-310): (A, app_dim)."""
    f = plane_features_ref_order(sc.app_plane_space, sc.app_plane_time, xyzt, manual)
    return F.linear(f.T, sc.basis_mat)


# --------------------------------------------------------------------------------------
# a16: alpha / transmittance / weights
# --------------------------------------------------------------------------------------
def raw2alpha(sigma: torch.Tensor, dist: torch.Tensor):
    """models/tensorf_model_utils.py:186-197."""
    alpha = 1.0 - torch.exp(-sigma * dist)
    T = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1), 1.0 - alpha + 1e-10], -1), -1)
    return alpha, alpha * T[:, :-1], T[:, -1:]


# --------------------------------------------------------------------------------------
# a18 / a19 / a23: appearance decoders, mask field
# --------------------------------------------------------------------------------------
def positional_encoding(p: torch.Tensor, freqs: int) -> torch.Tensor:
    """models/tensorf_model_utils.py:176-183: dim-major (d*freqs+f), [sin(...), cos(...)]."""
    bands = 2 ** torch.arange(freqs).float()
    q = (p[..., None] * bands).reshape(p.shape[:-1] + (freqs * p.shape[-1],))
    return torch.cat([torch.sin(q), torch.cos(q)], dim=-1)


def mlp_pe_render(sc: Scene, pts: torch.Tensor, viewdirs: torch.Tensor, feat: torch.Tensor):
    """MLPRender_PE.forward (models/tensorf_base.py:88-98)."""
    parts = [feat, viewdirs, pts]
    if sc.pos_pe > 0:
        parts.append(positional_encoding(pts, sc.pos_pe))
    if sc.view_pe > 0:
        parts.append(positional_encoding(viewdirs, sc.view_pe))
    return torch.sigmoid(mlp_forward(sc.render_mlp, torch.cat(parts, dim=-1), F.relu))


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
          -1.0925484305920792, 0.5462742152960396)


def sh_bases_deg2(d: torch.Tensor) -> torch.Tensor:
    """models/sh.py:87-116 for deg == 2 (9 bases)."""
    x, y, z = d.unbind(-1)
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    return torch.stack([torch.full_like(x, _SH_C0), -_SH_C1 * y, _SH_C1 * z, -_SH_C1 * x,
                        _SH_C2[0] * xy, _SH_C2[1] * yz, _SH_C2[2] * (2.0 * zz - xx - yy),
                        _SH_C2[3] * xz, _SH_C2[4] * (xx - yy)], dim=-1)


def sh_render(viewdirs: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """SHRender (models/tensorf_model_utils.py:292-296): feat (A,27) viewed (A,3,9)."""
    sh = sh_bases_deg2(viewdirs)[:, None]
    return torch.relu(torch.sum(sh * feat.view(-1, 3, sh.shape[-1]), dim=-1) + 0.5)


def mask_field_forward(layers, pts: torch.Tensor) -> torch.Tensor:
    """MaskField.forward as built at test_segm_render.py:75-80 (no skips hit for 4 layers,
    no point embedding): ReLU MLP + softmax(dim=1) (models/mask_field.py:68-83)."""
    h = pts
    for w, b in layers[:-1]:
        h = F.relu(F.linear(h, w, b))
    w, b = layers[-1]
    return F.softmax(F.linear(h, w, b), dim=1)


def sample_alpha(sc: Scene, xyz: torch.Tensor) -> torch.Tensor:
    """AlphaGridMask.sample_alpha (models/tensorf_model_utils.py:433-439)."""
    return F.grid_sample(sc.alpha_volume, xyz.view(1, -1, 1, 1, 3), align_corners=True).view(-1)


# --------------------------------------------------------------------------------------
# a6 / a10 / a20: one chunk of rays, end to end
# --------------------------------------------------------------------------------------
def render_chunk(sc: Scene, t: float, o: torch.Tensor, d: torch.Tensor, *, white_bg: bool,
                 training: bool, jitter: Optional[torch.Tensor] = None, random_bg: bool = False,
                 transfer_vel: bool = False, n_samples: Optional[int] = None,
                 manual_bilerp: bool = False, return_aux: bool = False,
                 app_mask_override: Optional[torch.Tensor] = None):
    """TensorVMKeyframeTimeKplane.forward + render_pts
    (models/tensorf_keyframe.py:613-755), non-NDC, non-contracted branch.

    ``jitter`` (N,1) replaces the CPU-RNG draw at tensorf_base.py:305 (training only);
    ``random_bg`` replaces the ``torch.rand((1,)) < 0.5`` draw at tensorf_keyframe.py:740.
    ``app_mask_override`` (N,S bool; tests only) replaces the membership test ``weight > thres`` of :719,
    which is discontinuous: weights within FP32 rounding of the threshold may land on either side, and
    a gradient comparison has to hold the membership fixed to compare like with like.
    """
    pts, z, valid = sample_ray(sc, o, d, jitter if training else None, n_samples)
    N, S = z.shape
    dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
    tt = (torch.ones_like(o[..., -1:]) * t).view(-1, 1, 1).expand(N, S, 1)
    xyz = normalize_coord(sc, pts)

    if transfer_vel:
        base = torch.zeros_like(tt)
    else:
        base = keyframe_snap(sc, tt)

    if sc.alpha_volume is not None and not training:          # :656-661
        keep = sample_alpha(sc, xyz[valid]) > 0
        valid = valid.clone()
        valid[valid.clone()] = keep

    sigma = torch.zeros(N, S)
    rgb = torch.zeros(N, S, 3)
    mask_dim = sc.mask_field[-1][0].shape[0] if sc.mask_field is not None else 3
    mask = torch.zeros(N, S, mask_dim)

    xyz_adv = xyz
    if valid.any():
        if sc.use_vel:
            key = torch.isclose(tt, base)
            not_key = (~key[..., 0]) & valid
            if not_key.any():
                adv = integrate_pos(sc, xyz[not_key], tt[not_key], base[not_key])
                xyz_adv = xyz.index_put((not_key,), adv)
            xyzt_eval = torch.cat([xyz_adv, normalize_time_coord(sc, base)], dim=-1)
        else:
            xyzt_eval = torch.cat([xyz, normalize_time_coord(sc, tt)], dim=-1)
        feat = density_feature(sc, xyzt_eval[valid], manual_bilerp)
        sigma = sigma.index_put((valid,), feature2density(sc, feat))
    else:
        xyzt_eval = torch.cat([xyz, normalize_time_coord(sc, tt)], dim=-1)

    alpha, weight, _ = raw2alpha(sigma, dists * sc.distance_scale)
    app_mask = weight > sc.ray_march_weight_thres
    if app_mask_override is not None:
        app_mask = app_mask_override
    viewdirs = d.view(-1, 1, 3).expand(N, S, 3)
    if app_mask.any():
        af = app_feature(sc, xyzt_eval[app_mask], manual_bilerp)
        p_app = xyzt_eval[..., :-1][app_mask]
        if sc.shading_mode == "SH":
            c = sh_render(viewdirs[app_mask], af)
        else:
            c = mlp_pe_render(sc, p_app, viewdirs[app_mask], af)
        rgb = rgb.index_put((app_mask,), c)
        if sc.mask_field is not None:
            mask = mask.index_put((app_mask,), mask_field_forward(sc.mask_field, p_app))

    acc = torch.sum(weight, -1)
    rgb_map = torch.sum(weight[..., None] * rgb, -2)
    if white_bg or (training and random_bg):
        rgb_map = rgb_map + (1.0 - acc[..., None])
    rgb_map = rgb_map.clamp(0, 1)
    depth = torch.sum(weight * z, -1) + (1.0 - acc) * sc.far
    mask_map = torch.sum(weight[..., None] * mask, -2)
    if return_aux:
        aux = dict(valid=valid, app_mask=app_mask, z=z, xyz=xyz, xyz_adv=xyz_adv, sigma=sigma,
                   alpha=alpha, rgb=rgb)
        return rgb_map, depth, acc, weight, mask_map, aux
    return rgb_map, depth, acc, weight, mask_map


def render(sc: Scene, t: float, o: torch.Tensor, d: torch.Tensor, *, ray_chunk: int = 2048,
           white_bg: bool, training: bool, jitter: Optional[torch.Tensor] = None,
           random_bg: Optional[Sequence[bool]] = None, transfer_vel: bool = False,
           n_samples: Optional[int] = None):
    """Renderer.forward chunk loop (models/renderer.py:22-56) over flattened rays."""
    o = o.reshape(-1, 3)
    d = d.reshape(-1, 3)
    outs = [[] for _ in range(5)]
    n = o.shape[0]
    n_chunks = n // ray_chunk + int(n % ray_chunk > 0)
    for c in range(n_chunks):
        sl = slice(c * ray_chunk, (c + 1) * ray_chunk)
        r = render_chunk(sc, t, o[sl], d[sl], white_bg=white_bg, training=training,
                         jitter=None if jitter is None else jitter[sl],
                         random_bg=bool(random_bg[c]) if random_bg is not None else False,
                         transfer_vel=transfer_vel, n_samples=n_samples)
        for k in range(5):
            outs[k].append(r[k])
    return tuple(torch.cat(x, 0) for x in outs)


# --------------------------------------------------------------------------------------
# a22: PDE (divergence + transport) loss
# --------------------------------------------------------------------------------------
def occupancy_filter(sc: Scene, points_n: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """The no-grad occupancy test of models/nvfi.py:50-64. points_n normalised (P,3), t (P,1).
    Returns a bool mask (P,)."""
    with torch.no_grad():
        base = keyframe_snap(sc, t)
        prev = integrate_pos(sc, points_n, t, base)
        xyzt = torch.cat([prev, normalize_time_coord(sc, base)], dim=-1)
        sigma = feature2density(sc, density_feature(sc, xyzt))
        alpha = 1 - torch.exp(-sigma * 0.01 * 25)
        return alpha >= sc.alpha_mask_thres


def vel_jacobian(sc: Scene, xyzt: torch.Tensor):
    """Value (P,6) and Jacobian (P,6,4) of VelBasis.forward, forward-mode through the MLP
    (equivalent of vmap(jacrev(u_func)) at models/nvfi.py:69-72), double precision optional."""
    def u(x):
        return vel_full(sc, x[None])[0]
    jac = torch.func.vmap(torch.func.jacfwd(u))(xyzt)
    return vel_full(sc, xyzt), jac


def pde_loss_from_points(sc: Scene, xyzt: torch.Tensor) -> torch.Tensor:
    """models/nvfi.py:69-84 on already-filtered points xyzt (P,4)."""
    val, jac = vel_jacobian(sc, xyzt)
    vel, a = val[..., :3], val[..., 3:]
    div = jac[..., 0, 0] + jac[..., 1, 1] + jac[..., 2, 2]
    transport = torch.einsum("poi,pi->po", jac[..., :3, :3], vel) + jac[..., :3, 3] - a
    return torch.mean(div ** 2) * 5 + torch.mean(transport ** 2) * 0.1


def vel_loss(sc: Scene, points_n: torch.Tensor, t: torch.Tensor):
    """NVFi.get_vel_loss (models/nvfi.py:42-84) with the random draws passed in.
    Returns python 0.0 when no point is occupied (:66-67)."""
    keep = occupancy_filter(sc, points_n, t)
    xyzt = torch.cat([points_n, t], dim=-1)[keep]
    if xyzt.shape[0] == 0:
        return 0.0
    return pde_loss_from_points(sc, xyzt)


# --------------------------------------------------------------------------------------
# (f)-1: dense alpha volume (next row), used by the eval prelude
# --------------------------------------------------------------------------------------
def compute_alpha(sc: Scene, xyzt_locs: torch.Tensor, length: torch.Tensor, transfer: bool = False):
    """models/tensorf_keyframe.py:508-537."""
    pts = normalize_coord(sc, xyzt_locs[..., :3])
    t = xyzt_locs[..., -1:]
    base = torch.zeros_like(t) if transfer else keyframe_snap(sc, t)
    prev = integrate_pos(sc, pts, t, base)
    xyzt = torch.cat([prev, normalize_time_coord(sc, base)], dim=-1)
    sigma = feature2density(sc, density_feature(sc, xyzt))
    return 1 - torch.exp(-sigma * length)
