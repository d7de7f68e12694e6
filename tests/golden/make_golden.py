"""Generate golden vectors by running the REFERENCE implementation (CPU, FP32).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/<scene>.npz.  Each file holds the synthetic parameters (seeded,
nvfi_b200/synth.py), the inputs of every case (rays, jitter draws, random-bg draws, PDE
points) and the outputs of the unmodified reference code (models/renderer.py,
models/nvfi.py, models/tensorf_keyframe.py).  The reference has no golden vectors of its
own, so these files are what pins the oracle (tests/test_oracle_golden.py)
and, through it and directly, the CUDA path (tests/test_gpu_parity.py).
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REF = os.environ.get("NVFI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
for m in ("lpips", "imageio", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))

import models as ref_models                      # noqa: E402  (the reference package)
from models.camera import Ray                    # noqa: E402
from utils import CfgNode                        # noqa: E402

from nvfi_b200 import configs, synth             # noqa: E402

torch.set_num_threads(8)


class DrawRecorder:
    """Records torch.rand_like / torch.rand draws made inside the reference render so the
    same numbers can be fed to the oracle / CUDA path (SURVEY.md section 7 'RNG parity')."""

    def __init__(self):
        self.jitter, self.bg = [], []

    def __enter__(self):
        self._rl, self._r = torch.rand_like, torch.rand
        rec = self

        def rand_like(x, *a, **k):
            out = rec._rl(x, *a, **k)
            rec.jitter.append(out.clone())
            return out

        def rand(*a, **k):
            out = rec._r(*a, **k)
            if tuple(out.shape) == (1,):
                rec.bg.append(bool(out.item() < 0.5))
            return out

        torch.rand_like, torch.rand = rand_like, rand
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.rand = self._rl, self._r


def build_reference(cfg, grid, K, sd):
    c = CfgNode(json.loads(json.dumps(cfg)))
    c.nvfi.num_keyframes = K
    aabb = synth.aabb_from_cfg(cfg)
    nv = ref_models.NVFi(c, "cpu", aabb, list(grid), [cfg.dataset.near, cfg.dataset.far])
    missing, unexpected = nv.load_state_dict({"nvfi." + k: v for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    # vel.vel_net.* aliases vel_net.*; frequency_bands are buffers; everything else must load
    bad = [m for m in missing if "frequency_bands" not in m and ".vel.vel_net." not in m and m != "nvfi.aabb"]
    assert not bad, bad
    return c, nv


def scene_specs():
    bat = configs.get_config("bat")
    chess = configs.get_config("chessboard")
    sh = configs.get_config("bat", shadingMode="SH", app_dim=27, tmax=1.0, max_n_samples=40)
    return [
        dict(name="bat_small", cfg=bat, grid=[22, 18, 26], K=16, seed=233, theta=30.0),
        dict(name="chess_small", cfg=chess, grid=[20, 24, 16], K=4, seed=234, theta=200.0),
        dict(name="sh_small", cfg=sh, grid=[16, 16, 16], K=2, seed=235, theta=100.0),
    ]


def loss_weights(gen, n, S):
    return dict(wr=torch.randn(n, 3, generator=gen), wd=torch.randn(n, generator=gen) * 0.1,
                wa=torch.randn(n, generator=gen), ww=torch.randn(n, S, generator=gen) * 0.1)


def scalar_loss(out, lw):
    rgb, depth, acc, w, _ = out
    return (rgb * lw["wr"]).sum() + (depth * lw["wd"]).sum() + (acc * lw["wa"]).sum() + (w * lw["ww"]).sum()


def main():
    for spec in scene_specs():
        cfg, grid, K = spec["cfg"], spec["grid"], spec["K"]
        sd = synth.synth_state(cfg, grid, K, seed=spec["seed"])
        c, nv = build_reference(cfg, grid, K, sd)
        field = nv.nvfi
        white = bool(cfg.dataset.white_background)
        ray_chunk = 160
        renderer = ref_models.Renderer(nv, 0, 0, ray_chunk)
        arrays = {"meta": np.array(json.dumps(dict(
            name=spec["name"], cfg=cfg, grid=grid, K=K, seed=spec["seed"], ray_chunk=ray_chunk,
            nSamples=int(field.nSamples), stepSize=float(field.stepSize))))}
        for k, v in sd.items():
            arrays["sd/" + k] = v.numpy()

        H = W = 20
        focal = synth.blender_focal(W)
        pose = synth.pose_spherical(spec["theta"], -30.0, 4.0)
        if spec["name"] == "chess_small":
            pose[2, 3] += 3.0      # the chessboard box spans z in [0,6]
        cam = ref_models.Camera(pose, H, W, focal, torch.zeros(H, W, 3), cfg.dataset.near, cfg.dataset.far)
        rays = cam.rays
        o = rays.ray_origins.reshape(-1, 3).clone()
        d = rays.ray_directions.reshape(-1, 3).clone()
        arrays["pose"] = pose.numpy()
        arrays["cam"] = np.array([H, W, focal], dtype=np.float64)
        arrays["rays_o"], arrays["rays_d"] = o.numpy(), d.numpy()
        n = o.shape[0]
        S = int(field.nSamples)
        tmax = float(cfg.nvfi.tmax)
        tsf = tmax / (K - 1)

        def record(case, out, extra=None):
            names = ("rgb", "depth", "acc", "weights", "mask_map")
            for nm, v in zip(names, out):
                arrays[f"case/{case}/{nm}"] = v.detach().reshape(n, -1).numpy() if nm in ("rgb", "weights", "mask_map") \
                    else v.detach().reshape(n).numpy()
            for k2, v in (extra or {}).items():
                arrays[f"case/{case}/{k2}"] = v

        # ---- eval cases ------------------------------------------------------------
        t_list = [0.0, 0.33 * tmax / 0.75, 2 * tsf, tmax, tmax + 0.25]
        for i, t in enumerate(t_list):
            out = renderer.render(float(t), Ray(o.view(H, W, 3), d.view(H, W, 3), 0, 0),
                                  white_background=white, mode="test")
            record(f"eval{i}", out, dict(t=np.float64(t)))
        out = renderer.render(0.2, Ray(o, d, 0, 0), white_background=white, mode="test", transfer_vel=True)
        record("transfer", out, dict(t=np.float64(0.2)))

        # ---- train cases (jitter + grads) --------------------------------------------
        gen = torch.Generator().manual_seed(spec["seed"] + 1000)
        lw = loss_weights(gen, n, S)
        for kk, v in lw.items():
            arrays[f"lossw/{kk}"] = v.numpy()
        params = dict(field.named_parameters())
        # vel.vel_net.* are the same Parameter objects as vel_net.*: named_parameters dedups
        for i, t in enumerate([0.33 * tmax / 0.75, 3 * tsf if K > 3 else tsf]):
            torch.manual_seed(spec["seed"] + 7 + i)
            for p in params.values():
                p.grad = None
            with DrawRecorder() as rec:
                out = renderer.render(float(t), Ray(o, d, 0, 0), white_background=white, mode="train")
            loss = scalar_loss(out, lw)
            loss.backward()
            jit = torch.cat(rec.jitter, 0)
            assert jit.shape == (n, 1), jit.shape
            extra = dict(t=np.float64(t), jitter=jit.numpy(), random_bg=np.array(rec.bg, dtype=np.bool_),
                         loss=np.float64(loss.item()))
            for pn, p in params.items():
                if p.grad is None:
                    continue
                gflat = p.grad.reshape(-1)
                if "plane" in pn:       # subsample big plane grads to keep fixtures small
                    extra[f"grad_sub/{pn}"] = gflat[::7].numpy()
                    extra[f"grad_norm/{pn}"] = np.float64(gflat.double().norm().item())
                else:
                    extra[f"grad/{pn}"] = p.grad.numpy().copy()
            record(f"train{i}", out, extra)

        # ---- per-op cases --------------------------------------------------------------
        field.eval()
        gen = torch.Generator().manual_seed(spec["seed"] + 2000)
        P = 300
        xyz = torch.rand(P, 3, generator=gen) * 2.3 - 1.15          # some points out of range
        tt = torch.rand(P, 1, generator=gen) * (tmax + 0.2)
        tt[:20] = torch.round(tt[:20] / tsf).clamp(0, K - 1) * tsf   # exact keyframes
        with torch.no_grad():
            base = torch.round((tt / field.time_scale_factor).clamp(0.0, K - 1)) * field.time_scale_factor
            adv = field.integrate_pos(xyz.clone(), tt.clone(), base.clone())
            xyzt = torch.cat([adv, field.normalize_time_coord(base)], -1)
            dfeat = field.compute_densityfeature(xyzt)
            sigma = field.feature2density(dfeat, {})
            afeat = field.compute_appfeature(xyzt)
            vfull = field.vel_net(torch.cat([xyz, tt], -1))
            vgate = field.vel(torch.cat([xyz, tt], -1))
            # forward (negative offset) advection as used by train_segm.py:161
            fwd = field.integrate_pos(xyz.clone(), torch.zeros(P, 1), tt.clone())
        arrays.update({"op/xyz": xyz.numpy(), "op/t": tt.numpy(), "op/base": base.numpy(),
                       "op/adv": adv.numpy(), "op/dfeat": dfeat.numpy(), "op/sigma": sigma.numpy(),
                       "op/afeat": afeat.numpy(), "op/vfull": vfull.numpy(), "op/vgate": vgate.numpy(),
                       "op/adv_fwd": fwd.numpy()})

        # ---- PDE loss ----------------------------------------------------------------------
        field.train()
        n_pts = 1536
        torch.manual_seed(spec["seed"] + 3000)
        st = torch.get_rng_state()
        pts_draw = torch.rand(n_pts, 3)
        t_draw = torch.rand(n_pts, 1)
        torch.set_rng_state(st)
        for p in params.values():
            p.grad = None
        lv = nv.get_vel_loss(n_pts)
        arrays["pde/points_u"] = pts_draw.numpy()       # U(0,1) draws, before aabb scaling
        arrays["pde/t"] = t_draw.numpy()
        if isinstance(lv, float):
            arrays["pde/loss"] = np.float64(lv)
            arrays["pde/empty"] = np.array(True)
        else:
            lv.backward()
            arrays["pde/loss"] = np.float64(lv.item())
            arrays["pde/empty"] = np.array(False)
            for pn, p in params.items():
                if p.grad is not None:
                    arrays[f"pde/grad/{pn}"] = p.grad.numpy().copy()

        # ---- alpha mask (next row f-1) + eval with the mask --------------------------------
        field.eval()
        mask_grid = [g for g in grid]
        with torch.no_grad():
            new_aabb = field.updateAlphaMask(tuple(mask_grid))
        arrays["alpha/grid"] = np.array(mask_grid)
        arrays["alpha/volume"] = field.alphaMask.alpha_volume.numpy().astype(np.uint8)
        arrays["alpha/new_aabb"] = new_aabb.numpy()
        out = renderer.render(float(t_list[1]), Ray(o, d, 0, 0), white_background=white, mode="test")
        record("eval_alpha", out, dict(t=np.float64(t_list[1])))
        out = renderer.render(float(t_list[4]), Ray(o, d, 0, 0), white_background=white, mode="test")
        record("eval_alpha_extrap", out, dict(t=np.float64(t_list[4])))

        # ---- mask field composite (config 5), mask_dim=3 (Renderer reshapes to 3) ----------
        torch.manual_seed(spec["seed"] + 4000)
        mf = ref_models.MaskField(n_layer=4, n_dim=128, input_dim=3, skips=[], mask_dim=3, mask_act="softmax")
        field.mask_field = mf
        for li, lin in enumerate(list(mf.point_fc) + [mf.mask_fc]):
            arrays[f"maskfield/{li}/weight"] = lin.weight.detach().numpy()
            arrays[f"maskfield/{li}/bias"] = lin.bias.detach().numpy()
        out = renderer.render(0.2, Ray(o, d, 0, 0), white_background=white, mode="test", transfer_vel=True)
        record("transfer_mask", out, dict(t=np.float64(0.2)))
        field.mask_field = None

        path = os.path.join(HERE, spec["name"] + ".npz")
        np.savez_compressed(path, **arrays)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB", "nSamples", S)


if __name__ == "__main__":
    main()
