"""Golden vectors for reference branches the three scene fixtures do not exercise:

  * sample_ray's slab-entry branch (models/tensorf_base.py:294-300): a camera whose origin is outside
    the box on ALL three axes, so the chunk-global inside test is False, plus rays with a zero
    direction component (the `d == 0 -> 1e-6` substitution);
  * MaskField with mask_dim = 8 (config 5), rendered through NVFi.render_ray_transfer /
    NVFi.render_ray (models/nvfi.py:27-31) — Renderer.forward cannot reshape an 8-wide mask
    (models/renderer.py:54), the field call can.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_branches.py
Writes tests/golden/branches_small.npz.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                       # noqa: E402  (sets up the reference imports)
from make_golden import ref_models, synth, configs    # noqa: E402


def main():
    cfg = configs.get_config("bat")
    grid, K, seed = [22, 18, 26], 16, 233
    sd = synth.synth_state(cfg, grid, K, seed=seed)
    c, nv = MG.build_reference(cfg, grid, K, sd)
    field = nv.nvfi
    field.eval()
    arrays = {"meta": np.array(json.dumps(dict(name="branches_small", cfg=cfg, grid=grid, K=K, seed=seed,
                                                ray_chunk=97, nSamples=int(field.nSamples))))}
    for k, v in sd.items():
        arrays["sd/" + k] = v.numpy()

    # ---- camera outside the box on all three axes, looking at the origin --------------------------
    H = W = 12
    eye = torch.tensor([4.6, 5.2, 3.9])
    fwd = -eye / eye.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    pose = torch.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, up, -fwd, eye
    cam = ref_models.Camera(pose, H, W, 14.0, torch.zeros(H, W, 3), cfg.dataset.near, cfg.dataset.far)
    o = cam.rays.ray_origins.reshape(-1, 3).clone()
    d = cam.rays.ray_directions.reshape(-1, 3).clone()
    # rays with zero direction components (still from the outside origin)
    extra_d = torch.tensor([[-1.0, -1.1, 0.0], [0.0, -1.0, -0.8], [-1.0, 0.0, 0.0], [-0.9, -1.0, -0.75]])
    o = torch.cat([o, eye[None].expand(4, 3)], 0).contiguous()
    d = torch.cat([d, extra_d], 0).contiguous()
    assert not bool(((field.aabb[0] <= o) & (o <= field.aabb[1])).any())      # the slab branch is taken
    arrays["rays_o"], arrays["rays_d"] = o.numpy(), d.numpy()
    n = o.shape[0]
    renderer = ref_models.Renderer(nv, 0, 0, 97)
    names = ("rgb", "depth", "acc", "weights", "mask_map")
    for i, t in enumerate([0.0, 0.33, 1.0]):
        out = renderer.render(float(t), MG.Ray(o, d, 0, 0), white_background=True, mode="test")
        for nm, v in zip(names, out):
            arrays[f"case/slab{i}/{nm}"] = v.detach().reshape(n, -1).numpy() if nm in ("rgb", "weights", "mask_map") \
                else v.detach().reshape(n).numpy()
        arrays[f"case/slab{i}/t"] = np.float64(t)
    acc = arrays["case/slab1/acc"]
    assert acc.max() > 0.3, "the outside camera should still see the cube"

    # ---- training-mode forward + gradients through the slab branch ---------------------------------
    gen = torch.Generator().manual_seed(77)
    wr = torch.randn(n, 3, generator=gen)
    nv.requires_grad_(True)
    torch.manual_seed(4242)
    with MG.DrawRecorder() as rec:
        out = renderer.render(0.33, MG.Ray(o, d, 0, 0), white_background=True, mode="train")
    (out[0] * wr).sum().backward()
    arrays["case/slab_train/jitter"] = torch.cat(rec.jitter, 0).numpy()
    arrays["case/slab_train/wr"] = wr.numpy()
    arrays["case/slab_train/rgb"] = out[0].detach().numpy()
    for pn, p in field.named_parameters():
        if p.grad is not None and "plane" not in pn:
            arrays[f"case/slab_train/grad/{pn}"] = p.grad.numpy().copy()
    nv.requires_grad_(False)
    field.eval()

    # ---- MaskField(mask_dim = 8) through NVFi.render_ray_transfer / render_ray -----------------------
    torch.manual_seed(99)
    mf = ref_models.MaskField(n_layer=4, n_dim=128, input_dim=3, skips=[], mask_dim=8, mask_act="softmax")
    for i, lin in enumerate(list(mf.point_fc) + [mf.mask_fc]):
        arrays[f"maskfield8/{i}/weight"] = lin.weight.detach().numpy().copy()
        arrays[f"maskfield8/{i}/bias"] = lin.bias.detach().numpy().copy()
    field.mask_field = mf
    # rays from the regular golden rig (inside test true), one reference chunk
    pose2 = synth.pose_spherical(30.0, -30.0, 4.0)
    cam2 = ref_models.Camera(pose2, 14, 14, synth.blender_focal(14), torch.zeros(14, 14, 3), cfg.dataset.near, cfg.dataset.far)
    o2 = cam2.rays.ray_origins.reshape(-1, 3).clone()
    d2 = cam2.rays.ray_directions.reshape(-1, 3).clone()
    arrays["mask8/rays_o"], arrays["mask8/rays_d"] = o2.numpy(), d2.numpy()
    with torch.no_grad():
        for nm, fn, t in (("transfer", nv.render_ray_transfer, 0.2), ("plain", nv.render_ray, 0.33)):
            out = fn(float(t), o2, d2, True, False)
            for k2, v in zip(names, out):
                arrays[f"case/mask8_{nm}/{k2}"] = v.detach().numpy()
            arrays[f"case/mask8_{nm}/t"] = np.float64(t)
    assert arrays["case/mask8_transfer/mask_map"].shape == (o2.shape[0], 8)
    np.savez_compressed(os.path.join(HERE, "branches_small.npz"), **arrays)
    print("wrote branches_small.npz", {k: v.shape for k, v in arrays.items() if k.startswith("case/") and k.endswith("acc")})


if __name__ == "__main__":
    main()
