"""Runs the reference's UNMODIFIED training driver (train_nvfi.py:21-369, copied verbatim into
baseline/_ref by tools/install_reference.py) on a tiny synthetic Blender-format dataset, either with the
reference's own ``models`` package (--impl reference, CPU) or with ``nvfi_b200.models`` aliased in its
place (--impl nvfi_b200, CUDA): the drop-in claim of INTEGRATION.md, executed.

    python tests/dropin/harness.py --impl nvfi_b200 --device cuda --work /tmp/x --iters 8 [--full]

Only third-party modules that are not installed in the image are stubbed (imageio -> PIL, matplotlib,
lpips); nothing of the reference is patched.  Prints the driver's own "[TRAIN] Iter: ..." lines."""
import argparse
import json
import math
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stub_third_party():
    import numpy as np
    from PIL import Image
    if "imageio" not in sys.modules:
        try:
            import imageio  # noqa: F401
        except Exception:
            m = types.ModuleType("imageio")
            v2 = types.ModuleType("imageio.v2")
            v2.imread = lambda f: np.array(Image.open(f))
            m.v2 = v2
            m.imread = v2.imread
            m.imwrite = lambda f, a: Image.fromarray(a).save(f)
            sys.modules["imageio"], sys.modules["imageio.v2"] = m, v2
    for name in ("matplotlib", "matplotlib.pyplot", "lpips"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def make_dataset(base, n_train=6, size=24, tmax=0.75, K=4):
    """transforms_{train,val,test}.json + RGBA PNGs (datasets/load_blender.py:68-124): a coloured disc
    that moves with time, cameras on the Blender circle."""
    import numpy as np
    from PIL import Image
    sys.path.insert(0, ROOT)
    from nvfi_b200 import synth
    if os.path.exists(os.path.join(base, "transforms_train.json")):
        return
    os.makedirs(base, exist_ok=True)
    rng = np.random.RandomState(0)
    key_times = [tmax * k / (K - 1) for k in range(K)]

    def frames(split, times):
        out = []
        os.makedirs(os.path.join(base, split), exist_ok=True)
        for i, t in enumerate(times):
            yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
            cx, cy = 0.35 + 0.3 * t, 0.5
            disc = ((xx - cx) ** 2 + (yy - cy) ** 2) < 0.06
            img = np.zeros((size, size, 4), np.uint8)
            img[..., 0] = (255 * xx).astype(np.uint8)
            img[..., 1] = (255 * yy).astype(np.uint8)
            img[..., 2] = rng.randint(0, 255)
            img[..., 3] = np.where(disc, 255, 0)
            Image.fromarray(img).save(os.path.join(base, split, f"r_{i:03d}.png"))
            pose = synth.pose_spherical(-180.0 + 60.0 * i, -30.0, 4.0)
            out.append({"file_path": f"./{split}/r_{i:03d}", "time": float(t),
                        "transform_matrix": [[float(x) for x in row] for row in pose.tolist()]})
        return {"camera_angle_x": synth.BLENDER_CAMERA_ANGLE_X, "frames": out}

    train_t = [key_times[i % K] if i % 2 == 0 else 0.1 + 0.09 * i for i in range(n_train)]
    for split, times in (("train", train_t), ("val", [0.3, 0.6]), ("test", [0.2])):
        json.dump(frames(split, times), open(os.path.join(base, f"transforms_{split}.json"), "w"))


def make_config(ref_root, work, device, iters, full):
    import yaml
    cfg = yaml.load(open(os.path.join(ref_root, "config", "InDoorObj", "bat.yaml")), Loader=yaml.FullLoader)
    cfg["experiment"].update(device=device, logdir=os.path.join(work, "logs") + "/", train_iters=iters,
                             print_every=1, validate_every=(4 if full else 10 ** 6), save_every=10 ** 6,
                             vel_reg_n_pts=2048, vel_reg_weight=(1 if full else 0))
    cfg["pbar"]["progress_refresh_rate"] = 10 ** 6
    cfg["dataset"].update(basedir=os.path.join(work, "data"), half_res=False)
    cfg["renderer"].update(n_rays=256)
    cfg["nvfi"].update(N_voxel_init=16 ** 3, N_voxel_final=24 ** 3, upsamp_list=([3] if full else [10 ** 6]),
                       update_AlphaMask_list=([5] if full else []), num_keyframes=4, num_keyframes_end=4,
                       density_shift=(0 if full else -10))
    path = os.path.join(work, f"cfg_{'full' if full else 'cmp'}.yaml")
    yaml.dump(cfg, open(path, "w"))
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["nvfi_b200", "reference"], required=True)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--work", required=True)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--full", action="store_true",
                    help="PDE loss, one upsample step, alpha-mask update + shrink, validation renders")
    ap.add_argument("--checkpoint", type=int, default=0, help="resume from this checkpoint (train_nvfi.py --checkpoint)")
    ap.add_argument("--ref-root", default=os.environ.get("NVFI_REFERENCE") or os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(a.ref_root, "train_nvfi.py")):
        sys.exit(f"no reference driver at {a.ref_root} (python tools/install_reference.py)")
    os.makedirs(a.work, exist_ok=True)
    stub_third_party()
    make_dataset(os.path.join(a.work, "data"))
    cfg = make_config(a.ref_root, a.work, a.device, a.iters, a.full)
    sys.path.insert(0, a.ref_root)          # utils/, datasets/ (and models/ for --impl reference)
    if a.impl == "nvfi_b200":
        sys.path.insert(0, ROOT)
        import nvfi_b200.models as M
        sys.modules["models"] = M            # `from models import *` in train_nvfi.py:16 now binds nvfi_b200's classes
    import torch
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    os.chdir(a.work)
    sys.argv = ["train_nvfi.py", "--config", cfg, "--static_dynamic"]
    if a.checkpoint:
        sys.argv += ["--checkpoint", str(a.checkpoint)]
    runpy.run_path(os.path.join(a.ref_root, "train_nvfi.py"), run_name="__main__")
    import models
    print("MODELS_PACKAGE " + os.path.abspath(models.__file__))


if __name__ == "__main__":
    main()
