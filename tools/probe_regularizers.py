"""Timing probe of the fused plane regularisers against the reference's torch expressions at the
final grid (199^3, K = 16): density_L1 + TV_loss_density + TV_loss_app, loss and backward
(train_nvfi.py:210-224)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import _lib
from nvfi_b200.scenes import build_scene
from tests.test_gpu_regularizers import RefTVLoss

cfg, nv, _ = build_scene("bat", grid=(199, 199, 199))
nv.requires_grad_(True)
f = nv.nvfi
reg = RefTVLoss(1.0)


def fused():
    nv.zero_grad(set_to_none=True)
    (f.density_L1() * 8e-4 + f.TV_loss_density(reg) + f.TV_loss_app(reg)).backward()


def eager():
    nv.zero_grad(set_to_none=True)
    tot = 0
    for k in range(3):
        ds, dt, as_ = f.density_plane_space[k], f.density_plane_time[k], f.app_plane_space[k]
        tot = tot + (torch.mean(torch.abs(ds)) + torch.mean(torch.abs(1 - dt))) * 8e-4
        tot = tot + (reg(ds) + reg(dt, t=True)) * 1e-2 + reg(as_) * 1e-2
    tot.backward()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_f, t_e = timeit(fused), timeit(eager)
nbytes = sum(p.numel() * 4 for k, p in f.named_parameters() if "plane" in k and "app_plane_time" not in k)
print(f"regularisers (9 planes, {nbytes/1e6:.1f} MB): fused {t_f:.3f} ms, torch eager {t_e:.3f} ms, x{t_e/t_f:.1f}")
_lib.load().nvfi_profile_enable(1)
fused()
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
_lib.load().nvfi_profile_enable(0)
for name, (ms, cnt) in prof.items():
    print(f"  {name}: {cnt} launches, {ms:.3f} ms total")
tv_ms = sum(ms for name, (ms, c) in prof.items() if "tv" in name)
tv_bytes = 2 * sum(p.numel() * 4 for k, p in f.named_parameters()
                   if "plane" in k and "app_plane_time" not in k)
print(f"  k_tv_plane: algorithmic (1 read + 1 gradient write) {tv_bytes/1e6:.1f} MB -> {tv_bytes/tv_ms/1e6:.0f} GB/s")
