"""Where one training iteration at the shipped batch size (2048 rays, train_nvfi.py:198-252) spends its
time: device time per kernel (library event hook) against the wall clock of the call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import _lib
from nvfi_b200.scenes import build_scene, frame_rays

cfg, nv, _ = build_scene("bat", step_ratio=1.79)
f = nv.nvfi
nv.requires_grad_(True)
f.train()
o_all, d_all = frame_rays(800, 800)
torch.manual_seed(0)
idx = torch.randperm(o_all.shape[0])[:2048]
o, d = o_all[idx].contiguous().cuda(), d_all[idx].contiguous().cuda()
target = torch.rand(2048, 3, device="cuda")


def step():
    nv.zero_grad(set_to_none=True)
    rgb, *_ = f.render_rays(0.33, o, d, white_bg=True, ray_chunk=2048)
    loss = torch.nn.functional.mse_loss(rgb, target)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 50
for _ in range(N):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
_lib.profile_read(reset=True)
_lib.profile_enable(True)
for _ in range(10):
    step()
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
_lib.profile_enable(False)
dev = sum(ms for ms, c in prof.values()) / 10
print(f"2048 rays x 192 samples, t=0.33: wall {wall:.3f} ms/iter, library kernels {dev:.3f} ms/iter "
      f"({sum(c for ms, c in prof.values()) / 10:.0f} launches)")
for name, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {name:24s} {ms / 10:8.3f} ms/iter  {c / 10:4.0f} launches")
