"""Per-kernel device time of one full-frame train step (library event hook), without the CPU leg."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import _lib
from nvfi_b200.scenes import build_scene, frame_rays
cfg, nv, _ = build_scene("bat", step_ratio=1.79)
f = nv.nvfi
nv.requires_grad_(True)
f.train()
o, d = frame_rays(800, 800)
o, d = o.cuda(), d.cuda()
n = o.shape[0]
gen = torch.Generator().manual_seed(1000)
target = torch.rand(n, 3, generator=gen).cuda()
jit = torch.rand(n, 1, generator=gen).cuda()
def step():
    nv.zero_grad(set_to_none=True)
    rgb, *_ = f.render_rays(0.33, o, d, white_bg=True, ray_chunk=2048, jitter=jit)
    torch.nn.functional.mse_loss(rgb, target).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.profile_read(reset=True)
_lib.profile_enable(True)
for _ in range(3):
    step()
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
_lib.profile_enable(False)
tot = sum(ms for ms, c in prof.values()) / 3
print(f"step: {tot:.1f} ms of kernels")
for name, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {name:24s} {ms / 3:8.3f} ms")
