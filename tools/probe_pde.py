"""Timing probe of NVFi.get_vel_loss (PDE loss) at the shipped size (262 144 points)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200.scenes import build_scene
cfg, nv, _ = build_scene("fallingball")
nv.requires_grad_(True)
f = nv.nvfi
n = int(cfg.experiment.vel_reg_n_pts)
for it in range(4):
    nv.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    loss = nv.get_vel_loss(n)
    e[1].record()
    if torch.is_tensor(loss):
        loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    print(f"get_vel_loss({n}): fwd(+grads) {e[0].elapsed_time(e[1]):.2f} ms, backward() {e[1].elapsed_time(e[2]):.2f} ms, loss {float(loss):.6f}")

from nvfi_b200 import _lib
_lib.profile_read(reset=True)
_lib.profile_enable(True)
for it in range(5):
    nv.zero_grad(set_to_none=True)
    loss = nv.get_vel_loss(n)
    if torch.is_tensor(loss):
        loss.backward()
torch.cuda.synchronize()
prof = _lib.profile_read(reset=True)
_lib.profile_enable(False)
for name, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {name:24s} {ms / 5:8.3f} ms/call  {c / 5:4.0f} launches")
