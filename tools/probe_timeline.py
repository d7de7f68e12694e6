"""Phase timeline of CTA 0 of the tensor-core backward (clock64 deltas between marks).

The marks are compiled out of the product build: rebuild the library with
    NVFI_TIMELINE=1 python -m nvfi_b200.build --force
before sending this probe to the GPU box (and rebuild without it afterwards)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import _lib as L
from nvfi_b200.scenes import build_scene, frame_rays
lib = L.load()
cfg, nv, _ = build_scene("bat", step_ratio=1.79)
f = nv.nvfi; nv.requires_grad_(True); f.train()
o, d = frame_rays(800, 800, crop=(0, 350, 800, 100))
o, d = o.cuda(), d.cuda()
n = o.shape[0]
target = torch.rand(n, 3, device="cuda"); jit = torch.rand(n, 1)
def step():
    nv.zero_grad(set_to_none=True)
    rgb, *_ = f.render_rays(0.33, o, d, white_bg=True, ray_chunk=2048, jitter=jit)
    torch.nn.functional.mse_loss(rgb, target).backward()
step(); torch.cuda.synchronize()
buf = torch.zeros(20000, dtype=torch.int64, device="cuda")
lib.nvfi_debug_timeline(buf.data_ptr(), buf.numel())
step(); torch.cuda.synchronize()
lib.nvfi_debug_timeline(None, 0)
b = buf.cpu().tolist()
ev = [(b[i], b[i + 1]) for i in range(0, len(b), 2) if b[i + 1] != 0]
print("events", len(ev))
# per-transition statistics over tiles 2.. (skip warm-up)
d = collections.defaultdict(list)
for (t0, c0), (t1, c1) in zip(ev, ev[1:]):
    d[(t0, t1)].append(c1 - c0)
tot = 0
for k in sorted(d, key=lambda k: -sum(d[k])):
    v = d[k]; tot += sum(v)
for k in sorted(d, key=lambda k: -sum(d[k]))[:40]:
    v = d[k]
    print(f"{k[0]:4d}->{k[1]:4d}  n={len(v):4d} mean {sum(v)/len(v):9.0f} cyc  share {100*sum(v)/tot:5.1f}%")
tiles = [c for t, c in ev if t == 0]
if len(tiles) > 2: print("cycles per tile", (tiles[-1] - tiles[1]) / (len(tiles) - 2))
