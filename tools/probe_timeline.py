"""Phase timeline of CTA 0 of the tensor-core backward (clock64 deltas between marks).

The marks are compiled out of the product build.  Build a variant with them and point the loader at it:
    tools/build_variant.sh tl -DNVFI_TIMELINE
    NVFI_LIB_PATH=$PWD/nvfi_b200/_variants/libtl.so python tools/probe_timeline.py      (on the GPU box)
Marks of worker thread 0 (tags < 1000) and of the issuer warp (tags >= 1000) are taken on the same SM clock;
tags 200+l / 300+l / 400+l are pseudo-events of a stashed forward layer: the last of the 16 kready[3]
arrivals, the issuer's commit of the next accumulator, the moment the issuer sees it published.
TL_TOP=n prints the n largest transitions (default 70)."""
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
buf = torch.zeros(80000, dtype=torch.int64, device="cuda")
lib.nvfi_debug_timeline(buf.data_ptr(), buf.numel())
step(); torch.cuda.synchronize()
lib.nvfi_debug_timeline(None, 0)
b = buf.cpu().tolist()
half = len(b) // 2
from nvfi_b200 import engine
h16 = engine.get_mlp_mode() == "f16x3"
parts = [("worker thread 0", b[:half]), ("issuer warp", b[half:])] if h16 else [("thread 0", b)]
for who, bb in parts:
    ev = [(bb[i], bb[i + 1]) for i in range(0, len(bb), 2) if bb[i + 1] != 0]
    print(f"== {who}: {len(ev)} events")
    if len(ev) < 3:
        continue
    # skip the first tile (warm-up)
    starts = [i for i, (t, c) in enumerate(ev) if t == (0 if who != "issuer warp" else ev[0][0])]
    d = collections.defaultdict(list)
    for (t0, c0), (t1, c1) in zip(ev, ev[1:]):
        d[(t0, t1)].append(c1 - c0)
    tot = sum(sum(v) for v in d.values())
    for k in sorted(d, key=lambda k: -sum(d[k]))[:int(os.environ.get('TL_TOP', '70'))]:
        v = d[k]
        print(f"{k[0]:5d}->{k[1]:5d}  n={len(v):4d} mean {sum(v)/len(v):9.0f} cyc  share {100*sum(v)/tot:5.1f}%")
    tiles = [c for t, c in ev if t == 0]
    if len(tiles) > 2:
        print("cycles per tile", (tiles[-1] - tiles[1]) / (len(tiles) - 2))
