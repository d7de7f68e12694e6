"""Timing probe of the train step (forward + MSE + backward) — not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200.scenes import build_scene, frame_rays

cfg, nv, _ = build_scene("bat", step_ratio=1.79)
f = nv.nvfi
nv.requires_grad_(True)
f.train()
o_all, d_all = frame_rays(800, 800)
o_all, d_all = o_all.cuda(), d_all.cuda()
torch.manual_seed(0)
for n in (2048, 640000):
    idx = torch.randperm(o_all.shape[0], device="cuda")[:n]
    o, d = o_all[idx].contiguous(), d_all[idx].contiguous()
    target = torch.rand(n, 3, device="cuda")
    for t in (0.33, 0.25):
        for it in range(3):
            nv.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            rgb, depth, acc, w, _ = f.render_rays(t, o, d, white_bg=True, ray_chunk=2048)
            loss = torch.nn.functional.mse_loss(rgb, target)
            e[1].record()
            loss.backward()
            e[2].record()
            torch.cuda.synchronize()
        fw, bw = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        print(f"n={n} t={t}: fwd {fw:.2f} ms bwd {bw:.2f} ms -> {n/(fw+bw)*1e3:.3e} rays/s  loss {loss.item():.5f} "
              f"mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
