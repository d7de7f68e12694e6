"""Quick timing probe of the forward render (not the bench): python tools/probe_eval.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import engine
from nvfi_b200.scenes import build_scene, frame_rays

H = W = int(os.environ.get("RES", 800))
cfg, nv, _ = build_scene("bat", step_ratio=1.79)
f = nv.nvfi
f.eval()
o, d = frame_rays(H, W)
o, d = o.cuda(), d.cuda()
print("rays", o.shape[0], "S", f.nSamples)
for t in (0.0, 0.33, 1.0):
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = engine.render_forward(f.binding, o, d, t, white_bg=True, training=False, ray_chunk=2048,
                                    want_stats=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    st = out.stats.cpu().tolist()
    print(f"t={t}: {ms:.2f} ms  {o.shape[0]/ms*1e3:.3e} rays/s  valid={st[0]} adv={st[1]} app={st[2]} "
          f"acc_mean={out.acc_map.mean().item():.4f} rgb_mean={out.rgb_map.mean().item():.4f}")
