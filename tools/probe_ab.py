"""A/B timing of library variants on ONE GPU box (development; tools/build_variant.sh builds the variants).

    python tools/probe_ab.py name=path/to/lib.so [name=path ...]      (first = the baseline)

Each variant runs in its own process (NVFI_LIB_PATH): full-frame bat train step (the bench workload), per-kernel
device time through the library's event hook, and the gradients' relative-norm distance to the baseline's."""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(out_path, reps):
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
    g = torch.Generator().manual_seed(7)
    target = torch.rand(n, 3, generator=g).cuda()
    jit = torch.rand(n, 1, generator=g)

    def step():
        nv.zero_grad(set_to_none=True)
        rgb, *_ = f.render_rays(0.33, o, d, white_bg=True, ray_chunk=2048, jitter=jit)
        loss = torch.nn.functional.mse_loss(rgb, target)
        loss.backward()
        return loss
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    loss = step()
    torch.cuda.synchronize()
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    grads = {k: p.grad.detach().float().cpu() for k, p in nv.named_parameters() if p.grad is not None}
    torch.save({"grads": grads, "loss": float(loss)}, out_path + ".pt")
    top = sorted(((k, v[0]) for k, v in prof.items()), key=lambda kv: -kv[1])[:6]
    json.dump({"ms_step": ms, "kernels": {k: round(v, 3) for k, v in top}, "loss": float(loss)}, open(out_path, "w"))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    import torch
    reps = int(os.environ.get("AB_REPS", "3"))
    base = None
    for spec in sys.argv[1:]:
        name, path = spec.split("=", 1)
        out = os.path.join(tempfile.gettempdir(), f"ab_{name}.json")
        env = dict(os.environ, NVFI_LIB_PATH=os.path.abspath(path))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", out, str(reps)], env=env,
                           capture_output=True, text=True, timeout=600)
        if p.returncode != 0:
            print(f"{name}: FAILED rc={p.returncode}\n{p.stderr[-1500:]}")
            continue
        r = json.load(open(out))
        g = torch.load(out + ".pt")
        line = f"{name:12s} step {r['ms_step']:8.2f} ms  loss {r['loss']:.7f}  " + "  ".join(f"{k} {v}" for k, v in r["kernels"].items())
        if base is None:
            base = g
        else:
            worst = max(((float((g["grads"][k].double() - base["grads"][k].double()).norm() /
                                base["grads"][k].double().norm().clamp_min(1e-30)), k) for k in base["grads"]))
            line += f"\n{'':12s} max grad rel-norm distance to baseline {worst[0]:.2e} ({worst[1]})"
        print(line, flush=True)
