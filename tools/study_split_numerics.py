"""CPU study for round 2: operand-splitting schemes for the velocity-MLP GEMMs on the tensor cores.
Products are formed exactly (float64) from the rounded operands and accumulated in float64, so
the figures isolate the OPERAND rounding of each scheme (tensor cores accumulate in FP32).

   fp32        reference arithmetic of the PyTorch path
   tf32        one TF32 pass
   tf32x3      A_hi W_hi + A_hi W_lo + A_lo W_hi, all TF32 (the product path today): 3 MMA units
   tf32+bf16   A_hi W_hi in TF32, the two corrections in BF16 (K = 16 per MMA): 2 MMA units
   fp16x3      2-way FP16 split, 3 products in FP16 (K = 16): 1.5 MMA units
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200.scenes import build_scene

torch.manual_seed(0)


def tf32(x):
    xi = x.float().contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32)


def bf16(x):
    return x.float().to(torch.bfloat16).float()


def fp16(x):
    return x.float().to(torch.float16).float()


def mm(a, w):  # exact products, float64 accumulation
    return a.double() @ w.double().t()


def gemm(a, w, scheme):
    a, w = a.float(), w.float()
    if scheme == "fp32":
        return mm(a, w)
    if scheme == "tf32":
        return mm(tf32(a), tf32(w))
    ah, wh = tf32(a), tf32(w)
    al, wl = a - ah, w - wh
    if scheme == "tf32x3":
        return mm(ah, wh) + mm(ah, tf32(wl)) + mm(tf32(al), wh)
    if scheme == "tf32+bf16":
        return mm(ah, wh) + mm(bf16(ah), bf16(wl)) + mm(bf16(al), bf16(wh))
    if scheme == "fp16x3":
        ah, wh = fp16(a), fp16(w)
        al, wl = fp16(a - ah), fp16(w - wh)
        return mm(ah, wh) + mm(ah, wl) + mm(al, wh)
    raise ValueError(scheme)


cfg, nv, sd = build_scene("bat", grid=(16, 16, 16), device="cpu")
keys = ["1", "3.0", "4.0", "5.0", "6.0", "7.0"]
Ws = [(sd[f"vel_net.weight_net.{k}.weight"], sd[f"vel_net.weight_net.{k}.bias"]) for k in keys]
n = 8192
q = torch.rand(n, 4) * 2 - 1
q[:, 3] = torch.rand(n)
enc = torch.cat([q] + [f(q * s) for s in (1, 2, 4) for f in (torch.sin, torch.cos)], -1)


def forward(scheme):
    a = enc
    for i, (w, b) in enumerate(Ws):
        h = gemm(a, w, scheme) + b.double()
        a = torch.nn.functional.silu(h).float() if i < 5 else h
    return a


ref = forward("fp32") if False else None
# float64 truth: exact weights / activations
a = enc.double()
for i, (w, b) in enumerate(Ws):
    h = a @ w.double().t() + b.double()
    a = torch.nn.functional.silu(h) if i < 5 else h
truth = a
print(f"velocity MLP 28-128x5-6, {n} random inputs; max |w| {float(truth.abs().max()):.3f}")
for scheme in ("fp32", "tf32", "tf32x3", "tf32+bf16", "fp16x3"):
    out = forward(scheme)
    err = (out - truth).abs()
    print(f"  {scheme:10s} max abs {float(err.max()):.2e}  rms {float(err.pow(2).mean().sqrt()):.2e}  "
          f"rel-to-max {float(err.max() / truth.abs().max()):.2e}")


# ------------------------------------------------------------------------------------------
# End to end: the oracle's render of the golden scenes with the velocity-MLP GEMMs replaced by each
# scheme, against the reference's own outputs (tests/golden/*.npz) under the 1e-4 parity metric.
# ------------------------------------------------------------------------------------------
def end_to_end():
    from oracle import nvfi_oracle as O
    from tests.helpers import GOLDEN_SCENES, Golden

    orig = O.mlp_forward

    def make(scheme):
        def mlp_forward(layers, x, act):
            if len(layers) != 6 or layers[0][0].shape[1] != 28:      # only the velocity weight net
                return orig(layers, x, act)
            h = x
            for i, (w, b) in enumerate(layers):
                h = (gemm(h, w, scheme) + b.double()).float()
                if i < len(layers) - 1:
                    h = act(h)
            return h
        return mlp_forward

    print("\nend to end (oracle render with emulated velocity GEMMs vs reference outputs; metric max |d| / max(|ref|, 1)):")
    for name in GOLDEN_SCENES:
        g = Golden(name)
        sc = g.scene()
        o, d = g.rays()
        for case_name in ("eval1", "eval4"):      # one RK2 step / extrapolated (many steps)
            case = g.case(case_name)
            row = []
            for scheme in ("fp32", "tf32", "tf32x3", "tf32+bf16", "fp16x3"):
                O.mlp_forward = make(scheme) if scheme != "fp32" else orig
                try:
                    with torch.no_grad():
                        out = O.render(sc, float(case["t"]), o, d, ray_chunk=g.ray_chunk,
                                       white_bg=bool(g.cfg.dataset.white_background), training=False)
                finally:
                    O.mlp_forward = orig
                errs = []
                for k, i in (("rgb", 0), ("depth", 1), ("acc", 2), ("weights", 3)):
                    ref = torch.from_numpy(case[k]).float()
                    errs.append(float(((out[i] - ref).abs() / ref.abs().clamp_min(1.0)).max()))
                row.append(f"{scheme} {max(errs):.1e}")
            print(f"  {name:12s} {case_name} t={float(case['t']):.2f}: " + "  ".join(row))


if __name__ == "__main__":
    end_to_end()
