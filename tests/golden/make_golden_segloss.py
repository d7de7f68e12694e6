"""Golden vectors of the segmentation losses from the REFERENCE's own module (build container only):

    python tests/golden/make_golden_segloss.py     # writes tests/golden/segloss_small.npz

/root/reference/utils/seg_loss.py imports pytorch3d (absent from the image) for knn_points / knn_gather
only; those two names are provided by the oracle's restatement, everything else is the reference's code."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import seg_loss_oracle as O  # noqa: E402

ops = types.ModuleType("pytorch3d.ops")
ops.knn_points = lambda a, b, K: (*O.knn_points(a, b, K), None)
ops.knn_gather = O.knn_gather
sys.modules["pytorch3d"] = types.ModuleType("pytorch3d")
sys.modules["pytorch3d.ops"] = ops
spec = importlib.util.spec_from_file_location("ref_seg_loss", "/root/reference/utils/seg_loss.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

g = torch.Generator().manual_seed(5)
B, N, K = 2, 300, 4
pc = torch.rand(B, N, 3, generator=g) * 2 - 1
# two rigidly moving halves + noise, soft masks
ang = torch.tensor(0.3)
Rz = torch.tensor([[torch.cos(ang), -torch.sin(ang), 0], [torch.sin(ang), torch.cos(ang), 0], [0, 0, 1.0]])
flow = torch.where((pc[..., :1] > 0), pc @ Rz.t() - pc + 0.1, torch.full_like(pc, -0.05))
flow = flow + 0.01 * torch.randn(B, N, 3, generator=g)
mask = torch.softmax(3 * torch.randn(B, N, K, generator=g), dim=-1)
out = {"pc": pc, "flow": flow, "mask": mask}
R, t = ref.fit_motion_svd_batch(pc, pc + flow, mask[..., 0])
out["fit_R"], out["fit_t"] = R, t
R0, t0 = ref.fit_motion_svd_batch(pc, pc + flow)
out["fit_R_nomask"], out["fit_t_nomask"] = R0, t0
mk = mask.clone().requires_grad_(True)
ld, ptf = ref.dynamic_loss(pc, mk, flow)
ld.backward()
out["dynamic_loss"], out["dynamic_pc"], out["dynamic_grad_mask"] = ld.detach(), ptf.detach(), mk.grad.clone()
for name, kw in (("smooth_k4", dict(k=4, radius=0.01)), ("smooth_k16", dict(k=16, radius=0.1)),
                 ("smooth_k8_l2", dict(k=8, radius=0.05, loss_norm=2))):
    mk = mask.clone().requires_grad_(True)
    ls = ref.smooth_loss(pc, mk, **kw)
    ls.backward()
    out[name], out[name + "_grad_mask"] = ls.detach(), mk.grad.clone()
out["entropy"] = ref.entropy_loss(mask)
out["rank"] = ref.rank_loss(mask)
d8, i8 = O.knn_points(pc, pc, 8)
out["knn8_dist"], out["knn8_idx"] = d8, i8
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "segloss_small.npz"),
                    **{k: v.numpy() for k, v in out.items()})
print({k: tuple(v.shape) for k, v in out.items()})
