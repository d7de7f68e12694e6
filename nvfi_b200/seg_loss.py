"""Losses of the segmentation trainer (reference: utils/seg_loss.py:6-121, called at train_segm.py:186-197):
same names, arguments and return values.  SURVEY.md section 8 row f4 ("next"): not on the render path.

The reference needs pytorch3d for ``knn_points`` / ``knn_gather``; here the neighbour search is the CUDA
kernel ``nvfi_knn_points`` (csrc/knn.cu) and the rest is the reference's tensor algebra, with one change of
evaluation order that matters at the trainer's point counts: the masked covariance of
``fit_motion_svd_batch`` is formed as (pc1 * mask)^T pc2 instead of pc1^T diag(mask) pc2 — the reference
materialises the (B, N, N) diagonal matrix (utils/seg_loss.py:31: 29 GB for 8 objects x 30 000 points).
"""
from __future__ import annotations

import torch

from . import _lib as L


def knn_points(p1: torch.Tensor, p2: torch.Tensor, K: int = 1):
    """(dists, idx, None): squared distances (B, N1, K) ascending, indices (B, N1, K) into p2."""
    if not (p1.is_cuda and p2.is_cuda):
        raise RuntimeError("nvfi_b200: knn_points needs CUDA tensors (no CPU fallback)")
    q = p1.detach().contiguous().float()
    p = p2.detach().contiguous().float()
    B, N1, _ = q.shape
    N2 = p.shape[1]
    dist = torch.empty(B, N1, K, device=q.device, dtype=torch.float32)
    idx = torch.empty(B, N1, K, device=q.device, dtype=torch.int64)
    L.check(L.load().nvfi_knn_points(q.data_ptr(), p.data_ptr(), B, N1, N2, K, dist.data_ptr(), idx.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), "knn_points")
    return dist, idx, None


def knn_gather(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x (B, N, C), idx (B, M, K) -> (B, M, K, C); differentiable in x."""
    B, M, K = idx.shape
    C_ = x.shape[-1]
    return torch.gather(x, 1, idx.reshape(B, M * K, 1).expand(B, M * K, C_)).reshape(B, M, K, C_)


def fit_motion_svd_batch(pc1, pc2, mask=None):
    """Weighted Kabsch fit (utils/seg_loss.py:6-55): R (B, 3, 3), t (B, 3) with pc2 ~ R pc1 + t."""
    n_batch = pc1.shape[0]
    if mask is None:
        pc1_mean = pc1.mean(dim=1, keepdim=True)
        pc2_mean = pc2.mean(dim=1, keepdim=True)
    else:
        den = mask.sum(dim=1, keepdim=True)
        pc1_mean = (torch.einsum("bnd,bn->bd", pc1, mask) / den).unsqueeze(1)
        pc2_mean = (torch.einsum("bnd,bn->bd", pc2, mask) / den).unsqueeze(1)
    c1, c2 = pc1 - pc1_mean, pc2 - pc2_mean
    if mask is None:
        S = torch.bmm(c1.transpose(1, 2), c2)
    else:
        S = torch.bmm((c1 * mask.unsqueeze(-1)).transpose(1, 2), c2)
    valid = ~torch.isnan(S).any(dim=1).any(dim=1)
    R_base = torch.eye(3, device=pc1.device).unsqueeze(0).repeat(n_batch, 1, 1)
    t_base = torch.zeros((n_batch, 3), device=pc1.device)
    if valid.any():
        S = S[valid]
        u, s, vh = torch.linalg.svd(S, full_matrices=True)
        v = vh.transpose(1, 2)
        R = torch.bmm(v, u.transpose(1, 2))
        det = torch.det(R)
        diag = torch.ones_like(S[..., 0], requires_grad=False)
        diag[:, 2] = det
        R = v.bmm(torch.diag_embed(diag).bmm(u.transpose(1, 2)))
        m1, m2 = pc1_mean[valid], pc2_mean[valid]
        t = m2.squeeze(1) - torch.bmm(R, m1.transpose(1, 2)).squeeze(2)
        R_base[valid] = R
        t_base[valid] = t
    return R_base, t_base


def dynamic_loss(pc, mask, flow):
    """utils/seg_loss.py:58-85: per-object rigid fit of the flow, discrepancy of the mask-blended motion."""
    n_batch, n_point, n_object = mask.size()
    pc2 = pc + flow
    mask = mask.transpose(1, 2).reshape(n_batch * n_object, n_point)
    pc_rep = pc.unsqueeze(1).expand(n_batch, n_object, n_point, 3).reshape(n_batch * n_object, n_point, 3)
    pc2_rep = pc2.unsqueeze(1).expand(n_batch, n_object, n_point, 3).reshape(n_batch * n_object, n_point, 3)
    object_R, object_t = fit_motion_svd_batch(pc_rep, pc2_rep, mask)
    pc_transformed = torch.einsum("bij,bnj->bni", object_R, pc_rep) + object_t.unsqueeze(1)
    pc_transformed = pc_transformed.reshape(n_batch, n_object, n_point, 3).detach()
    mask = mask.reshape(n_batch, n_object, n_point).unsqueeze(-1)
    pc_transformed = (mask * pc_transformed).sum(1)
    loss = (pc_transformed - pc2).norm(p=2, dim=-1)
    return loss.mean(), pc_transformed


def smooth_loss(pc, mask, k=16, radius=0.1, loss_norm=1):
    """utils/seg_loss.py:78-90: neighbours further than `radius` (compared with the SQUARED distance, as the
    reference does) are replaced by the nearest one."""
    dist, idx, _ = knn_points(pc, pc, K=k)
    first = idx[:, :, 0].unsqueeze(2).repeat(1, 1, k)
    far = dist > radius
    idx[far] = first[far]
    nn_mask = knn_gather(mask, idx.detach())
    loss = (mask.unsqueeze(2) - nn_mask).norm(p=loss_norm, dim=-1)
    return loss.mean()


def entropy_loss(mask, epsilon=1e-5):
    """utils/seg_loss.py:93-102."""
    return (-(mask * torch.log(mask.clamp(epsilon)))).sum(-1).mean()


def rank_loss(mask):
    """utils/seg_loss.py:105-112."""
    return mask.norm(p="nuc", dim=(1, 2)).mean()
