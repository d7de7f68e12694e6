"""TEST INFRASTRUCTURE ONLY (never imported by nvfi_b200/): CPU restatement of the segmentation losses
(reference utils/seg_loss.py:6-121) for tests/test_seg_loss.py and tests/test_gpu_seg_loss.py.

Pinning: fit_motion_svd_batch / dynamic_loss / entropy_loss / rank_loss are checked against outputs of the
reference module itself (tests/golden/segloss_small.npz, made by tests/golden/make_golden_segloss.py, which
imports /root/reference/utils/seg_loss.py).  ``knn_points`` lives in pytorch3d (requirements.txt:
``pytorch3d``, un-vendored and absent from the image), so its published behaviour is restated here —
squared L2 distances, K smallest per query in ascending order — and smooth_loss is pinned to the
reference's own code running on top of that restatement."""
import torch


def knn_points(p1, p2, K):
    """pytorch3d.ops.knn_points(p1, p2, K=K)[:2]: (squared distances ascending, indices)."""
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    dist, idx = torch.topk(d, K, dim=-1, largest=False, sorted=True)
    return dist, idx


def knn_gather(x, idx):
    """pytorch3d.ops.knn_gather: x (B, N, C), idx (B, M, K) -> (B, M, K, C)."""
    B, M, K = idx.shape
    return torch.stack([x[b][idx[b]] for b in range(B)], 0)


def fit_motion_svd_batch(pc1, pc2, mask=None):
    """utils/seg_loss.py:6-55, statement for statement (including the (B, N, N) diag_embed)."""
    n_batch = pc1.shape[0]
    if mask is None:
        m1, m2 = pc1.mean(1, keepdim=True), pc2.mean(1, keepdim=True)
    else:
        m1 = (torch.einsum("bnd,bn->bd", pc1, mask) / mask.sum(1, keepdim=True)).unsqueeze(1)
        m2 = (torch.einsum("bnd,bn->bd", pc2, mask) / mask.sum(1, keepdim=True)).unsqueeze(1)
    c1, c2 = pc1 - m1, pc2 - m2
    S = torch.bmm(c1.transpose(1, 2), c2) if mask is None else c1.transpose(1, 2).bmm(torch.diag_embed(mask).bmm(c2))
    valid = ~torch.isnan(S).any(1).any(1)
    R_base = torch.eye(3).unsqueeze(0).repeat(n_batch, 1, 1)
    t_base = torch.zeros(n_batch, 3)
    if valid.any():
        S = S[valid]
        u, s, v = torch.svd(S, some=False, compute_uv=True)
        R = torch.bmm(v, u.transpose(1, 2))
        diag = torch.ones_like(S[..., 0])
        diag[:, 2] = torch.det(R)
        R = v.bmm(torch.diag_embed(diag).bmm(u.transpose(1, 2)))
        t = m2[valid].squeeze(1) - torch.bmm(R, m1[valid].transpose(1, 2)).squeeze(2)
        R_base[valid], t_base[valid] = R, t
    return R_base, t_base


def dynamic_loss(pc, mask, flow):
    """utils/seg_loss.py:58-85."""
    B, N, K = mask.shape
    pc2 = pc + flow
    m = mask.transpose(1, 2).reshape(B * K, N)
    pr = pc.unsqueeze(1).repeat(1, K, 1, 1).reshape(B * K, N, 3)
    p2r = pc2.unsqueeze(1).repeat(1, K, 1, 1).reshape(B * K, N, 3)
    R, t = fit_motion_svd_batch(pr, p2r, m)
    pt = (torch.einsum("bij,bnj->bni", R, pr) + t.unsqueeze(1)).reshape(B, K, N, 3).detach()
    pt = (m.reshape(B, K, N).unsqueeze(-1) * pt).sum(1)
    return (pt - pc2).norm(p=2, dim=-1).mean(), pt


def smooth_loss(pc, mask, k=16, radius=0.1, loss_norm=1):
    """utils/seg_loss.py:78-90."""
    dist, idx = knn_points(pc, pc, k)
    first = idx[:, :, 0].unsqueeze(2).repeat(1, 1, k)
    idx = idx.clone()
    idx[dist > radius] = first[dist > radius]
    return (mask.unsqueeze(2) - knn_gather(mask, idx)).norm(p=loss_norm, dim=-1).mean()


def entropy_loss(mask, epsilon=1e-5):
    """utils/seg_loss.py:93-102."""
    return (-(mask * torch.log(mask.clamp(epsilon)))).sum(-1).mean()


def rank_loss(mask):
    """utils/seg_loss.py:105-112."""
    return mask.norm(p="nuc", dim=(1, 2)).mean()
