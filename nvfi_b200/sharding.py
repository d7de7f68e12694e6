"""Ray sharding across the GPUs of one box.

The reference is single-process (no torch.distributed anywhere); rays are independent, so
the B200 build adds plain data parallelism: one process per GPU, parameters replicated, a
contiguous block of rays per rank and ONE collective per step —

  * eval:  every rank renders its block of the frame, then one all-gather of the
           (rgb, depth, acc[, mask]) slabs (``gather_frame``);
  * train: every rank back-propagates its block, then one all-reduce(sum) over a single
           flat FP32 buffer holding every parameter gradient plus the loss numerator and
           the ray count (``allreduce_grads``), so the result equals the single-GPU
           mean-reduced loss / gradient.

The only cross-ray coupling in the reference is the chunk-global inside-box predicate of
``sample_ray`` (models/tensorf_base.py:294): shards are cut at multiples of the reference
chunk size (``ray_chunk``), so every reference chunk lives on exactly one rank and the
predicate is evaluated on the same set of rays as in the single-process run.

Works with any initialised ``torch.distributed`` backend (NCCL on the GPUs; the CPU tests
use gloo).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_rays: int, rank: int, world_size: int, ray_chunk: int = 2048) -> Tuple[int, int]:
    """[begin, end) of the rays rank ``rank`` renders: contiguous, chunk-aligned, sizes
    differing by at most one chunk, the union covering [0, n_rays) exactly once."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if n_rays < 0 or ray_chunk <= 0:
        raise ValueError("n_rays must be >= 0 and ray_chunk > 0")
    n_chunks = (n_rays + ray_chunk - 1) // ray_chunk
    base, extra = divmod(n_chunks, world_size)
    c0 = rank * base + min(rank, extra)
    c1 = c0 + base + (1 if rank < extra else 0)
    return min(c0 * ray_chunk, n_rays), min(c1 * ray_chunk, n_rays)


def shard_rays(tensors: Sequence[Optional[torch.Tensor]], rank: int, world_size: int,
               ray_chunk: int = 2048) -> List[Optional[torch.Tensor]]:
    """Slice per-ray tensors (first dim = rays) to this rank's block."""
    n = next(t.shape[0] for t in tensors if t is not None)
    b, e = shard_bounds(n, rank, world_size, ray_chunk)
    return [None if t is None else t[b:e] for t in tensors]


def gather_frame(parts: Sequence[torch.Tensor], n_rays: int, ray_chunk: int = 2048,
                 group=None) -> List[torch.Tensor]:
    """Eval: all-gather the per-rank slabs of per-ray outputs into full-frame tensors.

    ``parts`` are this rank's (n_local, c_i) / (n_local,) outputs.  They are packed into one
    (n_local, sum c_i) buffer so that a frame costs exactly one collective."""
    rank, ws = world()
    if ws == 1:
        return [p for p in parts]
    widths = [1 if p.dim() == 1 else p.shape[1] for p in parts]
    n_local = parts[0].shape[0]
    packed = torch.cat([p.reshape(n_local, -1).float() for p in parts], dim=1).contiguous()
    sizes = [shard_bounds(n_rays, r, ws, ray_chunk) for r in range(ws)]
    max_n = max(e - b for b, e in sizes)
    if n_local < max_n:   # all_gather wants equal shapes: pad the short slabs
        pad = torch.zeros(max_n - n_local, packed.shape[1], device=packed.device, dtype=packed.dtype)
        packed = torch.cat([packed, pad], 0)
    out = torch.empty(ws * max_n, packed.shape[1], device=packed.device, dtype=packed.dtype)
    dist.all_gather_into_tensor(out, packed, group=group)
    full = torch.cat([out[r * max_n: r * max_n + (e - b)] for r, (b, e) in enumerate(sizes)], 0)
    res, c = [], 0
    for p, w in zip(parts, widths):
        col = full[:, c:c + w]
        res.append(col.reshape(n_rays) if p.dim() == 1 else col.reshape(n_rays, w))
        c += w
    return res


def shard_index(n_rays: int, rank: int, world_size: int, ray_chunk: int = 2048) -> torch.Tensor:
    """Ray indices (1-D int64, ascending) of rank ``rank`` under the INTERLEAVED partition: reference
    chunk c goes to rank c % world_size.  Every reference chunk still lives on exactly one rank (the
    chunk-global predicate of sample_ray is unchanged), but the ranks' loads are balanced: contiguous
    blocks of an image give the ranks that own the empty top and bottom rows far less work than the
    ranks that own the object (strong scaling of one frame)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if n_rays < 0 or ray_chunk <= 0:
        raise ValueError("n_rays must be >= 0 and ray_chunk > 0")
    n_chunks = (n_rays + ray_chunk - 1) // ray_chunk
    if rank >= n_chunks:
        return torch.zeros(0, dtype=torch.int64)
    chunks = torch.arange(rank, n_chunks, world_size, dtype=torch.int64)
    idx = (chunks[:, None] * ray_chunk + torch.arange(ray_chunk, dtype=torch.int64)[None, :]).reshape(-1)
    return idx[idx < n_rays]


def gather_frame_interleaved(parts: Sequence[torch.Tensor], n_rays: int, ray_chunk: int = 2048,
                             group=None) -> List[torch.Tensor]:
    """Eval with the interleaved partition (``shard_index``): ONE all-gather of the packed per-ray
    outputs, then each rank's rows are written to their places in the full frame."""
    rank, ws = world()
    if ws == 1:
        return [p for p in parts]
    widths = [1 if p.dim() == 1 else p.shape[1] for p in parts]
    n_local = parts[0].shape[0]
    dev = parts[0].device
    packed = torch.cat([p.reshape(n_local, -1).float() for p in parts], dim=1).contiguous()
    index = [shard_index(n_rays, r, ws, ray_chunk) for r in range(ws)]
    max_n = max(int(i.numel()) for i in index)
    if n_local < max_n:
        pad = torch.zeros(max_n - n_local, packed.shape[1], device=dev, dtype=packed.dtype)
        packed = torch.cat([packed, pad], 0)
    out = torch.empty(ws * max_n, packed.shape[1], device=dev, dtype=packed.dtype)
    dist.all_gather_into_tensor(out, packed, group=group)
    full = torch.empty(n_rays, packed.shape[1], device=dev, dtype=packed.dtype)
    for r, i in enumerate(index):
        full[i.to(dev)] = out[r * max_n: r * max_n + i.numel()]
    res, c = [], 0
    for p, w in zip(parts, widths):
        col = full[:, c:c + w]
        res.append(col.reshape(n_rays) if p.dim() == 1 else col.reshape(n_rays, w))
        c += w
    return res


_PENDING = None      # (summed has-grad flags of the last in-place all-reduce, world size): verified lazily


def check_pending() -> None:
    """The in-place all-reduce requires that a parameter has a gradient on every rank or on none.  The summed
    flags of a step are checked here — at the start of the next in-place step, or when the caller asks —
    instead of with a device->host sync at the end of every step."""
    global _PENDING
    if _PENDING is None:
        return
    flags, ws = _PENDING
    _PENDING = None
    bad = [i for i, f in enumerate(flags.tolist()) if f not in (0.0, ws)]
    if bad:
        raise RuntimeError(f"nvfi_b200.sharding: parameters {bad} had a gradient on some ranks only in the previous "
                           "in-place all-reduce; call allreduce_grads without flat= for such steps")


def _allreduce_in_place(plist, extras, group, average, flat, ws) -> Optional[torch.Tensor]:
    """Zero-copy variant: every existing gradient is a view of ``flat[0]`` (engine.render_backward hands out
    views of one buffer), so that buffer — with the has-grad flags and the extras written into its tail — is
    all-reduced in place.  Returns None when the precondition does not hold (the caller then copies)."""
    buf, used = flat
    n_p = len(plist)
    n_e = extras.numel() if extras is not None else 0
    if used + n_p + n_e > buf.numel():
        return None
    base = buf.untyped_storage().data_ptr()
    lo, hi = buf.data_ptr(), buf.data_ptr() + used * 4
    for p in plist:
        g = p.grad
        if g is None:
            continue
        if (g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base
                or not (lo <= g.data_ptr() and g.data_ptr() + g.numel() * 4 <= hi)):
            return None
    flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in plist], device=buf.device)
    buf[used:used + n_p].copy_(flags)
    if n_e:
        buf[used + n_p:used + n_p + n_e].copy_(extras.reshape(-1).float())
    view = buf[:used + n_p + n_e]
    check_pending()          # the previous step's flags, one step late: no host sync inside a step
    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
    global _PENDING
    _PENDING = (buf[used:used + n_p].clone(), float(ws))
    if average:
        buf[:used].mul_(1.0 / ws)
    return buf[used + n_p:used + n_p + n_e].clone() if n_e else torch.zeros(0, device=buf.device)


def allreduce_grads(params: Iterable[torch.nn.Parameter], extras: Optional[torch.Tensor] = None,
                    group=None, average: bool = False, flat=None) -> Optional[torch.Tensor]:
    """Train: ONE all-reduce(sum) over [all param grads || has-grad flags || extras] in a flat FP32 buffer.

    ``extras`` (1-D tensor) carries scalars that must be summed across ranks as well (loss
    numerators, ray / point counts); the reduced copy is returned.  The buffer has a slot for EVERY
    parameter that requires grad, so its length is the same on every rank whatever each rank's rays
    hit; a parameter whose grad is None contributes zeros and a 0 flag.  After the reduction a
    parameter receives a gradient iff at least one rank had one (e.g. ``get_vel_loss`` returns python
    0.0 on a rank with no occupied point and leaves ``a_weight_net`` without grads there, but not on
    the others); a parameter without a gradient on every rank stays None, as in the single-process
    run.  With ``average`` the gradients are divided by the world size afterwards (use it when every
    rank's loss is already a mean over its own, equally sized block).

    ``flat`` = ``engine.last_grad_flat()``: when every gradient of the step came out of ONE
    ``render_backward`` call they are views of one buffer, which is then reduced in place — no flattening
    copies of the 38 MB of gradients before and after the collective (strong scaling at 8 GPUs: the
    copies were 1.2 ms of a 50 ms step).  Falls back to the copying path when a gradient lives elsewhere."""
    rank, ws = world()
    plist = [p for p in params if p.requires_grad]
    if ws == 1:
        return extras
    if not plist and extras is None:
        return None
    if flat is not None:
        res = _allreduce_in_place(plist, extras, group, average, flat, ws)
        if res is not None:
            return res if extras is not None else None
    dev = plist[0].device if plist else extras.device
    n_p = len(plist)
    n_e = extras.numel() if extras is not None else 0
    numel = sum(p.numel() for p in plist) + n_p + n_e
    flat = torch.zeros(numel, device=dev, dtype=torch.float32)
    off = 0
    for p in plist:
        n = p.numel()
        if p.grad is not None:
            flat[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    flags_off = off
    flat[off:off + n_p].copy_(torch.tensor([0.0 if p.grad is None else 1.0 for p in plist], device=dev))
    off += n_p
    if extras is not None:
        flat[off:].copy_(extras.reshape(-1).float())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    has = (flat[flags_off:flags_off + n_p] > 0).tolist()
    scale = 1.0 / ws if average else 1.0
    off = 0
    for p, h in zip(plist, has):
        n = p.numel()
        if h:
            g = flat[off:off + n].view_as(p)
            if average:
                g = g * scale
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
        off += n
    return flat[flags_off + n_p:].clone() if extras is not None else None
