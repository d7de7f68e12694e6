"""Worker of tests/test_gpu_multi.py (launched under torch.distributed.run, one rank per GPU, NCCL):
one frame cut with sharding.shard_index, rendered by the ranks and re-assembled with ONE all-gather, must
equal the single-GPU frame BITWISE; the all-reduced train gradients must equal the single-GPU gradients
(up to the order of the floating-point reductions)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from nvfi_b200 import sharding
    from nvfi_b200.scenes import build_scene, frame_rays
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {}
    Hh = Ww = 96
    chunk = 512
    for scene, t, white in (("bat", 0.33, True), ("chessboard", 1.0, False)):
        zs = 3.0 if scene == "chessboard" else 0.0
        cfg, nv, _ = build_scene(scene, grid=(40, 40, 40), device=dev, step_ratio=1.5)
        f = nv.nvfi
        o, d = frame_rays(Hh, Ww, theta=30.0, z_shift=zs)
        n = o.shape[0]
        idx = sharding.shard_index(n, rank, world, chunk)
        f.eval()
        with torch.no_grad():
            part = f.render_rays(t, o[idx].cuda(), d[idx].cuda(), white_bg=white, ray_chunk=chunk)
            full = sharding.gather_frame_interleaved([part[0], part[1], part[2]], n, chunk)
            one = f.render_rays(t, o.cuda(), d.cuda(), white_bg=white, ray_chunk=chunk)
        res[f"{scene}_eval_bitwise"] = all(bool(torch.equal(a, b)) for a, b in zip(full, one[:3]))
        res[f"{scene}_acc_mean"] = float(one[2].mean())
        if scene != "bat":
            continue
        # train: gradients of the frame's MSE, sharded + all-reduce(sum) vs single GPU
        gen = torch.Generator().manual_seed(3)
        target = torch.rand(n, 3, generator=gen)
        jitter = torch.rand(n, 1, generator=gen)
        nv.requires_grad_(True)
        f.train()
        params = [p for p in nv.parameters()]

        def grads(sel):
            for p in params:
                p.grad = None
            rgb = f.render_rays(t, o[sel].cuda(), d[sel].cuda(), white_bg=white, ray_chunk=chunk,
                                jitter=jitter[sel].cuda())[0]
            loss = ((rgb - target[sel].cuda()) ** 2).sum() / float(3 * n)
            loss.backward()
            return loss.detach()
        loss_l = grads(idx)
        loss_sum = sharding.allreduce_grads(params, extras=loss_l.reshape(1))
        sharded = [None if p.grad is None else p.grad.clone() for p in params]
        loss_1 = grads(torch.arange(n))
        worst = 0.0
        same_none = True
        for g, p in zip(sharded, params):
            if (g is None) != (p.grad is None):
                same_none = False
                continue
            if g is None:
                continue
            den = float(p.grad.norm())
            if den > 0:
                worst = max(worst, float((g - p.grad).norm()) / den)
        res["train_grad_worst_rel"] = worst
        res["train_same_none"] = same_none
        res["train_loss_rel"] = abs(float(loss_sum[0]) - float(loss_1)) / max(abs(float(loss_1)), 1e-12)
    flags = torch.tensor([1.0 if all(v for k, v in res.items() if k.endswith("bitwise")) else 0.0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res["all_ranks_bitwise"] = bool(flags.item() == 1.0)
    if rank == 0:
        print("RESULT " + json.dumps(res), flush=True)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
