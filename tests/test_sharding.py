"""Ray sharding: partition arithmetic and the two collectives, on CPU with
gloo and world_size 2 (the GPU path uses the same code over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nvfi_b200 import sharding


@pytest.mark.parametrize("n,ws,chunk", [(640000, 8, 2048), (640000, 3, 2048), (4096, 2, 2048), (100, 4, 7),
                                        (0, 2, 2048), (2047, 2, 2048), (5, 8, 1)])
def test_shard_bounds_partition(n, ws, chunk):
    spans = [sharding.shard_bounds(n, r, ws, chunk) for r in range(ws)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
        assert e0 == b1 and b0 <= e0
    for b, e in spans:      # chunk-aligned starts: every reference chunk lives on one rank
        assert b % chunk == 0 or b == n
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) < 2 * chunk     # one chunk of imbalance + the ragged tail


def test_shard_bounds_rejects_bad_args():
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 0, 1, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)
        n, chunk = 1000, 64
        rgb = torch.rand(n, 3)
        depth = torch.rand(n)
        b, e = sharding.shard_bounds(n, rank, ws, chunk)
        full = sharding.gather_frame([rgb[b:e], depth[b:e]], n, chunk)
        ok_gather = torch.equal(full[0], rgb) and torch.equal(full[1], depth)
        idx = sharding.shard_index(n, rank, ws, chunk)
        full2 = sharding.gather_frame_interleaved([rgb[idx], depth[idx]], n, chunk)
        ok_gather = ok_gather and torch.equal(full2[0], rgb) and torch.equal(full2[1], depth)

        # gradient all-reduce: sum over ranks of per-shard gradients == single-process gradient
        w = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
        unused = torch.nn.Parameter(torch.ones(2))
        x = torch.rand(n, 3)
        loss_num = ((x[b:e] @ w.t()) ** 2).sum()
        loss_num.backward()
        extras = sharding.allreduce_grads([w, unused], extras=torch.tensor([float(loss_num), float(e - b)]))
        w_ref = torch.nn.Parameter(w.detach().clone())
        ref = ((x @ w_ref.t()) ** 2).sum()
        ref.backward()
        ok_grad = torch.allclose(w.grad, w_ref.grad, rtol=1e-5, atol=1e-5) and unused.grad is None
        # a parameter that has a gradient on ONE rank only: same buffer length on
        # every rank, and the gradient arrives everywhere
        lone = torch.nn.Parameter(torch.zeros(3))
        if rank == 1:
            lone.grad = torch.tensor([1.0, 2.0, 3.0])
        sharding.allreduce_grads([w, lone, unused])
        ok_grad = ok_grad and lone.grad is not None and torch.equal(lone.grad, torch.tensor([1.0, 2.0, 3.0])) \
            and unused.grad is None
        ok_extra = abs(float(extras[0]) - float(ref)) < 1e-3 * float(ref) and int(extras[1]) == n
        # zero-copy variant: the gradients are views of ONE buffer with a tail (engine.render_backward's layout)
        buf = torch.zeros(64 + 32)
        pa, pb, pn = (torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(5)),
                      torch.nn.Parameter(torch.zeros(4)))
        pa.grad, pb.grad = buf[0:6].view(2, 3), buf[32:37]
        pa.grad.fill_(float(rank + 1))
        pb.grad.fill_(10.0 * (rank + 1))
        ex = sharding.allreduce_grads([pa, pb, pn], extras=torch.tensor([float(rank)]), flat=(buf, 64))
        ok_flat = (pa.grad.data_ptr() == buf.data_ptr() and torch.equal(pa.grad, torch.full((2, 3), 3.0))
                   and torch.equal(pb.grad, torch.full((5,), 30.0)) and pn.grad is None and float(ex[0]) == 1.0)
        sharding.check_pending()         # flags of the in-place step agree on both ranks
        # a gradient outside the buffer -> silent fall back to the copying path, same result
        pc = torch.nn.Parameter(torch.zeros(3))
        pc.grad = torch.full((3,), float(rank + 1))
        sharding.allreduce_grads([pa, pc], flat=(buf, 64))
        ok_flat = ok_flat and torch.equal(pc.grad, torch.full((3,), 3.0)) and torch.equal(pa.grad, torch.full((2, 3), 6.0))
        # ... and a gradient that exists on one rank only is reported (one step late, by check_pending)
        pd = torch.nn.Parameter(torch.zeros(4))
        if rank == 1:
            pd.grad = buf[40:44]
        sharding.allreduce_grads([pa, pd], flat=(buf, 64))
        try:
            sharding.check_pending()
            ok_flat = False
        except RuntimeError:
            pass
        ok_grad = ok_grad and ok_flat
        q.put((rank, ok_gather, ok_grad, ok_extra))
    finally:
        dist.destroy_process_group()


def test_collectives_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res


@pytest.mark.parametrize("n,ws,chunk", [(640000, 8, 2048), (640000, 3, 2048), (100, 4, 7), (5, 8, 1)])
def test_shard_index_partition(n, ws, chunk):
    parts = [sharding.shard_index(n, r, ws, chunk) for r in range(ws)]
    allidx = torch.cat(parts).sort().values
    assert torch.equal(allidx, torch.arange(n))                   # every ray exactly once
    for r, p in enumerate(parts):                                 # whole reference chunks, chunk c on rank c % ws
        assert bool(((p // chunk) % ws == r).all())
    sizes = [p.numel() for p in parts]
    assert max(sizes) - min(sizes) <= chunk


def test_single_process_passthrough():
    a = torch.rand(5, 3)
    out = sharding.gather_frame([a], 5)
    assert out[0] is a
    assert sharding.world() == (0, 1)
