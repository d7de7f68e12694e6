"""The north-star multi-GPU split on hardware: ONE frame ray-sharded over the GPUs
of the box (models/renderer.py:29-42's chunk loop spread over the ranks) with a single collective.
Needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on a
one-GPU box (the gloo tests in test_sharding.py cover the host logic there)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_frame_equals_single_gpu_frame():
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "mp", "sharded_frame_worker.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, p.stdout[-2000:]
    r = json.loads(line[-1][7:])
    print(r)
    assert r["bat_eval_bitwise"] and r["chessboard_eval_bitwise"] and r["all_ranks_bitwise"]
    assert 0.01 < r["bat_acc_mean"] < 0.99       # the frame is neither empty nor saturated
    assert r["train_same_none"]
    assert r["train_loss_rel"] < 1e-5
    assert r["train_grad_worst_rel"] < 1e-5      # same sums in another order (atomics, all-reduce tree)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_second_device_in_one_process():
    """the shared-memory opt-ins (cudaFuncSetAttribute) and SM counts are cached per device —
    a process that renders on cuda:0 and then on cuda:1 must get the same frame on both."""
    from nvfi_b200.scenes import build_scene, frame_rays
    o, d = frame_rays(64, 64, theta=30.0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            cfg, nv, _ = build_scene("bat", grid=(40, 40, 40), device=dev, step_ratio=1.5)
            f = nv.nvfi
            nv.requires_grad_(True)
            f.train()
            gen = torch.Generator().manual_seed(1)
            jit = torch.rand(o.shape[0], 1, generator=gen)
            rgb, depth, acc, w, _ = f.render_rays(0.33, o.to(dev), d.to(dev), white_bg=True, ray_chunk=1024, jitter=jit)
            rgb.sum().backward()                 # the tensor-core backward and the gather kernels on this device
            g = f.vel_net.weight_net[1].weight.grad
            loss = nv.get_vel_loss(2048, points=torch.rand(2048, 3, generator=gen).to(dev) * 1.2 - 0.6,
                                   t=torch.rand(2048, 1, generator=gen).to(dev) * 0.75)
            torch.cuda.synchronize()
            outs.append((rgb.detach().cpu(), depth.detach().cpu(), g.cpu(), float(loss)))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert float((outs[0][2] - outs[1][2]).norm() / outs[0][2].norm()) < 1e-5      # atomics: summation order
    assert abs(outs[0][3] - outs[1][3]) < 1e-6 * max(1.0, abs(outs[0][3]))
