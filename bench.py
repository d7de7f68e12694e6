#!/usr/bin/env python
"""bench.py — rays/s of the NVFi train step (render forward + MSE + hand-written backward)
on the BASELINE.json workload: bat.yaml, ONE 800x800 frame (640 000 rays) per step,
192 samples/ray, final 199^3 grid, K = 16 keyframes, non-keyframe time (every valid sample is
advected by one RK2 step = 2 velocity-MLP evaluations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  Keys (see DESIGN.md "Measurement"):
  value         whole-job rays/s, inputs resident in HBM, CUDA events, max over ranks
  e2e           same metric through the public API (models.Renderer.render + loss.backward +
                loss.item()) with HOST ray / target buffers: H2D and D2H inside the timed region
  roofline      dominant kernel: algorithmic FLOPs (or bytes) per launch / its CUDA-event time,
                measured live through the library's per-launch event hook (no profiler)
  kernels       the same for every kernel of the step (+ L2 figures for the gather kernels)
  cpu_baseline  the UNMODIFIED reference (baseline/_ref; the oracle port if it is absent) on a bounded,
                frame-wide sample of the same workload, all host threads
  --impl reference  times only that CPU leg
  pde_262144, chessboard_eval_t1.0, fan_mask_render   BASELINE.json configs[2..4] (extra keys)
Multi-GPU: STRONG scaling — the same frame for every world size, the reference's 2 048-ray chunks dealt
round-robin to the ranks, ONE in-place all-reduce of the gradient buffer per step; `weak` (every rank a
whole frame) and `strong_breakdown` are extra keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rays/sec (train step, 800x800, 192 samples/ray)"
UNIT = "rays/s"
H = W = 800
GRID = (199, 199, 199)
STEP_RATIO = 1.79          # update_stepSize then computes nSamples = 192 spanning the box
T_RENDER = 0.33            # non-keyframe time: base 0.35, one RK2 step backwards
RAY_CHUNK = 2048           # the reference's chunk size (renderer.n_rays)
VEL_EVAL_FLOP = 2 * (28 * 128 + 4 * 128 * 128 + 128 * 6)     # 139 776
APP_EVAL_FLOP = 2 * (110 * 128 + 128 * 128 + 128 * 3) + 2 * 48 * 32   # 61 696 + 3 072
DENSITY_BYTES = 6 * 4 * 24 * 4    # 2 304 B per valid sample
APP_BYTES = 6 * 4 * 48 * 4        # 4 608 B per appearance sample


_REAL_STDOUT = None


def capture_stdout():
    """Everything that libraries print to stdout while the bench runs (the reference's module dumps, NCCL's
    version banner) goes to stderr; stdout carries exactly ONE line, written by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def env_world():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def workload_config(world):
    return {"workload": "bat.yaml, ONE 800x800 frame per step (sharded over the GPUs), 192 samples/ray "
                        "(step_ratio 1.79 @199^3), K=16, t=0.33 (1 RK2 step), train step = render fwd + MSE + bwd",
            "rays_per_step": H * W, "samples_per_ray": 192, "grid": list(GRID),
            "ray_chunk": RAY_CHUNK, "parallelism": f"ray-sharded dp{world}",
            "l2": "per-sample buffers (6 GB/step) exceed L2; the 37 MB factor planes are L2-resident by design"}


# ------------------------------------------------------------------------------------------
# CPU leg (oracle port of the reference algorithm)
# ------------------------------------------------------------------------------------------
class CpuLeg:
    """The reference's CPU path on the host: train render fwd + MSE + backward on 2 048-ray chunks
    of the same frame (the reference's own chunk size), all host threads.

    kind "reference": the UNMODIFIED reference (models.Renderer.render + autograd) imported from
    baseline/_ref (tools/install_reference.py; git-ignored, shipped to the GPU box);
    kind "port": the oracle restatement (oracle/nvfi_oracle.py), used when baseline/_ref is absent.
    Chunks are taken at a stride across the WHOLE frame (the frame has 313 chunks; stride 37 is
    coprime), so the sample sees empty borders and the cube in the frame's proportions; the valid-
    sample fraction of the sample is reported next to the frame's."""

    STRIDE = 37

    def __init__(self):
        from nvfi_b200 import configs, synth
        from nvfi_b200.scenes import frame_rays
        from oracle import reference_loader

        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        cfg = configs.get_config("bat", step_ratio=STEP_RATIO)
        K = int(cfg.nvfi.num_keyframes)
        sd = synth.synth_state(cfg, list(GRID), K, seed=233)
        ref_path = reference_loader.find_reference(allow_checkout=False)
        if ref_path is not None:
            self.kind = "reference"
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):     # the reference prints its modules while it builds them
                self.ref_models, _, self.nv = reference_loader.build_reference(ref_path, cfg, GRID, K, sd)
            self.nv.requires_grad_(True)
            self.renderer = self.ref_models.Renderer(self.nv, 0, 0, RAY_CHUNK)
            self.field = self.nv.nvfi
        else:
            from oracle.scene_io import scene_from_state
            self.kind = "port"
            self.sc = scene_from_state(cfg, list(GRID), K, sd, requires_grad=True)
        self.o, self.d = frame_rays(H, W, theta=30.0)
        self.n_chunks = (self.o.shape[0] + RAY_CHUNK - 1) // RAY_CHUNK
        self.gen = torch.Generator().manual_seed(7)
        self.next_chunk = 0
        self.valid = 0
        self.samples = 0

    def _valid_fraction(self, oo, dd):
        """Share of in-box samples of a chunk (the reference's own sampler, eval spacing)."""
        from oracle import nvfi_oracle as O
        with torch.no_grad():
            if self.kind == "reference":
                was = self.field.training
                self.field.eval()
                valid = self.field.sample_ray(oo, dd, N_samples=-1)[2]
                self.field.train(was)
            else:
                valid = O.sample_ray(self.sc, oo, dd, None)[2]
        return int(valid.sum()), int(valid.numel())

    def run(self, budget_s: float, chunks_cap: int):
        """Returns (seconds, rays, chunks) for up to `chunks_cap` chunks or `budget_s` seconds."""
        done, t_total, k = 0, 0.0, 0
        while k < chunks_cap and (k < 1 or t_total < budget_s):
            c = (self.next_chunk * self.STRIDE + 5) % self.n_chunks
            sl = slice(c * RAY_CHUNK, min((c + 1) * RAY_CHUNK, self.o.shape[0]))
            oo, dd = self.o[sl], self.d[sl]
            target = torch.rand(oo.shape[0], 3, generator=self.gen)
            if self.kind == "reference":
                for p in self.nv.parameters():
                    p.grad = None
                t0 = time.perf_counter()
                out = self.renderer.render(T_RENDER, self.ref_models.Ray(oo, dd, 0, 0), white_background=True,
                                           mode="train")
                loss = torch.nn.functional.mse_loss(out[0], target)
                loss.backward()
                t_total += time.perf_counter() - t0
            else:
                from oracle import nvfi_oracle as O
                jit = torch.rand(oo.shape[0], 1, generator=self.gen)
                for p in self.sc.parameters():
                    p.grad = None
                t0 = time.perf_counter()
                out = O.render_chunk(self.sc, T_RENDER, oo, dd, white_bg=True, training=True, jitter=jit)
                loss = torch.nn.functional.mse_loss(out[0], target)
                loss.backward()
                t_total += time.perf_counter() - t0
            try:
                v, tot = self._valid_fraction(oo, dd)
                self.valid += v
                self.samples += tot
            except Exception:
                pass
            done += oo.shape[0]
            k += 1
            self.next_chunk += 1
        return t_total, done, k

    def describe(self, k, frame_valid_fraction=None):
        s = (f"{k} chunks x {RAY_CHUNK} rays at stride {self.STRIDE} across the {self.n_chunks} chunks of the frame, "
             f"fwd+MSE+bwd, t={T_RENDER}, "
             + ("unmodified reference (baseline/_ref) on CPU" if self.kind == "reference" else "oracle port on CPU"))
        if self.samples:
            s += f"; valid-sample fraction of the sample {self.valid / self.samples:.3f}"
            if frame_valid_fraction is not None:
                s += f" (whole frame {frame_valid_fraction:.3f})"
        return s


def run_reference(args):
    rank, _, world = env_world()
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 8)), min(max(0, args.warmup), 1)
    leg = CpuLeg()
    for _ in range(warm):
        leg.run(0.0, 1)
    t_sum, r_sum, k_sum = 0.0, 0, 0
    for _ in range(steps):      # each step = a bounded sample of the frame: 2 chunks
        tt, rr, kk = leg.run(1e9, 2)
        t_sum += tt
        r_sum += rr
        k_sum += kk
    value = r_sum / t_sum
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t_sum / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": leg.cores, "kind": leg.kind,
                             "sample": leg.describe(k_sum)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.path = tempfile.mktemp(prefix="nvfi_clocks_", suffix=".csv")
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            ident = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        except Exception:
            ident = str(device_index)
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ident, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) >= 7:
                    rows.append(p)
            os.unlink(self.path)
        except Exception:
            pass
        if not rows:
            return out
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        out.update(sm_mhz=sm[len(sm) // 2] if sm else None,
                   sm_max_mhz=float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None,
                   power_w_max=max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None),
                   reasons=reasons, samples=len(rows))
        return out


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
                "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def measure_live_peaks(dev):
    """Dense-GEMM and L2 figures measured on this box, right now (extra keys; the roofline `peak` itself
    stays MEASURED_PEAKS.json's): TF32 / FP16 8192^3 torch.matmul (best of 5) and the bandwidth of a copy
    between two 24 MB buffers that stay in the 126 MB L2 (read + write bytes)."""
    out = {}
    try:
        n = 8192
        for name, dt, tf32 in (("tf32_tflops", torch.float32, True), ("fp16_tflops", torch.float16, False)):
            a = torch.randn(n, n, device=dev, dtype=dt)
            b = torch.randn(n, n, device=dev, dtype=dt)
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            best = 0.0
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            torch.backends.cuda.matmul.allow_tf32 = prev
            out[name] = best
            del a, b
        x = torch.empty(24 << 20, device=dev, dtype=torch.uint8)
        y = torch.empty_like(x)
        for _ in range(5):
            y.copy_(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            y.copy_(x)
        e1.record()
        torch.cuda.synchronize()
        out["l2_copy_gbs"] = 50 * 2 * x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
        out["how"] = "torch.matmul 8192^3 (allow_tf32 / fp16), best of 5; b.copy_(a) on 24 MB buffers x50 (L2-resident)"
    except Exception as exc:      # never lose the bench line over an extra
        out["error"] = str(exc)[:200]
    return out


def run_gpu(args):
    import torch.distributed as dist
    from nvfi_b200 import _lib, engine, sharding
    from nvfi_b200 import models as M
    from nvfi_b200.scenes import build_scene, frame_rays

    rank, local_rank, world = env_world()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the nvfi_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()     # fail loudly when the CUDA library is missing

    cfg, nv, _ = build_scene("bat", grid=GRID, device=dev, step_ratio=STEP_RATIO)
    field = nv.nvfi
    assert field.nSamples == 192, field.nSamples
    nv.requires_grad_(True)
    renderer = M.Renderer(nv, 0, 0, RAY_CHUNK)
    # ONE 800x800 frame per step, identical for every world size (same camera, same jitter, same target):
    # rank r renders the reference chunks c with c % world == r (sharding.shard_index: every 2 048-ray chunk
    # of models/renderer.py:29-42 lives on exactly one rank, and the empty borders and the object are
    # spread evenly over the ranks).
    o_f, d_f = frame_rays(H, W, theta=30.0)
    if args.rows != H:      # profiling aid: a horizontal band through the middle of the frame
        r0 = (H - args.rows) // 2
        o_f, d_f = o_f[r0 * W:(r0 + args.rows) * W].contiguous(), d_f[r0 * W:(r0 + args.rows) * W].contiguous()
    n_frame = o_f.shape[0]
    gen = torch.Generator().manual_seed(1000)
    target_f = torch.rand(n_frame, 3, generator=gen)
    jitter_f = torch.rand(n_frame, 1, generator=gen)
    params = [p for p in nv.parameters()]

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    class Job:
        """The train step on a set of rays of the frame (host-pinned and device-resident copies)."""

        def __init__(self, idx):
            sel = (lambda x: x) if idx is None else (lambda x: x[idx])
            self.o_h, self.d_h, self.target_h, self.jitter_h = (
                sel(x).contiguous().pin_memory() for x in (o_f, d_f, target_f, jitter_f))
            self.o_d, self.d_d, self.target_d, self.jitter_d = (
                x.to(dev) for x in (self.o_h, self.d_h, self.target_h, self.jitter_h))
            self.n = self.o_h.shape[0]

        def loss_of(self, rgb, target):
            # the frame's MSE: every rank contributes its rays' squared errors over the FRAME's element count,
            # so the all-reduce(sum) of the gradients gives exactly the single-GPU gradient
            return ((rgb - target) ** 2).sum() / float(3 * n_frame)

        def resident(self, collective=True):
            """Hot path with inputs resident in HBM."""
            for p in params:
                p.grad = None
            field.train()
            rgb, depth, acc, w, _ = field.render_rays(T_RENDER, self.o_d, self.d_d, white_bg=True,
                                                      ray_chunk=RAY_CHUNK, jitter=self.jitter_d)
            loss = self.loss_of(rgb, self.target_d)
            loss.backward()
            if world > 1 and collective:
                loss = sharding.allreduce_grads(params, extras=loss.detach().reshape(1), flat=engine.last_grad_flat())
            return loss

        def e2e(self):
            """Public API, host buffers: Ray.to(device) + Renderer.render(mode='train') + MSE + backward
            + loss.item() (what one iteration of train_nvfi.py does around the render, :156-164, :241-252)."""
            for p in params:
                p.grad = None
            rays = M.Ray(self.o_h, self.d_h, cfg.dataset.near, cfg.dataset.far).to(dev, non_blocking=True)
            tgt = self.target_h.to(dev, non_blocking=True)
            rgb, depth, acc, w, _ = renderer.render(T_RENDER, rays, white_background=True, mode="train")
            loss = self.loss_of(rgb, tgt)
            loss.backward()
            if world > 1:
                loss = sharding.allreduce_grads(params, extras=loss.detach().reshape(1), flat=engine.last_grad_flat())
            return float(loss.item()), rays

    # ---- headline: STRONG scaling, one frame per step sharded over the ranks
    job = Job(sharding.shard_index(n_frame, rank, world, RAY_CHUNK) if world > 1 else None)
    K, Wm = max(1, args.steps), max(3, args.warmup)
    for _ in range(Wm):
        job.resident()
    launches0 = _lib.launch_count()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ms_total = timed(job.resident, K)
    clk = clocks.stop() if clocks else None
    launches = (_lib.launch_count() - launches0)
    ms_step = ms_total / K
    value = n_frame / (ms_step * 1e-3)

    # ---- e2e through the public API with host buffers
    job.e2e()
    ms_e2e = timed(job.e2e, K) / K
    _, rays_obj = job.e2e()
    h2d = sum(b.numel() * b.element_size() for b in rays_obj.buffers()) + job.target_h.numel() * 4
    e2e = {"value": n_frame / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": 4 * world,
           "api": "models.Ray.to(device) + models.Renderer.render(mode='train') + mse + backward + "
                  "[all-reduce] + loss.item(); bytes are the whole job's (all ranks)"}
    del rays_obj

    sharding.check_pending()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world), "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clk}

    # ---- the same step with every in-box sample evaluated, as the reference does (early ray termination off)
    prev_et = engine.set_early_termination(False)
    try:
        job.resident()
        kk = max(1, K // 2)
        ms_noet = timed(job.resident, kk) / kk
    finally:
        engine.set_early_termination(prev_et)
    line["no_early_termination"] = {
        "value": n_frame / (ms_noet * 1e-3), "unit": UNIT, "ms_per_step": ms_noet,
        "note": "engine.set_early_termination(False): samples behind the point where a ray's FP32 transmittance is "
                "exactly 0 are advected and gathered too (they cannot change any output or gradient; "
                "tests/test_gpu_fullsize.py::test_early_termination_changes_nothing)"}

    # ---- the step's fixed costs beside the kernels (strong scaling exposes them): the gradient all-reduce
    if world > 1:
        ms_ar = timed(lambda: sharding.allreduce_grads(params, extras=torch.zeros(1, device=dev),
                                                       flat=engine.last_grad_flat()), K) / K
        ms_nocoll = timed(lambda: job.resident(collective=False), K) / K
        line["strong_breakdown"] = {"ms_step": ms_step, "ms_step_without_collective": ms_nocoll,
                                    "ms_allreduce_alone": ms_ar,
                                    "rays_per_rank": job.n, "note": "max over ranks, CUDA events"}
        # ---- weak scaling (round-1 figure): every rank a whole frame, one all-reduce(mean) per step
        wjob = Job(None)
        wjob.resident()
        ms_weak = timed(wjob.resident, max(1, K // 2)) / max(1, K // 2)
        line["weak"] = {"value": world * n_frame / (ms_weak * 1e-3), "unit": UNIT, "ms_per_step": ms_weak,
                        "note": "each rank renders its own 800x800 frame (identical work per GPU), gradients all-reduced"}
        del wjob

    # ---- per-kernel device time (CUDA events around every launch, on the launching stream)
    if rank == 0:
        out = engine.render_forward(field.binding, job.o_d, job.d_d, T_RENDER, white_bg=True, training=True,
                                    jitter=job.jitter_d, ray_chunk=RAY_CHUNK, want_stats=True)
        torch.cuda.synchronize()
        n_valid, n_adv, n_app, n_mlp = (int(x) for x in out.stats.tolist())
        n_mlp = n_mlp or n_adv
        del out
        _lib.profile_read(reset=True)
        _lib.profile_enable(True)
        P = 2
        for _ in range(P):
            job.resident(collective=False)     # rank 0 only: no collective in the profiled passes
        prof = _lib.profile_read(reset=True)
        _lib.profile_enable(False)
        cnt = engine.LAST_BWD_COUNTERS.view(torch.int64).tolist()
        n_app_bwd, n_adv_bwd = int(cnt[4]), int(cnt[5])
        pk = peaks()
        live = measure_live_peaks(dev)
        line["measured_live"] = live
        tf32_peak = 0.5 * pk["bf16_tflops_sustained"]
        n, S = job.n, 192
        alg = {   # kernel -> (bound, algorithmic work per launch)
            "k_sample_advect": ("tensor", n_adv * 2 * VEL_EVAL_FLOP),
            "k_sample_advect_tc": ("tensor", n_adv * 2 * VEL_EVAL_FLOP),
            "k_advect_bwd": ("tensor", n_adv_bwd * 6 * VEL_EVAL_FLOP),   # 2 evals: fwd recompute + dX + dW GEMMs
            "k_advect_bwd_tc": ("tensor", n_adv_bwd * 6 * VEL_EVAL_FLOP),
            "k_sample_advect_h": ("tensor16", n_mlp * 2 * VEL_EVAL_FLOP),
            "k_advect_bwd_h": ("tensor16", n_adv_bwd * 6 * VEL_EVAL_FLOP),
            "k_march": ("hbm", n_adv * DENSITY_BYTES + n * (44 + 4 * S)),   # gathers only the evaluated samples
            "k_density_bwd": ("hbm", n_valid * 2 * DENSITY_BYTES),
            "k_appearance": ("hbm", n_app * APP_BYTES),
            "k_app_bwd": ("hbm", n_app_bwd * 3 * APP_BYTES),
            "k_march_bwd": ("hbm", n * S * (4 + 4 + 1 + 4) + n * 40),
            "k_composite": ("hbm", n * S * 4 + n_app * 12 + n * 12),
        }
        total_ms = sum(v[0] for v in prof.values())
        kern = {}
        for name, (ms, cnt_l) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            e = {"ms_per_launch": ms / cnt_l, "launches_per_step": cnt_l / P, "share": ms / total_ms}
            if name in alg:
                bound, work = alg[name]
                work = work / (cnt_l / P)      # per launch (the forward kernels run once per depth wave)
                sec = (ms / cnt_l) * 1e-3
                if bound == "tensor16":   # FP16-split path: the denominator is the measured 16-bit dense peak
                    ach = work / sec / 1e12
                    e.update(bound="tensor", achieved=ach, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                             frac=ach / pk["bf16_tflops_sustained"], mma_per_gemm=3,
                             # FLOPs the tensor cores actually execute (3 MMAs per algorithmic GEMM) against the
                             # same peak: how busy the MMA pipe is kept, as opposed to how much useful work it does
                             mma_executed_tflops=3 * ach, mma_executed_frac=3 * ach / pk["bf16_tflops_sustained"])
                elif bound == "tensor":
                    ach = work / sec / 1e12
                    e.update(bound="tensor", achieved=ach, peak=tf32_peak, unit="TFLOP/s", frac=ach / tf32_peak)
                else:
                    ach = work / sec / 1e9
                    e.update(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"])
                    if live.get("l2_copy_gbs"):
                        # the factor planes (<= 37 MB) are L2-resident by design: the meaningful ceiling of a
                        # gather kernel is the L2, not HBM
                        e["l2"] = {"achieved": ach, "peak": live["l2_copy_gbs"], "unit": "GB/s",
                                   "frac": ach / live["l2_copy_gbs"],
                                   "peak_source": "L2-resident copy measured live (measured_live.l2_copy_gbs)"}
            kern[name] = e
        # DRAM / L2 traffic per launch: per-unit dram__bytes (read + write) and lts__t_bytes of one
        # `ncu --set full` capture (profiles/ncu_traffic.json, tools/ncu_traffic.py) x the units this
        # run's launch processed
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            units = {"valid_samples": n_valid, "advected_samples": n_adv, "advected_samples_bwd": n_adv_bwd,
                     "app_samples": n_app, "app_samples_bwd": n_app_bwd}
            for kname, kv in tj.get("kernels", {}).items():
                if kname in kern and kv["unit"] in units:
                    kern[kname]["traffic"] = kv["dram_bytes_per_unit"] * units[kv["unit"]]
                    if "lts_bytes_per_unit" in kv:
                        kern[kname]["l2_traffic"] = kv["lts_bytes_per_unit"] * units[kv["unit"]]
                    for extra in ("lts_pct_of_peak", "tensor_pipe_pct", "l2_hit_pct"):
                        if extra in kv:
                            kern[kname]["ncu_" + extra] = kv[extra]
        top = next(iter(kern))
        r = dict(kern[top])
        line["roofline"] = {"kernel": top, "bound": r.get("bound"), "achieved": r.get("achieved"),
                            "peak": r.get("peak"), "unit": r.get("unit"), "frac": r.get("frac"),
                            "traffic": r.get("traffic"), "traffic_unit": "bytes/launch",
                            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per unit "
                                              "(profiles/ncu_traffic.json, 50-row capture) x units of this launch",
                            "share_of_step": r["share"],
                            "peak_source": pk["source"] + (
                                "; 16-bit dense tensor peak (sustained); the FP16 hi+lo operand split issues 3 MMAs "
                                "per algorithmic GEMM, so the ceiling of this fraction is 1/3"
                                if r.get("mma_per_gemm") == 3 else
                                ("; TF32 peak = 0.5 x measured sustained bf16" if r.get("bound") == "tensor" else ""))}
        line["roofline_gather"] = dict(kern.get("k_march", {}), kernel="k_march",
                                       note="TensoRF density gather + alpha scan; planes are L2-resident, "
                                            "so algorithmic GB/s may exceed the HBM copy peak: see the 'l2' entry")
        line["kernels"] = kern
        line["counts"] = {"valid_samples": n_valid, "advected_samples": n_adv, "mlp_samples": n_mlp, "app_samples": n_app,
                          "note": "valid = in-box samples; advected = those in front of each ray's termination; mlp = advected "
                                  "samples inside the velocity gate (the others do not move)",
                          "app_samples_bwd": n_app_bwd, "advected_samples_bwd": n_adv_bwd}
    for p in params:
        p.grad = None

    # ---- secondary figure (SURVEY.md 8d ii): one 800x800 eval frame, mode='test', sharded + ONE all-gather
    def eval_leg(fld, t, o_all, d_all, white_bg, transfer, chunk=RAY_CHUNK, reps=3):
        """Full-frame eval render through the sharded path: each rank renders its interleaved chunks, one
        all-gather assembles the frame on every rank (sharding.gather_frame_interleaved)."""
        nfr = o_all.shape[0]
        idx = sharding.shard_index(nfr, rank, world, chunk) if world > 1 else None
        o_l = (o_all if idx is None else o_all[idx]).contiguous().to(dev)
        d_l = (d_all if idx is None else d_all[idx]).contiguous().to(dev)
        fld.eval()

        def once():
            with torch.no_grad():
                rgb, depth, acc, w, mk = fld.render_rays(t, o_l, d_l, white_bg=white_bg, ray_chunk=chunk,
                                                        transfer_vel=transfer)
                parts = [rgb, depth, acc] + ([mk] if fld.mask_field is not None else [])
                if world > 1:
                    parts = sharding.gather_frame_interleaved(parts, nfr, chunk)
            return parts
        once()
        ms = timed(once, reps) / reps
        parts = once()
        chk = float(parts[0].double().sum().item())
        return {"ms": ms, "rays_per_s": nfr / (ms * 1e-3), "rays": nfr, "n_gpus": world,
                "samples_per_ray": int(fld.nSamples), "rgb_checksum": chk}

    try:
        ev = eval_leg(field, T_RENDER, o_f, d_f, True, False)
        ev["note"] = "bat, t=0.33, render only (mode='test'), inputs resident, frame sharded + all-gather when n_gpus > 1"
        line["eval_frame"] = ev
    except Exception as exc:   # never lose the bench line over a secondary figure
        line["eval_frame"] = {"error": str(exc)[:200]}
    field.train()

    if args.rows == H and not args.no_extras:
        del job
        line.update(extra_configs(args, dev, rank, world, eval_leg, timed))

    if rank == 0:
        if world == 1 and not args.no_cpu:
            leg = CpuLeg()
            leg.run(0.0, 1)     # warm-up chunk (thread pool, allocator)
            leg.valid = leg.samples = 0
            tt, rr, k = leg.run(args.cpu_budget, 10)
            line["cpu_baseline"] = {"value": rr / tt, "unit": UNIT, "cores": leg.cores, "kind": leg.kind,
                                    "sample": leg.describe(k, n_valid / float(n_frame * 192))}
        if args.rows != H:
            line["invalid_for_bench"] = f"profiling run on {args.rows} of {H} rows"
        emit(line)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


def extra_configs(args, dev, rank, world, eval_leg, timed):
    """BASELINE.json configs[2..4] as extra keys of the line (never the headline):
      pde_262144            fallingball: NVFi.get_vel_loss(262 144) forward + backward (models/nvfi.py:42-84)
      chessboard_eval_t1.0  chessboard (VelocityAABBSur, K=4, non-white background, the shipped step_ratio):
                            one 800x800 frame at the extrapolated time t=1.0 (2 RK2 steps), ray-sharded
      fan_mask_render       fan + MaskField(mask_dim=8) composited in the render, transfer_vel (test_segm_render.py:75-99)
    """
    from nvfi_b200 import _lib
    from nvfi_b200 import models as M
    from nvfi_b200.scenes import build_scene, frame_rays
    res = {}
    torch.cuda.empty_cache()
    # -- config 3: PDE loss
    try:
        cfg, nv, _ = build_scene("fallingball", grid=GRID, device=dev)
        nv.requires_grad_(True)
        npts = int(cfg.experiment.vel_reg_n_pts)
        g = torch.Generator(device=dev).manual_seed(5 + rank)
        lo, hi = nv.nvfi.aabb
        pts = nv.nvfi.normalize_coord(torch.rand(npts, 3, device=dev, generator=g) * (hi - lo) + lo)
        tt = torch.rand(npts, 1, device=dev, generator=g)

        def pde_step():
            nv.zero_grad(set_to_none=True)
            loss = nv.get_vel_loss(npts, points=pts, t=tt)
            if torch.is_tensor(loss):
                loss.backward()
            return loss
        for _ in range(3):
            pde_step()
        ms = timed(pde_step, 5) / 5
        if rank == 0:
            _lib.profile_read(reset=True)
            _lib.profile_enable(True)
            loss = pde_step()
            torch.cuda.synchronize()
            prof = _lib.profile_read(reset=True)
            _lib.profile_enable(False)
            from nvfi_b200 import pde as _pde
            with torch.no_grad():
                kept = int(_pde.occupancy_filter(nv.nvfi, pts, tt).sum())
            # forward-mode Jacobian: 5 rows per point through weight_net + 1 through a_weight_net, and the
            # reverse pass of both (x2): algorithmic FLOPs of the loss AND its gradients
            flop = kept * (5 + 1) * VEL_EVAL_FLOP * 3
            pk = peaks()
            res["pde_262144"] = {
                "ms": ms, "points": npts, "occupied_points": kept, "loss": float(loss),
                "kernels_ms": {k: v[0] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]},
                "roofline": {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12,
                             "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": flop / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                             "note": "whole get_vel_loss call (occupancy filter + forward-mode Jacobian + reverse pass, "
                                     "host glue included) against the 16-bit dense peak; k_pde_jac_h / k_accnet_bwd_h "
                                     "run on tcgen05 with FP16-split operands (3 MMAs per GEMM: ceiling 1/3)"},
                "note": "fallingball.yaml: get_vel_loss(262144) + backward, per rank (replicated), max over ranks"}
        del nv
    except Exception as exc:
        res["pde_262144"] = {"error": str(exc)[:300]}
    torch.cuda.empty_cache()
    # -- config 4: chessboard future-frame extrapolation, ray-sharded eval
    try:
        cfg, nv, _ = build_scene("chessboard", grid=GRID, device=dev)
        o, d = frame_rays(H, W, theta=30.0, phi=-35.0, radius=4.5, z_shift=3.0)
        r = eval_leg(nv.nvfi, 1.0, o, d, False, False, reps=2)
        r["note"] = ("chessboard.yaml (VelocityAABBSur, K=4, step_ratio 0.5), 800x800 eval at t=1.0 "
                     "(extrapolation: 2 RK2 steps), sharded over the ranks + one all-gather")
        res["chessboard_eval_t1.0"] = r
        del nv
    except Exception as exc:
        res["chessboard_eval_t1.0"] = {"error": str(exc)[:300]}
    torch.cuda.empty_cache()
    # -- config 5: fan + mask field
    try:
        cfg, nv, _ = build_scene("fan", grid=GRID, device=dev)
        torch.manual_seed(17)
        mf = M.MaskField(n_layer=4, n_dim=128, input_dim=3, skips=[], mask_dim=int(cfg.segmentation.n_object),
                         mask_act="softmax")
        nv.nvfi.mask_field = mf.to(dev)
        o, d = frame_rays(H, W, theta=30.0)
        r = eval_leg(nv.nvfi, 0.5, o, d, True, True, reps=2)
        r["note"] = ("fan.yaml + MaskField(n_layer=4, n_dim=128, mask_dim=8) composited in the render, "
                     "mode='test', transfer_vel=True (test_segm_render.py:75-99), sharded + one all-gather")
        res["fan_mask_render"] = r
        del nv
    except Exception as exc:
        res["fan_mask_render"] = {"error": str(exc)[:300]}
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nvfi_b200", choices=["nvfi_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs for BASELINE.json configs 3-5")
    ap.add_argument("--rows", type=int, default=H,
                    help="PROFILING ONLY: render the first ROWS rows of the 800x800 frame (ncu replays are "
                         "slow on the 6 GB full-frame working set); the line is marked invalid_for_bench")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
