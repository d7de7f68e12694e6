"""Host logic that decides the integers of a render call: key-frame snap (round half to even),
the isclose key-frame test and the RK2 step count (models/tensorf_keyframe.py:646-654, 683-699,
575-609).  CPU only."""
import math
from types import SimpleNamespace

import pytest
import torch

from nvfi_b200 import engine


def _field(K=16, tmax=0.75, use_vel=True):
    return SimpleNamespace(num_keyframes=K, tmax=tmax, use_vel=use_vel)


def _reference_plan(t, K, tmax, transfer):
    """Restatement of models/tensorf_keyframe.py:646-654, 683 with FP32 tensors."""
    tt = torch.ones(1, dtype=torch.float32) * t
    tsf = tmax / (K - 1)
    base = torch.zeros_like(tt) if transfer else torch.round((tt / tsf).clamp(0.0, K - 1)) * tsf
    key = bool(torch.isclose(tt, base))
    return float(tt), float(base), key


@pytest.mark.parametrize("K,tmax", [(16, 0.75), (4, 0.75), (2, 1.0)])
@pytest.mark.parametrize("t", [0.0, 0.025, 0.075, 0.1, 0.125, 0.33, 0.375, 0.5, 0.74999, 0.75, 0.8, 1.0, 1.25, 2.0])
@pytest.mark.parametrize("transfer", [False, True])
def test_time_plan_matches_reference_semantics(K, tmax, t, transfer):
    f = _field(K, tmax)
    tt, base, tnb, advect = engine.time_plan(f, t, transfer)
    rt, rbase, rkey = _reference_plan(t, K, tmax, transfer)
    assert tt == rt and base == rbase
    assert advect == (not rkey)
    assert tnb == pytest.approx(2 * rbase / tmax - 1, abs=1e-6)


def test_half_way_times_round_to_even_keyframes():
    """torch.round is half-to-even: t / tsf = 0.5 -> key frame 0, 1.5 -> 2, 2.5 -> 2."""
    f = _field(K=16, tmax=0.75)        # tsf = 0.05
    tsf = 0.75 / 15
    for k_half, want in ((0.5, 0), (1.5, 2), (2.5, 2), (3.5, 4)):
        t = float(torch.tensor(k_half, dtype=torch.float32) * tsf)
        q = float(torch.ones(1) * t / tsf)
        if q != k_half:      # not exactly representable: the FP32 quotient decides, as in the reference
            want = int(torch.round(torch.tensor(q)))
        _, base, _, _ = engine.time_plan(f, t, False)
        assert base == pytest.approx(want * tsf, abs=1e-7)


def test_rk2_step_count_of_the_schedule():
    """dt_max = 0.5 tmax / (K - 1); the loop runs until the offset is exactly 0 (:577-609)."""
    for K, tmax, t, want in ((16, 0.75, 0.33, 1), (16, 0.75, 1.0, 10), (4, 0.75, 1.0, 2), (16, 0.75, 0.3, 0)):
        f = _field(K, tmax)
        tt, base, _, advect = engine.time_plan(f, t, False)
        dt_max = 0.5 * tmax / (K - 1)
        off, steps = torch.tensor(tt - base, dtype=torch.float32), 0
        while advect and float(off.abs()) > 0 and steps < 64:
            dt = torch.sign(off) * torch.minimum(off.abs(), torch.tensor(dt_max, dtype=torch.float32))
            off = off - dt
            steps += 1
        assert steps == want, (K, t, steps)


def test_no_velocity_field_never_advects():
    f = _field(use_vel=False)
    tt, base, tnb, advect = engine.time_plan(f, 0.33, False)
    assert not advect and tnb == pytest.approx(2 * 0.33 / 0.75 - 1, abs=1e-6)


def test_packed_copy_cache_key_and_invalidate():
    """engine._Tracked: in-place ops bump the version and make the key stale; a write
    through .data does not — FieldBinding.invalidate() (called by load_state_dict / upsample / shrink /
    updateAlphaMask) does."""
    import torch
    from nvfi_b200 import engine
    p = torch.nn.Parameter(torch.zeros(4))
    tr = engine._Tracked()
    assert tr.stale((p,)) and not tr.stale((p,))
    with torch.no_grad():
        p.add_(1)
    assert tr.stale((p,)) and not tr.stale((p,))
    p.data.add_(1)                       # bypasses the version counter
    assert not tr.stale((p,))
    engine.FieldBinding.invalidate()
    assert tr.stale((p,)) and not tr.stale((p,))


def test_single_jitter_draw_equals_per_chunk_draws():
    """field.render_rays draws the stratified jitter of all chunks at once when no per-chunk background
    draw is interleaved (white background): the CPU generator's stream must not depend on the split."""
    import torch
    for sizes in ([2048] * 5 + [1024], [7, 9, 16, 5], [1, 1, 1, 31], [4096, 3]):
        torch.manual_seed(3)
        a = torch.cat([torch.rand(s, 1) for s in sizes])
        torch.manual_seed(3)
        b = torch.rand(sum(sizes), 1)
        assert torch.equal(a, b), sizes
