"""CPU tests: the C oracle (oracle/mvp_oracle.c) against the reference's own PyTorch autograd loop restated in
oracle/torch_ref.py (mvpraymarch.py:567-633), in fp64 where both are exact up to measure-zero boundary rules."""
import numpy as np
import torch

from oracle import oracle, torch_ref
from tests.helpers import gradcheck_like_scene, relerr, scene_args_np


def _torch_fwd_bwd(s, grad, dtype):
    t = {k: (v.to(dtype) if torch.is_tensor(v) else v) for k, v in s.items()}
    return torch_ref.raymarch_torch_fwd_bwd(t["raypos"], t["raydir"], t["stepsize"], t["tminmax"], t["primpos"],
                                            t["primrot"], t["primscale"], t["template"], grad.to(dtype),
                                            fadescale=s["fadescale"], fadeexp=s["fadeexp"])


def test_oracle_f64_matches_torch_autograd_loop():
    s = gradcheck_like_scene(N=2, H=13, W=13, k3=2, M=4)
    g = torch.Generator().manual_seed(5)
    grad = torch.randn(2, 13, 13, 4, generator=g)
    ref = _torch_fwd_bwd(s, grad, torch.float64)
    a, kw = scene_args_np(s, np.float64)
    rgba, raysat = oracle.forward(*a, dtype=np.float64, **kw)
    nsat = int((raysat[..., 0] > -1).sum())
    assert 10 < nsat < 2 * 13 * 13 - 10, "scene must exercise both saturated and unsaturated rays (%d)" % nsat
    assert relerr(rgba, ref[0].numpy()) < 1e-11
    grads = oracle.backward(*a, grad.numpy().astype(np.float64), raysat, dtype=np.float64, **kw)
    for name, mine, theirs in zip(("primpos", "primrot", "primscale", "template"), grads, ref[1:]):
        assert relerr(mine, theirs.numpy()) < 1e-9, name


def test_oracle_f32_close_to_f64():
    s = gradcheck_like_scene(N=1, H=24, W=20, k3=3, M=6)
    a32, kw = scene_args_np(s, np.float32)
    a64, _ = scene_args_np(s, np.float64)
    r32, s32 = oracle.forward(*a32, dtype=np.float32, **kw)
    r64, s64 = oracle.forward(*a64, dtype=np.float64, **kw)
    assert relerr(r32, r64) < 2e-5
    g = np.random.default_rng(3).standard_normal(r32.shape).astype(np.float32)
    g32 = oracle.backward(*a32, g, s32, dtype=np.float32, **kw)
    g64 = oracle.backward(*a64, g.astype(np.float64), s64, dtype=np.float64, **kw)
    for x, y in zip(g32, g64):
        assert relerr(x, y) < 2e-4


def test_oracle_blocksize_invariance_without_cap():
    """Below the 512 cap the result does not depend on how pixels are grouped into warps."""
    s = gradcheck_like_scene(N=1, H=17, W=19, k3=2, M=4)
    a, kw = scene_args_np(s, np.float64)
    r1, _ = oracle.forward(*a, dtype=np.float64, blocksize=(8, 16), **kw)
    r2, _ = oracle.forward(*a, dtype=np.float64, blocksize=(32, 1), **kw)
    assert relerr(r1, r2) < 1e-12


def test_oracle_hitbox_cap_changes_result():
    """With maxhitboxes smaller than the candidate count the warp list is truncated in DFS order (utils.h:779-781)."""
    s = gradcheck_like_scene(N=1, H=9, W=9, k3=2, M=4)
    a, kw = scene_args_np(s, np.float64)
    full, _, st = oracle.forward(*a, dtype=np.float64, return_stats=True, **kw)
    cut, _, st2 = oracle.forward(*a, dtype=np.float64, maxhitboxes=2, return_stats=True, **kw)
    assert st2["capped_warps"] > 0 and st["capped_warps"] == 0
    assert relerr(cut, full) > 1e-3


def test_oracle_aabb_contains_corners():
    s = gradcheck_like_scene(N=1, k3=2)
    pos, rot, sc = (s[k][0].numpy().astype(np.float64) for k in ("primpos", "primrot", "primscale"))
    bb = oracle.aabb(pos, rot, sc, dtype=np.float64)
    K = pos.shape[0]
    for k in range(K):
        for c in range(8):
            sg = np.array([(c & 1) * 2 - 1, ((c >> 1) & 1) * 2 - 1, ((c >> 2) & 1) * 2 - 1], np.float64)
            p = rot[k] @ (sg / sc[k]) + pos[k]
            assert (p >= bb[K - 1 + k, 0] - 1e-12).all() and (p <= bb[K - 1 + k, 1] + 1e-12).all()
    assert (bb[0, 0] <= bb[K - 1:, 0].min(0) + 1e-12).all() and (bb[0, 1] >= bb[K - 1:, 1].max(0) - 1e-12).all()


# ---------------------------------------------------------------------------------------------------------
# Pin against the REFERENCE ITSELF: tests/golden/*.npz are outputs of the unmodified reference CUDA extension
# (compiled for sm_100 by oracle/build_ref.py, run on a B200 by tests/golden/make_golden.py).
# ---------------------------------------------------------------------------------------------------------
import os  # noqa: E402

import pytest  # noqa: E402

from tests.helpers import CASES, build_case  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_f32_matches_reference_cuda_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    assert os.path.exists(path), "golden vector missing: run tests/golden/make_golden.py on the GPU box"
    gold = np.load(path)
    s, grad = build_case(name)
    a, kw = scene_args_np(s)
    rgba, raysat = oracle.forward(*a, **kw)
    assert relerr(rgba, gold["rayrgba"]) < 5e-6
    # same rays saturate, same saturation colour
    assert np.array_equal(raysat[..., 0] > -1, gold["raysat"][..., 0] > -1)
    assert relerr(raysat, gold["raysat"]) < 5e-6
    g = oracle.backward(*a, grad.numpy(), gold["raysat"], **kw)
    for nm, x in zip(("primpos", "primrot", "primscale", "template", "warp"), g):
        assert relerr(x, gold["grad_" + nm]) < 2e-5, nm


def test_oracle_f64_warp_field_matches_torch_autograd_loop():
    """algo 1 (PrimSamplerTW<true>, primsampler.h:53-58, 82-88): payload sampled at a warp-field-displaced position."""
    from tests.helpers import make_warp
    s = gradcheck_like_scene(N=1, H=11, W=10, k3=2, M=4, alpha_gain=25.0)
    warp = make_warp(1, 8, 3, 3, 3, amp=0.15)      # large noise: warped positions leave [-1,1] (zero-padding path)
    g = torch.Generator().manual_seed(9)
    grad = torch.randn(1, 11, 10, 4, generator=g)
    t = {k: (v.double() if torch.is_tensor(v) else v) for k, v in s.items()}
    ref = torch_ref.raymarch_torch_fwd_bwd(t["raypos"], t["raydir"], t["stepsize"], t["tminmax"], t["primpos"], t["primrot"],
                                           t["primscale"], t["template"], grad.double(), fadescale=s["fadescale"],
                                           fadeexp=s["fadeexp"], warp=warp.double())
    a, kw = scene_args_np(s, np.float64)
    rgba, raysat = oracle.forward(*a, dtype=np.float64, warp=warp.numpy().astype(np.float64), **kw)
    assert 5 < int((raysat[..., 0] > -1).sum()) < 105
    assert relerr(rgba, ref[0].numpy()) < 1e-11
    grads = oracle.backward(*a, grad.numpy().astype(np.float64), raysat, dtype=np.float64, warp=warp.numpy().astype(np.float64), **kw)
    for name, mine, theirs in zip(("primpos", "primrot", "primscale", "template", "warp"), grads, ref[1:]):
        assert relerr(mine, theirs.numpy()) < 1e-9, name
    # and it is really a different image than without the warp
    plain, _ = oracle.forward(*a, dtype=np.float64, **kw)
    assert relerr(rgba, plain) > 1e-3
