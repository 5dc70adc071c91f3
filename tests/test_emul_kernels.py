"""The PRODUCT kernels' source (ava-256_b200/csrc/mvp_kernels.cu) compiled for the host on a CPU emulation of warps,
blocks and shared memory (tests/emul/), run against the oracle and the reference's golden vectors WITHOUT a GPU.

This is test infrastructure: it checks the kernels' logic (accel build, tile lists, interval marching, sample compaction,
slab-major adjoint, the 256/512-entry variants, algo 1) in a container that has no GPU; floating point is IEEE on the
CPU (no MUFU approximations), so agreement is to rounding.  The GPU tests (`-m gpu`) remain the parity gate."""
import os

import numpy as np
import pytest

import torch

from tests.helpers import CASES, EDGE_KINDS, build_case, edge_scene, relerr, scene_args_np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD_TOL, BWD_TOL = 1e-5, 1e-4        # vs the oracle (same IEEE arithmetic; gradients differ by summation order)


@pytest.fixture()
def kernels():
    from tests.emul import kernels as k
    k.use_variant(())
    k.load()
    yield k
    k.set_lane_order("forward")
    k.use_variant(())


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_kernels_vs_oracle(kernels, name):
    from oracle import oracle
    s, grad = build_case(name)
    a, kw = scene_args_np(s)
    kernels.set_lane_order("forward")
    out, raysat, grads = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    ref, rsat = oracle.forward(*a, **kw)
    assert relerr(out, ref) <= FWD_TOL
    assert np.array_equal(raysat[..., 0] > -1.0, rsat[..., 0] > -1.0)            # same rays saturate
    gref = oracle.backward(*a, grad.numpy(), rsat, **kw)
    assert len(grads) == len(gref)
    for nm, g, r in zip(("primpos", "primrot", "primscale", "template", "warp"), grads, gref):
        assert relerr(g, r) <= BWD_TOL, nm
    # inference path (no raysat / rayaux): same image
    out2, _, _ = kernels.forward_backward(*a, **kw)
    assert np.array_equal(out2, out)


@pytest.mark.parametrize("name", ["head_small", "many_overlaps", "warp_small", "gradcheck_ragged"])
def test_emulated_kernels_vs_reference_golden(kernels, name):
    """Vectors produced by the unmodified reference CUDA extension on a B200 (tests/golden/make_golden.py)."""
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    s, grad = build_case(name)
    a, kw = scene_args_np(s)
    out, _, grads = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    assert relerr(out, gold["rayrgba"]) <= 1e-4
    for nm, g in zip(("primpos", "primrot", "primscale", "template", "warp"), grads):
        assert relerr(g, gold["grad_" + nm]) <= 1e-3, nm


@pytest.mark.parametrize("order", ["reverse", "random"])
@pytest.mark.parametrize("name", ["head_small", "many_overlaps", "warp_small", "tiny"])
def test_result_does_not_depend_on_lane_schedule(kernels, name, order):
    """Lanes of a warp have no defined order between collectives; a kernel that relies on one (missing __syncwarp
    between a shared-memory write and another lane's read) gives different images under different schedules."""
    s, grad = build_case(name)
    a, kw = scene_args_np(s)
    kernels.set_lane_order("forward")
    out0, sat0, g0 = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    kernels.set_lane_order(order)
    out1, sat1, g1 = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    kernels.set_lane_order("forward")
    assert np.array_equal(out0, out1) and np.array_equal(sat0, sat1)        # forward has no atomics: bit-identical
    for x, y in zip(g0, g1):
        assert relerr(x, y) <= 1e-5                                         # atomics: summation order only


def test_emulated_argument_validation_matches_product(kernels):
    import ctypes
    from ava256_b200 import lib
    L = kernels.load()
    a = lib.ForwardArgs()
    a.shape = lib.Shape(1, 8, 8, 4, 2, 2, 2)
    a.stepsize = 0.1
    assert L.mvp_raymarch_forward(ctypes.byref(a), None) == lib.LIB.mvp_raymarch_forward(ctypes.byref(a), None) == -1
    s = lib.Shape(2, 64, 42, 64, 8, 8, 8)
    assert L.mvp_workspace_bytes(ctypes.byref(s)) == lib.LIB.mvp_workspace_bytes(ctypes.byref(s)) > 0


@pytest.mark.parametrize("variant", [(), ("MVP_LIST_MARGIN=1", "MVP_LIST_REUSE=1"), ("MVP_LIST_MARGIN=0", "MVP_LIST_REUSE=0")])
@pytest.mark.parametrize("kind", EDGE_KINDS)
def test_emulated_edge_cases_vs_oracle(kernels, kind, variant):
    """Degenerate inputs the reference tolerates (zero scale = infinite slab, rays missing the volume, one-step and
    many-step marches), for the default build and the list variants either way round."""
    from oracle import oracle
    s = edge_scene(kind)
    grad = torch.randn(*s["raypos"].shape[:3], 4, generator=torch.Generator().manual_seed(17))
    a, kw = scene_args_np(s)
    kernels.use_variant(variant)
    out, _, grads = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    ref, raysat = oracle.forward(*a, **kw)
    assert np.isfinite(out).all() and relerr(out, ref) <= FWD_TOL
    gref = oracle.backward(*a, grad.numpy(), raysat, **kw)
    for nm, g, r in zip(("primpos", "primrot", "primscale", "template"), grads, gref):
        assert np.isfinite(g).all(), nm
        assert relerr(g, r) <= BWD_TOL, nm


# build-time variants of the kernels (former defaults, and candidates waiting for a GPU measurement): same results required
VARIANTS = {
    "list_rebuild": ("MVP_LIST_REUSE=0",),                                 # backward rebuilds the lists (the former default)
    "list_reuse": ("MVP_LIST_REUSE=1",),                                   # backward loads the lists the forward saved (default)
    "list_reuse_overflow": ("MVP_LIST_REUSE=1", "MVP_LIST_CAP_PER_TILE=1", "MVP_LIST_CAP_MIN=16"),  # most tiles do not fit: mixes both paths
    "bwd_record": ("MVP_BWD_OPAQUE=2",),
    "fwd_arrays": ("MVP_FWD_OPAQUE=0",),
    "xbuckets": ("MVP_XBUCKETS=1",),                                        # second bucketing level in x
    "xbuckets_all": ("MVP_XBUCKETS=1", "MVP_LIST_MARGIN=1"),
    "list_margin": ("MVP_LIST_MARGIN=1",),                                  # step intervals from the fp-drift bound
    "list_margin_reuse": ("MVP_LIST_MARGIN=1", "MVP_LIST_REUSE=1"),
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", ["head_small", "many_overlaps", "warp_small", "gradcheck_ragged", "noncubic", "tiny"])
def test_kernel_variants_match_the_default_build(kernels, name, variant):
    s, grad = build_case(name)
    a, kw = scene_args_np(s)
    out0, sat0, g0 = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    kernels.use_variant(VARIANTS[variant])
    for order in ("forward", "random"):
        kernels.set_lane_order(order)
        out1, sat1, g1 = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
        assert np.array_equal(out0, out1) and np.array_equal(sat0, sat1)
        for x, y in zip(g0, g1):
            assert relerr(x, y) <= 1e-5
    if variant in ("list_reuse", "list_reuse_overflow"):
        import ctypes
        loaded, rebuilt = ctypes.c_int(), ctypes.c_int()
        kernels.load().mvp_emul_saved_list_tiles(ctypes.byref(loaded), ctypes.byref(rebuilt))
        if variant == "list_reuse":
            assert loaded.value > 0 and rebuilt.value == 0                # every tile's saved list was loaded
        else:
            assert loaded.value + rebuilt.value > 0
            if name == "head_small":
                assert loaded.value > 0 and rebuilt.value > 0             # tiny workspace: both paths in one launch
