"""Epilogue kernels: image compositing after the raymarcher, payload hand-off before it.

CPU part: the host build of the kernels' per-element bodies (tests/emul) and the eager-PyTorch restatement
(oracle/epilogue_ref.py) against each other and against vectors generated with the reference's own modules
(tests/golden/make_epilogue_golden.py); argument validation of the C-ABI without a device.
GPU part: the CUDA kernels through the Python mirror against the same restatement and the same vectors."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle.epilogue_ref import assemble_payload_ref, composite_ref

HERE = os.path.dirname(os.path.abspath(__file__))
F32P = ctypes.POINTER(ctypes.c_float)


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(F32P)


@pytest.fixture(scope="module")
def emul():
    from tests.emul.build import build
    lib = ctypes.CDLL(build())
    lib.emul_composite_forward.argtypes = [ctypes.c_int] * 4 + [F32P] * 6
    lib.emul_composite_backward.argtypes = [ctypes.c_int] * 4 + [F32P] * 9
    lib.emul_payload_forward.argtypes = [ctypes.c_int] * 6 + [F32P] * 2 + [ctypes.c_float] * 2 + [F32P]
    lib.emul_payload_backward.argtypes = [ctypes.c_int] * 6 + [F32P] * 2 + [ctypes.c_float] + [F32P] * 2
    return lib


def _composite_case(N, H, W, with_cc, with_bg, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    rayrgba = torch.rand(N, H, W, 4, generator=g)
    rayrgba[..., :3] *= 255.0
    ccw = (1.0 + 0.3 * torch.randn(N, 3, generator=g)) if with_cc else None
    ccb = (5.0 * torch.randn(N, 3, generator=g)) if with_cc else None
    bg = (255.0 * torch.rand(N, 3, H, W, generator=g)) if with_bg else None
    g_rgb = torch.randn(N, 3, H, W, generator=g)
    g_alpha = torch.randn(N, 1, H, W, generator=g)
    mv = lambda t: None if t is None else t.to(device)  # noqa: E731
    return tuple(mv(t) for t in (rayrgba, ccw, ccb, bg, g_rgb, g_alpha))


def _composite_ref_grads(rayrgba, ccw, ccb, bg, g_rgb, g_alpha):
    leaves = [None if t is None else t.detach().clone().requires_grad_(True) for t in (rayrgba, ccw, ccb, bg)]
    rgb, alpha = composite_ref(*leaves)
    ((rgb * g_rgb).sum() + (alpha * g_alpha).sum()).backward()
    return rgb.detach(), alpha.detach(), [None if t is None else t.grad for t in leaves]


@pytest.mark.parametrize("V", [1, 4])
@pytest.mark.parametrize("with_cc,with_bg", [(False, False), (True, False), (False, True), (True, True)])
def test_emul_composite_matches_eager_torch(emul, V, with_cc, with_bg):
    N, H, W = 3, 6, 10
    rayrgba, ccw, ccb, bg, g_rgb, g_alpha = _composite_case(N, H, W, with_cc, with_bg, seed=V)
    ref_rgb, ref_alpha, (gr, gw, gb, gbg) = _composite_ref_grads(rayrgba, ccw, ccb, bg, g_rgb, g_alpha)
    a = {k: (None if v is None else np.ascontiguousarray(v.numpy())) for k, v in
         dict(rayrgba=rayrgba, ccw=ccw, ccb=ccb, bg=bg, g_rgb=g_rgb, g_alpha=g_alpha).items()}
    rgb = np.empty((N, 3, H, W), np.float32)
    alpha = np.empty((N, 1, H, W), np.float32)
    assert emul.emul_composite_forward(N, H, W, V, _np_ptr(a["rayrgba"]), _np_ptr(a["ccw"]), _np_ptr(a["ccb"]), _np_ptr(a["bg"]),
                                       _np_ptr(rgb), _np_ptr(alpha)) == 0
    assert np.array_equal(rgb, ref_rgb.numpy())           # bit-exact: same ops, same roundings
    assert np.array_equal(alpha, ref_alpha.numpy())
    grad_rayrgba = np.empty((N, H, W, 4), np.float32)
    grad_ccw = np.zeros((N, 3), np.float32) if with_cc else None
    grad_ccb = np.zeros((N, 3), np.float32) if with_cc else None
    grad_bg = np.empty((N, 3, H, W), np.float32) if with_bg else None
    assert emul.emul_composite_backward(N, H, W, V, _np_ptr(a["rayrgba"]), _np_ptr(a["ccw"]), _np_ptr(a["bg"]), _np_ptr(a["g_rgb"]),
                                        _np_ptr(a["g_alpha"]), _np_ptr(grad_rayrgba), _np_ptr(grad_ccw), _np_ptr(grad_ccb),
                                        _np_ptr(grad_bg)) == 0
    np.testing.assert_allclose(grad_rayrgba, gr.numpy(), rtol=1e-6, atol=1e-4)
    if with_cc:
        np.testing.assert_allclose(grad_ccw, gw.numpy(), rtol=1e-5, atol=1e-2)
        np.testing.assert_allclose(grad_ccb, gb.numpy(), rtol=1e-5, atol=1e-4)
    if with_bg:
        np.testing.assert_allclose(grad_bg, gbg.numpy(), rtol=1e-6, atol=1e-6)


PAYLOAD_SHAPES = [  # N, hb, wb, B, V, BT
    (2, 2, 3, 8, 4, 8), (2, 2, 3, 8, 4, 0), (2, 2, 3, 8, 1, 0), (1, 3, 2, 4, 4, 0), (1, 4, 2, 3, 1, 0), (1, 1, 1, 1, 1, 0),
    (1, 2, 2, 16, 4, 0),
]


def _payload_case(N, hb, wb, B, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    tex = torch.randn(N, 3 * B, hb * B, wb * B, generator=g) * 4.0 - 2.0
    opacity = torch.randn(N, B, hb * B, wb * B, generator=g)
    gt = torch.randn(N, hb * wb, B, B, B, 4, generator=g)
    return tex.to(device), opacity.to(device), gt.to(device)


def _payload_ref_grads(tex, opacity, gt, B):
    t, o = tex.detach().clone().requires_grad_(True), opacity.detach().clone().requires_grad_(True)
    tp = assemble_payload_ref(t, o, B)
    (tp * gt).sum().backward()
    return tp.detach(), t.grad, o.grad


@pytest.mark.parametrize("N,hb,wb,B,V,BT", PAYLOAD_SHAPES)
def test_emul_payload_matches_eager_torch(emul, N, hb, wb, B, V, BT):
    tex, opacity, gt = _payload_case(N, hb, wb, B, seed=B)
    ref, gtex, gop = _payload_ref_grads(tex, opacity, gt, B)
    tplate = np.full((N, hb * wb, B, B, B, 4), np.nan, np.float32)
    assert emul.emul_payload_forward(N, hb, wb, B, V, BT, _np_ptr(tex.numpy()), _np_ptr(opacity.numpy()), 25.0, 100.0,
                                     _np_ptr(tplate)) == 0
    assert np.array_equal(tplate, ref.numpy())
    grad_tex = np.full(tuple(tex.shape), np.nan, np.float32)
    grad_op = np.full(tuple(opacity.shape), np.nan, np.float32)
    assert emul.emul_payload_backward(N, hb, wb, B, V, BT, _np_ptr(tplate), _np_ptr(np.ascontiguousarray(gt.numpy())), 25.0,
                                      _np_ptr(grad_tex), _np_ptr(grad_op)) == 0
    assert np.array_equal(grad_tex, gtex.numpy())
    assert np.array_equal(grad_op, gop.numpy())


@pytest.mark.parametrize("count,off", [(4 * 700, 0), (4 * 700 + 3, 0), (256, 1)])
def test_emul_expand_views(count, off):
    """mvp_expand_views on the emulation: 16-byte path, scalar path for a count that is not a multiple of 4 and for a source
    that is only 4-byte aligned; untouched guard floats either side of the destination."""
    from tests.emul.build import build_aux
    emul = ctypes.CDLL(build_aux())                          # raydirs.cu + epilogue.cu on the CPU emulation
    emul.mvp_expand_views.argtypes = [F32P, F32P, ctypes.c_size_t, ctypes.c_int32, ctypes.c_void_p]
    n = 3
    raw = np.random.default_rng(count).standard_normal(count + 8).astype(np.float32)
    base = (-raw.ctypes.data // 4) % 4                       # index of a 16-byte aligned float
    src = raw[base + off:base + off + count]
    draw = np.full(n * count + 12, 7.0, np.float32)
    dbase = (-draw.ctypes.data // 4) % 4 + 4
    dst = draw[dbase:dbase + n * count]
    assert emul.mvp_expand_views(_np_ptr(src), _np_ptr(dst), count, n, None) == 0
    assert np.array_equal(dst.reshape(n, count), np.broadcast_to(src, (n, count)))
    assert (draw[:dbase] == 7.0).all() and (draw[dbase + n * count:] == 7.0).all()


@pytest.mark.parametrize("count,off,n", [(4 * 700, 0, 19), (4 * 700 + 3, 0, 3), (256, 1, 8), (64, 0, 1)])
def test_emul_sum_views(count, off, n):
    """mvp_sum_views on the emulation: dst = sum over the views in view order -- the 16-byte path (8 views per round + tail), the
    scalar path for a count that is not a multiple of 4 and for a 4-byte aligned destination; guard floats stay untouched."""
    from tests.emul.build import build_aux
    emul = ctypes.CDLL(build_aux())
    emul.mvp_sum_views.argtypes = [F32P, F32P, ctypes.c_size_t, ctypes.c_int32, ctypes.c_void_p]
    raw = np.random.default_rng(count + n).standard_normal(n * count + 8).astype(np.float32)
    base = (-raw.ctypes.data // 4) % 4
    src = raw[base:base + n * count]
    draw = np.full(count + 12, 7.0, np.float32)
    dbase = (-draw.ctypes.data // 4) % 4 + 4 + off
    dst = draw[dbase:dbase + count]
    assert emul.mvp_sum_views(_np_ptr(src), _np_ptr(dst), count, n, None) == 0
    ref = np.zeros(count, np.float32)
    for v in range(n):                                        # same order, same single roundings
        ref = (ref + src[v * count:(v + 1) * count]).astype(np.float32)
    assert np.array_equal(dst, ref)
    assert (draw[:dbase] == 7.0).all() and (draw[dbase + count:] == 7.0).all()
    assert emul.mvp_sum_views(_np_ptr(src), _np_ptr(dst), count, 0, None) == -2
    assert emul.mvp_sum_views(None, _np_ptr(dst), count, 1, None) == -1


def test_restatement_matches_reference_golden():
    """oracle/epilogue_ref.py against vectors produced with the reference's own Colorcal module / literal statements."""
    z = np.load(os.path.join(HERE, "golden", "epilogue_composite.npz"))
    for tag in ("a", "b"):
        t = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + "_")}
        rgb, alpha, (gr, gw, gb, gbg) = _composite_ref_grads(t["rayrgba"], t["ccw"], t["ccb"], t["bg"], t["g_rgb"], t["g_alpha"])
        assert torch.equal(rgb, t["irgbrec"]) and torch.equal(alpha, t["rayalpha"])
        torch.testing.assert_close(gr, t["grad_rayrgba"], rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(gbg, t["grad_bg"], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(gw, t["grad_ccw"], rtol=1e-5, atol=1e-2)
        torch.testing.assert_close(gb, t["grad_ccb"], rtol=1e-5, atol=1e-4)
    z = np.load(os.path.join(HERE, "golden", "epilogue_payload.npz"))
    for tag in ("a", "b"):
        B = int(z[tag + "_B"])
        tex, op, gt = (torch.from_numpy(z["%s_%s" % (tag, k)]) for k in ("tex", "opacity", "g_template"))
        ref, gtex, gop = _payload_ref_grads(tex, op, gt, B)
        assert np.array_equal(ref.numpy(), z[tag + "_template"])
        assert np.array_equal(gtex.numpy(), z[tag + "_grad_tex"]) and np.array_equal(gop.numpy(), z[tag + "_grad_opacity"])


def test_epilogue_argument_errors_do_not_need_a_device():
    from ava256_b200 import lib
    L = lib.LIB
    p = ctypes.c_void_p(4096)
    odd = ctypes.c_void_p(4100)
    assert L.mvp_composite_forward(1, 4, 4, None, None, None, None, p, None, None) == -1      # MVP_ERR_NULL
    assert L.mvp_composite_forward(1, 4, 4, p, p, None, None, p, None, None) == -1            # ccw without ccb
    assert L.mvp_composite_forward(0, 4, 4, p, None, None, None, p, None, None) == -2         # MVP_ERR_SHAPE
    assert L.mvp_composite_forward(1, 4, 40000, p, None, None, None, p, None, None) == -2
    assert L.mvp_composite_forward(1, 4, 4, odd, None, None, None, p, None, None) == -6       # MVP_ERR_ALIGN
    assert L.mvp_composite_backward(1, 4, 4, None, None, None, None, None, p, None, None, None, None) == -1
    assert L.mvp_composite_backward(1, 4, 4, None, None, None, p, None, p, p, None, None, None) == -1   # grad_ccw without grad_ccb
    assert L.mvp_composite_backward(1, 4, 4, None, None, None, p, None, p, p, p, None, None) == -1      # cc grads need rayrgba
    assert L.mvp_composite_backward(1, 4, 4, p, None, None, p, None, p, None, None, p, None) == -1      # grad_bg without bg
    assert L.mvp_composite_backward(1, 4, 4, None, None, None, p, None, odd, None, None, None, None) == -6
    assert L.mvp_assemble_payload_forward(1, 2, 2, 8, None, p, 25.0, 100.0, p, None) == -1
    assert L.mvp_assemble_payload_forward(1, 2, 2, 0, p, p, 25.0, 100.0, p, None) == -2
    assert L.mvp_assemble_payload_forward(1, 2, 2, 65, p, p, 25.0, 100.0, p, None) == -2
    assert L.mvp_assemble_payload_forward(1, 2, 2, 8, p, p, 25.0, 100.0, odd, None) == -6
    assert L.mvp_assemble_payload_backward(1, 2, 2, 8, p, p, 25.0, p, None, None) == -1
    assert L.mvp_assemble_payload_backward(1, 0, 2, 8, p, p, 25.0, p, p, None) == -2
    assert L.mvp_expand_views(None, p, 16, 2, None) == -1
    assert L.mvp_expand_views(p, p, 16, -1, None) == -2
    assert L.mvp_expand_views(p, p, 0, 2, None) == 0                                        # nothing to do, no launch
    assert b"aligned" in L.mvp_error_string(-6)


# ------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels through the Python mirror
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 6, 10), (2, 7, 9), (2, 64, 67), (1, 128, 96)])
@pytest.mark.parametrize("with_cc,with_bg", [(False, False), (True, False), (False, True), (True, True)])
def test_gpu_composite_matches_eager_torch(shape, with_cc, with_bg):
    from ava256_b200.composite import composite
    N, H, W = shape
    rayrgba, ccw, ccb, bg, g_rgb, g_alpha = _composite_case(N, H, W, with_cc, with_bg, seed=H, device="cuda")
    ref_rgb, ref_alpha, (gr, gw, gb, gbg) = _composite_ref_grads(rayrgba, ccw, ccb, bg, g_rgb, g_alpha)
    leaves = [None if t is None else t.detach().clone().requires_grad_(True) for t in (rayrgba, ccw, ccb, bg)]
    rgb, alpha = composite(*leaves)
    assert rgb.shape == (N, 3, H, W) and alpha.shape == (N, 1, H, W) and rgb.is_contiguous() and alpha.is_contiguous()
    assert torch.equal(rgb, ref_rgb) and torch.equal(alpha, ref_alpha)
    ((rgb * g_rgb).sum() + (alpha * g_alpha).sum()).backward()
    assert leaves[0].grad.is_contiguous()
    torch.testing.assert_close(leaves[0].grad, gr, rtol=1e-6, atol=1e-4)
    if with_cc:
        scale = float(H * W) ** 0.5
        torch.testing.assert_close(leaves[1].grad, gw, rtol=1e-4, atol=1e-2 * scale)
        torch.testing.assert_close(leaves[2].grad, gb, rtol=1e-4, atol=1e-4 * scale)
    if with_bg:
        torch.testing.assert_close(leaves[3].grad, gbg, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_gpu_composite_golden_broadcast_bg_and_raymarcher_mirror():
    from ava256_b200.composite import composite, split_rgba
    z = np.load(os.path.join(HERE, "golden", "epilogue_composite.npz"))
    for tag in ("a", "b"):
        t = {k[2:]: torch.from_numpy(z[k]).cuda() for k in z.files if k.startswith(tag + "_")}
        leaves = [t[k].clone().requires_grad_(True) for k in ("rayrgba", "ccw", "ccb", "bg")]
        rgb, alpha = composite(*leaves)
        assert torch.equal(rgb, t["irgbrec"]) and torch.equal(alpha, t["rayalpha"])
        ((rgb * t["g_rgb"]).sum() + (alpha * t["g_alpha"]).sum()).backward()
        torch.testing.assert_close(leaves[0].grad, t["grad_rayrgba"], rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(leaves[3].grad, t["grad_bg"], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(leaves[1].grad, t["grad_ccw"], rtol=1e-5, atol=1e-2)
        torch.testing.assert_close(leaves[2].grad, t["grad_ccb"], rtol=1e-5, atol=1e-4)
    # black background given as the [1,3,1,1] constant of models/autoencoder.py:268-270; gradient flows through the expand
    rayrgba = torch.rand(2, 8, 12, 4, device="cuda")
    colour = torch.tensor([10.0, 20.0, 30.0], device="cuda")[None, :, None, None].requires_grad_(True)
    rgb, alpha = composite(rayrgba, bg=colour)
    ref_rgb, _ = composite_ref(rayrgba, bg=colour)
    assert torch.equal(rgb, ref_rgb)
    rgb.sum().backward()
    torch.testing.assert_close(colour.grad.flatten(), (1.0 - rayrgba[..., 3]).sum().expand(3), rtol=1e-5, atol=1e-3)
    # split only
    rgb, alpha = split_rgba(rayrgba)
    assert torch.equal(rgb, rayrgba.permute(0, 3, 1, 2)[:, :3]) and torch.equal(alpha, rayrgba.permute(0, 3, 1, 2)[:, 3:4])


@pytest.mark.gpu
@pytest.mark.parametrize("N,hb,wb,B", [(2, 2, 3, 8), (1, 3, 2, 4), (1, 4, 2, 3), (1, 1, 1, 1), (1, 2, 2, 16), (2, 16, 16, 8)])
def test_gpu_payload_matches_eager_torch(N, hb, wb, B):
    from ava256_b200.payload import assemble_payload
    tex, opacity, gt = _payload_case(N, hb, wb, B, seed=B, device="cuda")
    ref, gtex, gop = _payload_ref_grads(tex, opacity, gt, B)
    t, o = tex.clone().requires_grad_(True), opacity.clone().requires_grad_(True)
    tp = assemble_payload(t, o, boxsize=B)
    assert tp.shape == (N, hb * wb, B, B, B, 4) and tp.is_contiguous()
    assert torch.equal(tp, ref)
    (tp * gt).sum().backward()
    assert torch.equal(t.grad, gtex) and torch.equal(o.grad, gop)


@pytest.mark.gpu
def test_gpu_payload_golden_and_feeds_the_raymarcher():
    from ava256_b200.payload import assemble_payload
    z = np.load(os.path.join(HERE, "golden", "epilogue_payload.npz"))
    for tag in ("a", "b"):
        B = int(z[tag + "_B"])
        tex = torch.from_numpy(z[tag + "_tex"]).cuda().requires_grad_(True)
        op = torch.from_numpy(z[tag + "_opacity"]).cuda().requires_grad_(True)
        tp = assemble_payload(tex, op, boxsize=B)
        assert np.array_equal(tp.detach().cpu().numpy(), z[tag + "_template"])
        (tp * torch.from_numpy(z[tag + "_g_template"]).cuda()).sum().backward()
        assert np.array_equal(tex.grad.cpu().numpy(), z[tag + "_grad_tex"])
        assert np.array_equal(op.grad.cpu().numpy(), z[tag + "_grad_opacity"])
    # end to end: decoder images -> payload kernel -> raymarch op -> composite kernel, gradients back to the images
    from ava256_b200.composite import composite
    from ava256_b200.op import mvpraymarch
    from tests.helpers import build_case
    case, _ = build_case("head_small")
    s = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in case.items()}
    N, K = s["primpos"].shape[:2]
    B = s["template"].shape[2]
    hb = wb = int(round(K ** 0.5))
    assert hb * wb == K and s["template"].shape[2:5] == (B, B, B)
    tex = torch.randn(N, 3 * B, hb * B, wb * B, device="cuda").requires_grad_(True)
    opa = (torch.randn(N, B, hb * B, wb * B, device="cuda") * 3.0).requires_grad_(True)
    tp = assemble_payload(tex, opa, boxsize=B)
    rgba = mvpraymarch(s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], (s["primpos"], s["primrot"], s["primscale"]), tp, None)
    rgb, alpha = composite(rgba, bg=torch.full((N, 3, rgba.size(1), rgba.size(2)), 7.0, device="cuda"))
    (rgb.sum() + alpha.sum()).backward()
    assert torch.isfinite(tex.grad).all() and torch.isfinite(opa.grad).all() and float(tex.grad.abs().sum()) > 0
    # same chain with the eager restatements around the same op
    tex2, opa2 = tex.detach().clone().requires_grad_(True), opa.detach().clone().requires_grad_(True)
    tp2 = assemble_payload_ref(tex2, opa2, B)
    rgba2 = mvpraymarch(s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], (s["primpos"], s["primrot"], s["primscale"]), tp2.contiguous(), None)
    rgb2, alpha2 = composite_ref(rgba2, bg=torch.full((N, 3, rgba.size(1), rgba.size(2)), 7.0, device="cuda"))
    (rgb2.sum() + alpha2.sum()).backward()
    assert torch.equal(rgb, rgb2) and torch.equal(alpha, alpha2)
    torch.testing.assert_close(tex.grad, tex2.grad, rtol=1e-3, atol=1e-3 * float(tex2.grad.abs().max()))
    torch.testing.assert_close(opa.grad, opa2.grad, rtol=1e-3, atol=1e-3 * float(opa2.grad.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n", [((5, 3), 4), ((7,), 3), ((16, 8, 8, 8, 4), 6), ((1, 1, 2), 1)])
def test_gpu_expand_views_equals_expand_contiguous(shape, n):
    """One subject's tensor materialised per view (`mvp_expand_views`): bit-identical to expand().contiguous(), vector and
    scalar (count % 4 != 0, unaligned slice) paths; the adjoint is the sum over the views."""
    from ava256_b200.payload import expand_views
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    y = expand_views(x, n)
    assert y.shape == (n,) + tuple(shape) and torch.equal(y, x[None].expand(n, *shape).contiguous())
    g = torch.randn_like(y)
    y.backward(g)
    assert torch.allclose(x.grad, g.sum(0), rtol=1e-6, atol=1e-6)
    flat = torch.randn(4 * 33 + 1, device="cuda")[1:]                                       # 4-byte aligned only
    assert torch.equal(expand_views(flat, 3), flat[None].expand(3, -1).contiguous())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n", [((5, 3), 4), ((64, 8, 8, 8, 4), 19), ((1024, 9), 10), ((7,), 1)])
def test_gpu_sum_views_equals_sequential_sum(shape, n):
    """`mvp_sum_views` (the local step of the per-subject gradient reduction, parallel.GradReducer): the sum over the views added
    in view order, bit for bit; into a fresh tensor and into a slice of a flat buffer (4-byte aligned: the scalar path)."""
    from ava256_b200.payload import sum_views
    x = torch.randn(n, *shape, device="cuda")
    ref = torch.zeros(*shape, device="cuda")
    for v in range(n):
        ref = ref + x[v]
    assert torch.equal(sum_views(x), ref)
    assert torch.allclose(sum_views(x), x.sum(0), rtol=1e-5, atol=1e-5)
    cnt = ref.numel()
    flat = torch.full((cnt + 9,), 7.0, device="cuda")
    sum_views(x, flat[5:5 + cnt])
    assert torch.equal(flat[5:5 + cnt].view(*shape), ref) and bool((flat[:5] == 7.0).all()) and bool((flat[5 + cnt:] == 7.0).all())
