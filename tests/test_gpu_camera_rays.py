"""GPU tests of the fused ray generation (C-ABI `mvp_camera`): the render kernels generate each tile's
rays from the camera parameters in their prologue -- the arithmetic of the reference's compute_raydirs kernel
(/root/reference/extensions/utils/utils_kernel.cu:32-46) -- instead of reading raypos / raydir / tminmax.  The gate is identity with the
two-call form `compute_raydirs(...)` -> `mvpraymarch(...)` (models/autoencoder.py:240-252): same image bits, same saturated rays,
gradients equal up to fp32 atomic order."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ("primpos", "primrot", "primscale", "template")


def _trelerr(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def _scene(n, H, W, K, T, view_offset=0, **kw):
    from ava256_b200 import scene
    cams = tuple(t.cuda() for t in scene.make_cameras(n, H, W, view_offset=view_offset))
    s = scene.make_scene(n, H, W, K, T, view_offset=view_offset, device="cuda", **kw)
    return cams, s


@pytest.mark.parametrize("n,H,W,K,T,mu,sg", [(2, 96, 70, 64, 8, 1.0, 2.0), (3, 61, 45, 256, 8, 6.0, 6.0), (1, 1024, 667, 16384, 8, 17.0, 6.0)])
def test_camera_rays_identical_to_two_call_form(n, H, W, K, T, mu, sg):
    from ava256_b200 import scene
    from ava256_b200.op import mvpraymarch_camera
    from extensions.mvpraymarch.mvpraymarch import mvpraymarch
    from extensions.utils.utils import compute_raydirs
    (viewpos, viewrot, focal, princpt), s = _scene(n, H, W, K, T, view_offset=2, alpha_mu=mu, alpha_sigma=sg)
    grad = torch.randn(n, H, W, 4, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    step = s["stepsize"] if H >= 512 else 1.0 / 64

    lv = [s[k].detach().clone().requires_grad_(True) for k in NAMES]
    rp, rd, tmm = compute_raydirs(viewpos, viewrot, focal, princpt, (W, H), scene.VOLRADIUS)
    out2 = mvpraymarch(rp, rd, step, tmm, (lv[0], lv[1], lv[2]), lv[3], None)
    out2.backward(grad)
    g2 = [x.grad for x in lv]

    lw = [s[k].detach().clone().requires_grad_(True) for k in NAMES]
    out1 = mvpraymarch_camera(viewpos, viewrot, focal, princpt, (W, H), scene.VOLRADIUS, step, (lw[0], lw[1], lw[2]), lw[3], None)
    out1.backward(grad)
    torch.cuda.synchronize()
    assert float(out2[..., 3].max()) > 0.05 and float((out2[..., 3] >= 1.0).float().mean()) > (0.005 if mu > 5 else 0.0)
    assert torch.equal(out1, out2)                                  # same rays, bit for bit -> same samples, same image
    for nm, a, b in zip(NAMES, [x.grad for x in lw], g2):
        assert bool(torch.isfinite(a).all()), nm
        assert _trelerr(a, b) <= 1e-5, nm

    with torch.no_grad():                                           # inference mode: the kGrad = false kernels
        out0 = mvpraymarch_camera(viewpos, viewrot, focal, princpt, (W, H), scene.VOLRADIUS, step,
                                  (s["primpos"], s["primrot"], s["primscale"]), s["template"], None)
    assert torch.equal(out0, out2.detach())


def test_raymarcher_forward_camera_matches_forward():
    from ava256_b200 import scene
    from ava256_b200.raymarcher import Raymarcher
    from extensions.utils.utils import compute_raydirs
    n, H, W, K, T = 2, 80, 56, 64, 8
    (viewpos, viewrot, focal, princpt), s = _scene(n, H, W, K, T, alpha_mu=2.0, alpha_sigma=2.0)
    rm = Raymarcher(scene.VOLRADIUS, dt=4.0)
    dec = {k: s[k].detach().clone().requires_grad_(True) for k in NAMES}
    dec2 = {k: s[k].detach().clone().requires_grad_(True) for k in NAMES}
    rp, rd, tmm = compute_raydirs(viewpos, viewrot, focal, princpt, (W, H), scene.VOLRADIUS)
    rgb_a, alpha_a, _, _ = rm(rp, rd, tmm, dec)
    rgb_b, alpha_b, _, _ = rm.forward_camera(viewpos, viewrot, focal, princpt, (W, H), dec2)
    assert torch.equal(rgb_a, rgb_b) and torch.equal(alpha_a, alpha_b)
    g_rgb, g_a = torch.randn_like(rgb_a), torch.randn_like(alpha_a)
    torch.autograd.backward([rgb_a, alpha_a], [g_rgb, g_a])
    torch.autograd.backward([rgb_b, alpha_b], [g_rgb, g_a])
    for k in NAMES:
        assert _trelerr(dec2[k].grad, dec[k].grad) <= 1e-5, k
    # an explicit pixelcoords tensor takes the two-call path and gives the same image
    py, px = torch.meshgrid(torch.arange(H, device="cuda").float(), torch.arange(W, device="cuda").float(), indexing="ij")
    pc = torch.stack([px, py], dim=-1)[None].repeat(n, 1, 1, 1).contiguous()
    with torch.no_grad():
        rgb_c, alpha_c, _, _ = rm.forward_camera(viewpos, viewrot, focal, princpt, pc, dec)
    assert torch.equal(rgb_c, rgb_a.detach()) and torch.equal(alpha_c, alpha_a.detach())


def test_build_accel_camera_then_forward_with_valid_accel():
    """C-ABI: mvp_build_accel_camera + forward(MVP_FLAG_ACCEL_VALID) == forward that builds the accel itself; a degenerate camera
    (focal 0) is reported per view and the view takes the all-slabs fallback instead of faulting."""
    from ava256_b200 import lib, scene
    n, H, W, K, T = 2, 64, 48, 64, 8
    (viewpos, viewrot, focal, princpt), s = _scene(n, H, W, K, T, alpha_mu=2.0, alpha_sigma=2.0)
    P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    shape = lib.Shape(n, H, W, K, T, T, T)
    wsb = lib.workspace_bytes(n, H, W, K, T, T, T)

    def run(prebuilt, foc):
        ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
        cam = lib.Camera(P(viewpos), P(viewrot), P(foc), P(princpt), scene.VOLRADIUS, 0)
        flags = 0
        if prebuilt:
            lib.check(lib.LIB.mvp_build_accel_camera(ctypes.byref(shape), 0, None, ctypes.byref(cam), P(s["primpos"]), P(s["primrot"]),
                                                     P(s["primscale"]), P(ws), wsb, st))
            flags = lib.FLAG_ACCEL_VALID
        rgba = torch.full((n, H, W, 4), float("nan"), device="cuda")
        a = lib.ForwardArgs()
        a.shape, a.stepsize, a.fadescale, a.fadeexp, a.flags = shape, 1.0 / 64, 8.0, 8.0, flags
        a.camera = cam
        a.primpos, a.primrot, a.primscale, a.tplate = P(s["primpos"]), P(s["primrot"]), P(s["primscale"]), P(s["template"])
        a.rayrgba, a.workspace, a.workspace_bytes = P(rgba), P(ws), wsb
        lib.check(lib.LIB.mvp_raymarch_forward(ctypes.byref(a), st))
        torch.cuda.synchronize()
        return rgba

    a0, a1 = run(False, focal), run(True, focal)
    assert float(a0[..., 3].max()) > 0.05 and torch.equal(a0, a1)
    bad_focal = focal.clone()
    bad_focal[1] = 0.0
    b = run(False, bad_focal)
    assert torch.equal(b[0], a0[0])                                  # the healthy view is untouched; view 1 has no meaningful rays
