"""Times fwd / bwd kernels (C-ABI, accel prebuilt) for one config; used to A/B experiment knobs via env vars."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ava256_b200 import lib, scene  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
N, H, W, K, T = (a + [8, 1024, 667, 16384, 8][len(a):])[:5]
mu = float(os.environ.get("ALPHA_MU", "3.0")); sg = float(os.environ.get("ALPHA_SIGMA", "3.0"))
s = scene.make_scene(N, H, W, K, T, alpha_mu=mu, alpha_sigma=sg, device="cuda")
dev = "cuda"
wsb = lib.workspace_bytes(N, H, W, K, T, T, T)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
rgba = torch.empty(N, H, W, 4, device=dev); rsat = torch.empty(N, H, W, 3, device=dev)
raux = torch.empty(N, H, W, 4, dtype=torch.int32, device=dev)
grad = torch.randn(N, H, W, 4, device=dev)
P = lambda x: ctypes.c_void_p(x.data_ptr())
fa = lib.ForwardArgs(); fa.shape = lib.Shape(N, H, W, K, T, T, T)
fa.stepsize, fa.fadescale, fa.fadeexp, fa.flags = s["stepsize"], 8.0, 8.0, 0
fa.raypos, fa.raydir, fa.tminmax = P(s["raypos"]), P(s["raydir"]), P(s["tminmax"])
CAMERA = os.environ.get("CAMERA", "0") == "1"     # rays generated in the kernels from the camera parameters (mvp_camera)
if CAMERA:
    cams = [c.cuda() for c in scene.make_cameras(N, H, W)]
    fa.camera = lib.Camera(P(cams[0]), P(cams[1]), P(cams[2]), P(cams[3]), scene.VOLRADIUS, 0)
    fa.raypos = fa.raydir = fa.tminmax = None
fa.primpos, fa.primrot, fa.primscale, fa.tplate = P(s["primpos"]), P(s["primrot"]), P(s["primscale"]), P(s["template"])
fa.rayrgba, fa.raysat, fa.rayaux, fa.workspace, fa.workspace_bytes = P(rgba), P(rsat), P(raux), P(ws), wsb
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.check(lib.LIB.mvp_raymarch_forward(ctypes.byref(fa), st))
fa.flags = 1
gs = [torch.zeros_like(s[k]) for k in ("primpos", "primrot", "primscale", "template")]
ba = lib.BackwardArgs(); ba.shape = fa.shape
ba.stepsize, ba.fadescale, ba.fadeexp, ba.flags = s["stepsize"], 8.0, 8.0, 1
ba.raypos, ba.raydir, ba.tminmax = fa.raypos, fa.raydir, fa.tminmax
if CAMERA:
    ba.camera = fa.camera
ba.primpos, ba.primrot, ba.primscale, ba.tplate = fa.primpos, fa.primrot, fa.primscale, fa.tplate
ba.grad_rayrgba, ba.raysat, ba.rayaux = P(grad), P(rsat), P(raux)
ba.grad_primpos, ba.grad_primrot, ba.grad_primscale, ba.grad_tplate = P(gs[0]), P(gs[1]), P(gs[2]), P(gs[3])
ba.workspace, ba.workspace_bytes = P(ws), wsb


def tm(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


f = tm(lambda: lib.check(lib.LIB.mvp_raymarch_forward(ctypes.byref(fa), st)))
b = tm(lambda: lib.check(lib.LIB.mvp_raymarch_backward(ctypes.byref(ba), st)))
flagged = int(ws[wsb - 1 - 0:wsb].sum()) if False else -1
print("CAMERA=%d ALIGN=%s mu=%.1f N=%d %dx%d K=%d T=%d: fwd %.3f ms (%.3f/view)  bwd %.3f ms (%.3f/view)  sat %.3f cover %.3f" % (
    int(CAMERA), os.environ.get("MVP_ALIGN", "default"), mu, N, H, W, K, T, f, f / N, b, b / N, float((rgba[..., 3] >= 0.999).float().mean()),
    float((rgba[..., 3] > 0).float().mean())), flush=True)
