"""Driver of scripts/emul_sanitize.sh: runs the named parity scenes and the edge cases through a sanitizer build of the
emulated kernels (no comparisons here -- the sanitizers are the check)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.emul import kernels, build
build.build_kernels = lambda *a, **k: os.environ['MVP_EMUL_SANITIZED_LIB']
kernels.build_kernels = build.build_kernels
from tests.helpers import build_case, scene_args_np, relerr, edge_scene, EDGE_KINDS
import torch
for name in ("tiny", "gradcheck_ragged", "head_small", "many_overlaps", "warp_small", "noncubic", "head_t16"):
    s, grad = build_case(name); a, kw = scene_args_np(s)
    out, sat, g = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    print(name, "ok", float(np.abs(out).max()), flush=True)
for kind in EDGE_KINDS:
    s = edge_scene(kind); a, kw = scene_args_np(s)
    grad = torch.randn(*s["raypos"].shape[:3], 4)
    out, sat, g = kernels.forward_backward(*a, grad_rayrgba=grad.numpy(), **kw)
    print(kind, "ok", flush=True)
# rays generated in the kernels from the camera (mvp_camera), incl. a ragged image and a degenerate camera
from ava256_b200 import scene
for (n, H, W, K, T) in ((2, 64, 42, 64, 8), (1, 13, 19, 16, 4)):
    cams = [c.numpy() for c in scene.make_cameras(n, H, W)]
    sc = scene.make_scene(n, H, W, K, T, alpha_mu=1.0, alpha_sigma=2.0, share_primitives=False)
    prim = (sc["primpos"].numpy(), sc["primrot"].numpy(), sc["primscale"].numpy(), sc["template"].numpy())
    grad = torch.randn(n, H, W, 4).numpy()
    for bad in (False, True):
        f = cams[2].copy()
        if bad:
            f[-1] = 0.0
        out, sat, g = kernels.forward_backward(None, None, 1.0 / 64, None, *prim, grad_rayrgba=grad,
                                               camera=(cams[0], cams[1], f, cams[3], scene.VOLRADIUS, H, W), clear_in_forward=True)
        print("camera %dx%d%s" % (H, W, " (one degenerate view)" if bad else ""), "ok", float(np.abs(out).max()), flush=True)
