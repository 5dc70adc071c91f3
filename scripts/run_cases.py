"""Runs named parity cases (tests/helpers.py CASES) forward + backward through the public op on cuda:0 and prints a checksum
per case -- the workload for compute-sanitizer (scripts/gpu_sanitize.sh) and quick smoke runs.  No oracle involved."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import build_case  # noqa: E402
from tests.test_gpu_parity import run_ours  # noqa: E402

def run_camera_head():
    """rays generated inside the kernels from the camera (mvp_camera): 2 views 64x42, K=64"""
    from ava256_b200 import scene
    from ava256_b200.op import mvpraymarch_camera
    n, H, W, K, T = 2, 64, 42, 64, 8
    cams = [c.cuda() for c in scene.make_cameras(n, H, W)]
    sc = scene.make_scene(n, H, W, K, T, device="cuda", alpha_mu=1.0, alpha_sigma=2.0, share_primitives=False)
    lv = [sc[k].clone().requires_grad_(True) for k in ("primpos", "primrot", "primscale", "template")]
    out = mvpraymarch_camera(cams[0], cams[1], cams[2], cams[3], (W, H), scene.VOLRADIUS, 1.0 / 64, (lv[0], lv[1], lv[2]), lv[3], None)
    out.backward(torch.randn(n, H, W, 4, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)))
    return out.detach(), [x.grad for x in lv]


for name in sys.argv[1:] or ["head_small", "many_overlaps", "warp_small", "camera_head"]:
    if name == "camera_head":
        out, grads = run_camera_head()
    else:
        s, grad = build_case(name)
        out, grads = run_ours(s, grad)
    torch.cuda.synchronize()
    print("%-16s rayrgba sum %.6e  grads %s" % (name, float(out.sum()), " ".join("%.4e" % float(abs(g).sum()) for g in grads)), flush=True)
