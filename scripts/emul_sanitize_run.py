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
