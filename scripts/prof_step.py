"""Small fwd+bwd driver for ncu captures: python scripts/prof_step.py [N H W K T iters]."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ava256_b200 import scene  # noqa: E402
from extensions.mvpraymarch.mvpraymarch import mvpraymarch  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
N, H, W, K, T, iters = (a + [2, 1024, 667, 16384, 8, 2][len(a):])[:6]
import os as _os
s = scene.make_scene(N, H, W, K, T, alpha_mu=float(_os.environ.get("ALPHA_MU", "3.0")), alpha_sigma=float(_os.environ.get("ALPHA_SIGMA", "3.0")), device="cuda")
leaves = [s[k].clone().requires_grad_(True) for k in ("primpos", "primrot", "primscale", "template")]
grad = torch.randn(N, H, W, 4, device="cuda")
for _ in range(iters):
    out = mvpraymarch(s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], (leaves[0], leaves[1], leaves[2]), leaves[3], None)
    out.backward(grad)
torch.cuda.synchronize()
print("done", float(out.mean()))
