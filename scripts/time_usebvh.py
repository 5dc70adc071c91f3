"""usebvh=True (Morton order as an indirection) vs fixed order on a few C3 views: fwd+bwd through the public op."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ava256_b200 import scene
from ava256_b200.op import mvpraymarch
nv = 8
s = scene.make_scene(nv, 1024, 667, 16384, 8, device="cuda", alpha_mu=17.0, alpha_sigma=6.0)
g = torch.randn(nv, 1024, 667, 4, device="cuda")
lv = [s[n].requires_grad_(True) for n in ("primpos", "primrot", "primscale", "template")]
def run(usebvh):
    for x in lv: x.grad = None
    out = mvpraymarch(s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], (lv[0], lv[1], lv[2]), lv[3], None, usebvh=usebvh)
    out.backward(g)
    return out
for mode in ("fixedorder", True):
    for _ in range(2): run(mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): out = run(mode)
    b.record(); torch.cuda.synchronize()
    print("usebvh=%-10s %.3f ms per fwd+bwd of %d views; peak memory %.2f GB" % (mode, a.elapsed_time(b) / 5, nv, torch.cuda.max_memory_allocated() / 1e9))
