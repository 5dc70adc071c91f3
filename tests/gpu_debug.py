"""First-contact GPU script: prints parity numbers (no asserts) for every named case vs oracle and vs the reference
extension, then a rough timing on a mid-size head scene."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tests import refext  # noqa: E402
from tests.helpers import CASES, build_case, relerr, scene_args_np  # noqa: E402
from tests.test_gpu_parity import run_ours  # noqa: E402

names = ("primpos", "primrot", "primscale", "template", "warp")
for name in CASES:
    s, grad = build_case(name)
    out, grads = run_ours(s, grad)
    a, kw = scene_args_np(s)
    ref, raysat = oracle.forward(*a, **kw)
    gref = oracle.backward(*a, grad.numpy(), raysat, **kw)
    line = "%-18s vs oracle: fwd %.2e" % (name, relerr(out, ref))
    for nm, g, r in zip(names, grads, gref):
        line += " | %s %.2e" % (nm, relerr(g, r))
    print(line, flush=True)
    if refext.available():
        t = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in s.items()}
        fs, fe = s.get("fadescale", 8.0), s.get("fadeexp", 8.0)
        rgba, sat, st = refext.forward(t["raypos"], t["raydir"], t["stepsize"], t["tminmax"], t["primpos"], t["primrot"],
                                       t["primscale"], t["template"], fs, fe, warp=t.get("warp"))
        g2 = refext.backward(t["raypos"], t["raydir"], t["stepsize"], t["tminmax"], t["primpos"], t["primrot"],
                             t["primscale"], t["template"], rgba, sat, st, grad.cuda(), fs, fe, warp=t.get("warp"))
        line = "%-18s vs refext: fwd %.2e" % (name, relerr(out, rgba.cpu().numpy()))
        for nm, g, r in zip(names, grads, g2):
            line += " | %s %.2e" % (nm, relerr(g, r.cpu().numpy()))
        line += " || oracle vs refext fwd %.2e" % relerr(ref, rgba.cpu().numpy())
        for nm, g, r in zip(names, gref, g2):
            line += " %.2e" % relerr(g, r.cpu().numpy())
        print(line, flush=True)

# timing, mid-size
from ava256_b200 import scene  # noqa: E402
from extensions.mvpraymarch.mvpraymarch import mvpraymarch  # noqa: E402

for (N, H, W, K, T) in ((4, 512, 334, 4096, 16), (4, 1024, 667, 16384, 8)):
    s = scene.make_scene(N, H, W, K, T, alpha_mu=3.0, alpha_sigma=3.0, device="cuda")
    leaves = [s[k].clone().requires_grad_(True) for k in ("primpos", "primrot", "primscale", "template")]
    grad = torch.randn(N, H, W, 4, device="cuda")

    def step():
        out = mvpraymarch(s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], (leaves[0], leaves[1], leaves[2]), leaves[3], None)
        return out

    for _ in range(2):
        o = step(); o.backward(grad)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); o = step(); e1.record(); o.backward(grad); e2.record(); torch.cuda.synchronize()
    print("ours  N=%d %dx%d K=%d T=%d: fwd %.3f ms, bwd %.3f ms; sat frac %.3f cover %.3f" % (
        N, H, W, K, T, e0.elapsed_time(e1), e1.elapsed_time(e2), float((o[..., 3] >= 0.999).float().mean()),
        float((o[..., 3] > 0).float().mean())), flush=True)
    if refext.available():
        t0 = time.time()
        args = (s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], s["primpos"], s["primrot"], s["primscale"], s["template"])
        rgba, sat, st = refext.forward(*args)
        torch.cuda.synchronize(); t1 = time.time()
        rgba, sat, st = refext.forward(*args)
        torch.cuda.synchronize(); t2 = time.time()
        g2 = refext.backward(*args, rgba, sat, st, grad)
        torch.cuda.synchronize(); t3 = time.time()
        print("refext same: fwd %.3f ms (first %.3f), bwd %.3f ms; parity fwd %.2e" % (
            (t2 - t1) * 1e3, (t1 - t0) * 1e3, (t3 - t2) * 1e3, relerr(o.detach().cpu().numpy(), rgba.cpu().numpy())), flush=True)
        for nm, g, r in zip(names, [x.grad for x in leaves], g2):
            print("   grad %s relerr %.2e" % (nm, relerr(g.cpu().numpy() / 3.0, r.cpu().numpy())), flush=True)
