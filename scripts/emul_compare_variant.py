"""Design study on the CPU emulation (tests/emul): renders several full-size views with the default build of the kernels and
with a build-time variant and checks that the images (and gradients) are the same.  No GPU needed.

  python scripts/emul_compare_variant.py [DEFINE ...]        (default variant: MVP_LIST_MARGIN=1)
"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ava256_b200 import scene
from tests.emul import kernels
VARIANT = tuple(sys.argv[1:]) or ("MVP_LIST_MARGIN=1",)


def run(tag, s, grad=False):
    args = [s[k].numpy() if hasattr(s[k], "numpy") else s[k] for k in ("raypos","raydir","stepsize","tminmax","primpos","primrot","primscale","template")]
    g = np.random.default_rng(0).standard_normal(s["raypos"].shape[:3] + (4,)).astype(np.float32) if grad else None
    res = []
    for defs in ((), VARIANT):
        kernels.use_variant(defs)
        t0 = time.time(); out, sat, gr = kernels.forward_backward(*args, grad_rayrgba=g); res.append((out, sat, gr, time.time()-t0))
    same = np.array_equal(res[0][0], res[1][0])
    msg = "%s: image identical=%s (%.1fs/%.1fs) sat frac %.3f" % (tag, same, res[0][3], res[1][3], float((res[0][0][...,3] >= 0.999).mean()))
    if grad:
        rel = [float(np.abs(a-b).max()/max(np.abs(a).max(),1e-30)) for a,b in zip(res[0][2], res[1][2])]
        msg += " raysat identical=%s grad relerr %s" % (np.array_equal(res[0][1], res[1][1]), ["%.1e" % r for r in rel])
    print(msg, flush=True)
# other C3 views
for vo in (7, 23, 41, 66):
    run("C3 view %d" % vo, scene.make_scene(1, 1024, 667, 16384, 8, seed=1112, view_offset=vo, alpha_mu=17.0, alpha_sigma=6.0))
# low-alpha (no saturation: rays march through everything)
run("C3 view 3 low alpha", scene.make_scene(1, 1024, 667, 16384, 8, seed=1112, view_offset=3, alpha_mu=1.0, alpha_sigma=1.0))
# C2-like
run("C2 view", scene.make_scene(1, 512, 334, 4096, 16, seed=1112, view_offset=5, alpha_mu=8.0, alpha_sigma=4.0))
# fwd+bwd on a quarter-res C3-like view
run("C3/4 fwd+bwd", scene.make_scene(1, 512, 334, 4096, 8, seed=1112, view_offset=11, alpha_mu=17.0, alpha_sigma=6.0), grad=True)
# small dt (many steps: larger drift bound)
s = scene.make_scene(1, 256, 167, 1024, 8, seed=1112, view_offset=2, alpha_mu=4.0, alpha_sigma=2.0); s["stepsize"] = 1.0/2048
run("small dt 1/2048", s, grad=True)
