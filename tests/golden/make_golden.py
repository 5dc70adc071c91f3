"""Generates tests/golden/*.npz ON THE GPU BOX from the unmodified reference CUDA extension (oracle/_ref).

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   then copy gpurun_out/golden/*.npz here.

Each file holds the seeded inputs' recipe name plus the reference outputs: rayrgba, raysat and the four gradients
for grad_rayrgba drawn from a seeded generator.  Inputs are re-created from tests/helpers.py (same seeds), so only
outputs are stored.  These vectors pin oracle/mvp_oracle.c (CPU test) and the CUDA path (GPU test) to the reference
itself -- the reference's own test-suite pins no number for this path."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import refext  # noqa: E402
from tests.helpers import CASES, build_case  # noqa: E402


def main(outdir, only=()):
    os.makedirs(outdir, exist_ok=True)
    for name in CASES:
        if only and name not in only:
            continue
        s, grad = build_case(name)
        dev = "cuda"
        t = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in s.items()}
        fs, fe = s.get("fadescale", 8.0), s.get("fadeexp", 8.0)
        w = t.get("warp")
        rgba, sat, st = refext.forward(t["raypos"], t["raydir"], t["stepsize"], t["tminmax"], t["primpos"], t["primrot"],
                                       t["primscale"], t["template"], fs, fe, warp=w)
        g = refext.backward(t["raypos"], t["raydir"], t["stepsize"], t["tminmax"], t["primpos"], t["primrot"],
                            t["primscale"], t["template"], rgba, sat, st, grad.to(dev), fs, fe, warp=w)
        extra = {"grad_warp": g[4].cpu().numpy()} if w is not None else {}
        np.savez_compressed(os.path.join(outdir, name + ".npz"), rayrgba=rgba.cpu().numpy(), raysat=sat.cpu().numpy(),
                            grad_primpos=g[0].cpu().numpy(), grad_primrot=g[1].cpu().numpy(),
                            grad_primscale=g[2].cpu().numpy(), grad_template=g[3].cpu().numpy(), **extra)
        print(name, "saved; saturated rays:", int((sat[..., 0] > -1).sum()), "of", sat[..., 0].numel())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"), tuple(sys.argv[2:]))   # [outdir [case ...]]
