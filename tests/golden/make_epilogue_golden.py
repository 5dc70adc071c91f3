"""Generates tests/golden/epilogue_*.npz by running the REFERENCE's own modules on the CPU in the build container
(/root/reference does not exist on the GPU box, so the vectors are committed).

  composite: models/colorcals/colorcal.py `Colorcal.forward` (imported, unmodified) for the colour calibration, then the
             matting line of models/autoencoder.py:263 applied to its output; the NHWC->NCHW split is
             models/raymarchers/mvpraymarcher.py:50-51.
  payload:   the decoder modules cannot be instantiated without the asset files, so the vectors come from the three
             literal statements rgb.py:137-143 / geometry.py:182-184 / assembler.py:261 executed here on seeded inputs.

Run:  python tests/golden/make_epilogue_golden.py
"""
import importlib.util
import os

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_colorcal():
    spec = importlib.util.spec_from_file_location("ref_colorcal", os.path.join(REF, "models/colorcals/colorcal.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.Colorcal


def main():
    g = torch.Generator().manual_seed(1112)
    # ---- composite: N=3 views, 10x14 image (HW % 4 == 0) and 7x9 (scalar path) ----
    out = {}
    Colorcal = load_colorcal()
    for tag, (N, H, W) in (("a", (3, 10, 14)), ("b", (2, 7, 9))):
        cc = Colorcal(ncams=5, nident=4)
        with torch.no_grad():
            for p in (cc.wcam, cc.bcam, cc.wident, cc.bident):
                p.add_(torch.randn(p.shape, generator=g) * 0.3)
        camindex = torch.randint(0, 5, (N,), generator=g)
        idindex = torch.randint(0, 4, (N,), generator=g)
        rayrgba = torch.rand(N, H, W, 4, generator=g)
        rayrgba[..., :3] *= 255.0
        rayrgba.requires_grad_(True)
        bg = (torch.rand(N, 3, H, W, generator=g) * 255.0).requires_grad_(True)
        r = rayrgba.permute(0, 3, 1, 2)
        rayrgb, rayalpha = r[:, :3].contiguous(), r[:, 3:4].contiguous()
        rayrgb = cc(rayrgb, camindex, idindex)                      # reference module
        irgbrec = rayrgb + (1.0 - rayalpha) * bg
        g_rgb = torch.randn(irgbrec.shape, generator=g)
        g_alpha = torch.randn(rayalpha.shape, generator=g)
        (irgbrec * g_rgb).sum().add((rayalpha * g_alpha).sum()).backward()
        w = (cc.wcam[camindex] + cc.wident[idindex]).detach()
        b = (cc.bcam[camindex] + cc.bident[idindex]).detach()
        # d/dw, d/db per sample: scatter the parameter gradients back is ambiguous with repeated indices, so store the
        # per-sample sums computed from the definition in float64
        gw = (rayrgba.detach()[..., :3].permute(0, 3, 1, 2).double() * g_rgb.double()).sum(dim=(2, 3)).float()
        gb = g_rgb.double().sum(dim=(2, 3)).float()
        for k, v in dict(rayrgba=rayrgba, bg=bg, ccw=w, ccb=b, irgbrec=irgbrec, rayalpha=rayalpha, g_rgb=g_rgb, g_alpha=g_alpha,
                         grad_rayrgba=rayrgba.grad, grad_bg=bg.grad, grad_ccw=gw, grad_ccb=gb).items():
            out["%s_%s" % (tag, k)] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "epilogue_composite.npz"), **out)

    # ---- payload: B=8 (vector path), B=3 (scalar), h != w in both ----
    out = {}
    for tag, (N, h, w, B) in (("a", (2, 2, 3, 8)), ("b", (1, 4, 2, 3))):
        tex = (torch.randn(N, B * 3, h * B, w * B, generator=g) * 4.0 - 2.0).requires_grad_(True)
        opacity = torch.randn(N, B, h * B, w * B, generator=g).requires_grad_(True)
        rgb = tex.view(N, B, 3, h, B, w, B).permute(0, 3, 5, 1, 4, 6, 2).reshape(N, h * w, B, B, B, 3)
        op = opacity.view(N, B, 1, h, B, w, B).permute(0, 3, 5, 1, 4, 6, 2).reshape(N, h * w, B, B, B, 1)
        template = torch.cat([F.relu(rgb * 25.0 + 100.0), F.relu(op)], dim=-1)
        gt = torch.randn(template.shape, generator=g)
        (template * gt).sum().backward()
        for k, v in dict(tex=tex, opacity=opacity, template=template, g_template=gt, grad_tex=tex.grad,
                         grad_opacity=opacity.grad).items():
            out["%s_%s" % (tag, k)] = v.detach().numpy()
        out["%s_B" % tag] = np.int32(B)
    np.savez_compressed(os.path.join(HERE, "epilogue_payload.npz"), **out)
    print("wrote", [f for f in os.listdir(HERE) if f.startswith("epilogue_")])


if __name__ == "__main__":
    main()
