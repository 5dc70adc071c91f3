"""TEST INFRASTRUCTURE -- eager-PyTorch restatement of the expressions the epilogue kernels replace
(SURVEY.md section 8f rows 2 and 4).  Only tests/ may import this module; it is never on the product path.

Each function follows the reference line by line (same ops in the same order, so the same roundings) and runs on
whatever device its inputs carry; autograd through it defines the gradients.
"""
import torch
import torch.nn.functional as F


def composite_ref(rayrgba, ccw=None, ccb=None, bg=None):
    """rayrgba [N,H,W,4] -> (irgbrec [N,3,H,W], rayalpha [N,1,H,W])."""
    r = rayrgba.permute(0, 3, 1, 2)                                       # models/raymarchers/mvpraymarcher.py:50
    rayrgb, rayalpha = r[:, :3].contiguous(), r[:, 3:4].contiguous()      # :51
    if ccw is not None:
        rayrgb = ccw.unsqueeze(-1).unsqueeze(-1) * rayrgb + ccb.unsqueeze(-1).unsqueeze(-1)   # models/colorcals/colorcal.py:29
    if bg is not None:
        rayrgb = rayrgb + (1.0 - rayalpha) * bg                           # models/autoencoder.py:263
    return rayrgb, rayalpha


def assemble_payload_ref(tex, opacity, boxsize, rgb_scale=25.0, rgb_bias=100.0):
    """tex [N, B*3, h*B, w*B], opacity [N, B, h*B, w*B] -> template [N, h*w, B, B, B, 4]."""
    B = boxsize
    N = tex.size(0)
    h, w = tex.size(2) // B, tex.size(3) // B
    rgb = tex.view(N, B, 3, h, B, w, B)                                   # models/decoders/rgb.py:137
    rgb = rgb.permute(0, 3, 5, 1, 4, 6, 2)                                # :140
    rgb = rgb.reshape(N, h * w, B, B, B, 3)                               # :143
    op = opacity.view(N, B, 1, h, B, w, B)                                # models/decoders/geometry.py:182
    op = op.permute(0, 3, 5, 1, 4, 6, 2)                                  # :183
    op = op.reshape(N, h * w, B, B, B, 1)                                 # :184
    return torch.cat([F.relu(rgb * rgb_scale + rgb_bias), F.relu(op)], dim=-1)   # models/decoders/assembler.py:261
