"""Times the epilogue kernels at the C3 shapes against the HBM roofline and against
the eager PyTorch chains they replace.  CUDA events on the current stream, inputs far larger than L2.

  python scripts/time_epilogue.py [--views 80] [--payload-views 8] [--out gpurun_out/epilogue_timing.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ava256_b200.composite import composite  # noqa: E402
from ava256_b200.payload import assemble_payload  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def eager_composite(rayrgba, w, b, bg):
    r = rayrgba.permute(0, 3, 1, 2)
    rgb, alpha = r[:, :3].contiguous(), r[:, 3:4].contiguous()
    rgb = w.unsqueeze(-1).unsqueeze(-1) * rgb + b.unsqueeze(-1).unsqueeze(-1)
    return rgb + (1.0 - alpha) * bg, alpha


def eager_payload(tex, opacity, B):
    N, h, w = tex.size(0), tex.size(2) // B, tex.size(3) // B
    rgb = tex.view(N, B, 3, h, B, w, B).permute(0, 3, 5, 1, 4, 6, 2).reshape(N, h * w, B, B, B, 3)
    op = opacity.view(N, B, 1, h, B, w, B).permute(0, 3, 5, 1, 4, 6, 2).reshape(N, h * w, B, B, B, 1)
    return torch.cat([torch.relu(rgb * 25.0 + 100.0), torch.relu(op)], dim=-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=80)
    ap.add_argument("--payload-views", type=int, default=8)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peak = 6570.3
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = {"peak_gbps": peak}
    dev = "cuda"
    N, H, W = args.views, 1024, 667
    rayrgba = torch.rand(N, H, W, 4, device=dev).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(N, 3, device=dev)).requires_grad_(True)
    b = torch.randn(N, 3, device=dev).requires_grad_(True)
    bg = torch.rand(N, 3, H, W, device=dev).requires_grad_(True)
    px = N * H * W
    with torch.no_grad():
        t_f = timeit(lambda: composite(rayrgba, w, b, bg))
        t_fe = timeit(lambda: eager_composite(rayrgba, w, b, bg))
    rgb, alpha = composite(rayrgba, w, b, bg)
    g1, g2 = torch.randn_like(rgb), torch.randn_like(alpha)
    t_b = timeit(lambda: torch.autograd.grad((rgb, alpha), (rayrgba, w, b, bg), (g1, g2), retain_graph=True))
    rgb_e, alpha_e = eager_composite(rayrgba, w, b, bg)
    t_be = timeit(lambda: torch.autograd.grad((rgb_e, alpha_e), (rayrgba, w, b, bg), (g1, g2), retain_graph=True))
    res["composite"] = {
        "shape": [N, H, W], "fwd_ms": t_f, "fwd_bytes_per_ray": 44, "fwd_gbps": px * 44 / t_f / 1e6, "fwd_frac": px * 44 / t_f / 1e6 / peak,
        "eager_fwd_ms": t_fe, "bwd_ms": t_b, "bwd_bytes_per_ray": 72, "bwd_gbps": px * 72 / t_b / 1e6,
        "bwd_frac": px * 72 / t_b / 1e6 / peak, "eager_bwd_ms": t_be,
        "note": "bwd_ms includes torch's zero-fill of grad_ccw/grad_ccb and autograd dispatch; with colour calibration and a background image",
    }
    del rayrgba, bg, rgb, alpha, rgb_e, alpha_e, g1, g2
    torch.cuda.empty_cache()
    Np, B, hb = args.payload_views, 8, 128
    tex = torch.randn(Np, 3 * B, hb * B, hb * B, device=dev).requires_grad_(True)
    opa = torch.randn(Np, B, hb * B, hb * B, device=dev).requires_grad_(True)
    texels = Np * hb * hb * B ** 3
    with torch.no_grad():
        t_f = timeit(lambda: assemble_payload(tex, opa, B))
        t_fe = timeit(lambda: eager_payload(tex, opa, B))
    tp = assemble_payload(tex, opa, B)
    gt = torch.randn_like(tp)
    t_b = timeit(lambda: torch.autograd.grad(tp, (tex, opa), gt, retain_graph=True))
    del tp
    tpe = eager_payload(tex, opa, B)
    t_be = timeit(lambda: torch.autograd.grad(tpe, (tex, opa), gt, retain_graph=True))
    res["payload"] = {
        "shape": [Np, hb, hb, B], "fwd_ms": t_f, "fwd_bytes_per_texel": 32, "fwd_gbps": texels * 32 / t_f / 1e6,
        "fwd_frac": texels * 32 / t_f / 1e6 / peak, "eager_fwd_ms": t_fe, "bwd_ms": t_b, "bwd_bytes_per_texel": 48,
        "bwd_gbps": texels * 48 / t_b / 1e6, "bwd_frac": texels * 48 / t_b / 1e6 / peak, "eager_bwd_ms": t_be,
    }
    line = json.dumps(res)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
