"""TEST INFRASTRUCTURE -- pure-PyTorch restatement of the reference's autograd raymarch loop.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It is never on the product path.

Follows /root/reference/extensions/mvpraymarch/mvpraymarch.py:567-633 (the `gradcheck` "python raymarching
implementation"), which hard-codes .to("cuda"); here the same loop runs on whatever device/dtype its inputs
carry, so it is (a) the CPU autograd baseline the north star names ("the reference's own CPU autograd path")
and (b) the independent definition of forward semantics and -- via autograd -- of all four gradients that
pins oracle/mvp_oracle.c in fp64.

Differences between that loop and the CUDA kernels, all on measure-zero boundaries:
inclusive [-1,1] validity (:604-606) vs strict (primtransf.h:112-117); non-incremental positions
raypos0 + raydir*dt*step (:623-624); every ray is marched over [tmin, tmax) (:579, 608).  `lattice=True`
switches the start of the march to the CUDA kernels' per-ray lattice origin only in the sense that both use
t_j = tmin + j*dt (already true), so no option is needed for that.

Template layout: this function takes the op's channels-last template [N,K,TD,TH,TW,4] and permutes it to the
channels-first layout F.grid_sample needs (:600).
"""
import torch
import torch.nn.functional as F


def raymarch_torch(raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template,
                   fadescale=8.0, fadeexp=8.0, max_steps=None, warp=None):
    """Returns rayrgba [N,H,W,4].  Differentiable w.r.t. primpos, primrot, primscale, template.

    raypos, raydir: [N,H,W,3]; tminmax: [N,H,W,2]; primpos/primscale: [N,K,3]; primrot: [N,K,3,3];
    template: [N,K,TD,TH,TW,4] (channels-last, as the op takes it).
    """
    N, H, W = raypos.shape[:3]
    K = primpos.shape[1]
    tplate = template.permute(0, 1, 5, 2, 3, 4)  # [N,K,4,TD,TH,TW]  (mvpraymarch.py:600 expects ch-first)
    wfield = warp.permute(0, 1, 5, 2, 3, 4) if warp is not None else None   # [N,K,3,WD,WH,WW]  (:594-597, dowarp)

    rayrgba = torch.zeros((N, H, W, 4), dtype=raypos.dtype, device=raypos.device)
    raypos_t = raypos + raydir * tminmax[:, :, :, 0, None]          # :571
    t = tminmax[:, :, :, 0]                                         # :572
    step = 0
    t0 = t.detach().clone()
    raypos0 = raypos_t.detach().clone()

    while (t < tminmax[:, :, :, 1]).any():                          # :579
        if max_steps is not None and step >= max_steps:
            break
        for k in range(K):                                          # :582
            y0 = (
                torch.bmm(
                    (raypos_t - primpos[:, k, None, None, :]).view(N, -1, 3),
                    primrot[:, k, :, :],
                ).view_as(raypos_t)
                * primscale[:, k, None, None, :]
            )                                                       # :583-589
            fade = torch.exp(-fadescale * torch.sum(torch.abs(y0) ** fadeexp, dim=-1, keepdim=True))  # :591
            if wfield is not None:                                  # :593-597
                y1 = F.grid_sample(wfield[:, k], y0[:, None, :, :, :], align_corners=True)[:, :, 0, :, :].permute(0, 2, 3, 1)
            else:
                y1 = y0
            sample = F.grid_sample(tplate[:, k], y1[:, None, :, :, :], align_corners=True)[
                :, :, 0, :, :
            ].permute(0, 2, 3, 1)                                   # :600-602
            valid1 = torch.prod(y0 >= -1.0, dim=-1, keepdim=True) * torch.prod(y0 <= 1.0, dim=-1, keepdim=True)
            valid = ((t >= tminmax[:, :, :, 0]) & (t < tminmax[:, :, :, 1])).to(raypos.dtype)[:, :, :, None]
            alpha0 = sample[:, :, :, 3:4]
            rgb = sample[:, :, :, 0:3] * valid * valid1
            alpha = alpha0 * fade * stepsize * valid * valid1
            newalpha = rayrgba[:, :, :, 3:4] + alpha                # :616
            contrib = (newalpha.clamp(max=1.0) - rayrgba[:, :, :, 3:4]) * valid * valid1
            rayrgba = rayrgba + contrib * torch.cat([rgb, torch.ones_like(alpha)], dim=-1)
        step += 1
        t = t0 + stepsize * step                                    # :623
        raypos_t = raypos0 + raydir * stepsize * step               # :624
    return rayrgba


def raymarch_torch_fwd_bwd(raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template,
                           grad_rayrgba, fadescale=8.0, fadeexp=8.0, warp=None):
    """Forward + autograd backward; returns (rayrgba, grad_primpos, grad_primrot, grad_primscale, grad_template
    [, grad_warp])."""
    leaves = [x.detach().clone().requires_grad_(True) for x in (primpos, primrot, primscale, template)]
    wl = warp.detach().clone().requires_grad_(True) if warp is not None else None
    out = raymarch_torch(raypos, raydir, stepsize, tminmax, *leaves, fadescale=fadescale, fadeexp=fadeexp, warp=wl)
    out.backward(grad_rayrgba)
    return (out.detach(),) + tuple(x.grad for x in leaves) + ((wl.grad,) if wl is not None else ())
