"""Image epilogue of the render step on top of the C-ABI (`mvp_composite_*`).

One kernel replaces the eager chain the reference runs on the raymarcher's output:

    rayrgba.permute(0, 3, 1, 2); [:, :3].contiguous(); [:, 3:4].contiguous()   models/raymarchers/mvpraymarcher.py:50-51
    w[:, :, None, None] * rayrgb + b[:, :, None, None]                          models/colorcals/colorcal.py:26-29
    rayrgb + (1.0 - rayalpha) * bg                                              models/autoencoder.py:262-270

and its adjoint hands `grad_rayrgba` back contiguous channels-last, which is what the raymarch backward consumes
(the reference copies the permuted gradient first, mvpraymarch.py:212).  Forward values are bit-identical to the eager
expressions (each op rounded once)."""
import ctypes

import torch
from torch.autograd import Function

from . import lib as _lib


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_f32_cuda(name, t, shape=None):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError("%s must have shape %r, got %r" % (name, tuple(shape), tuple(t.shape)))


class Composite(Function):
    """(rayrgba [N,H,W,4], ccw [N,3] | None, ccb [N,3] | None, bg [N,3,H,W] | None) -> (irgbrec [N,3,H,W], rayalpha [N,1,H,W])"""

    @staticmethod
    def forward(ctx, rayrgba, ccw, ccb, bg):
        _check_f32_cuda("rayrgba", rayrgba)
        assert rayrgba.dim() == 4 and rayrgba.size(3) == 4
        N, H, W, _ = rayrgba.shape
        rayrgba = rayrgba.contiguous()
        if (ccw is None) != (ccb is None):                      # colorcal always has both (colorcal.py:26-28)
            ref = ccw if ccw is not None else ccb
            ccw = torch.ones_like(ref) if ccw is None else ccw
            ccb = torch.zeros_like(ref) if ccb is None else ccb
        if ccw is not None:
            _check_f32_cuda("ccw", ccw, (N, 3))
            _check_f32_cuda("ccb", ccb, (N, 3))
            ccw, ccb = ccw.contiguous(), ccb.contiguous()
        if bg is not None:
            _check_f32_cuda("bg", bg)
            bg = bg.expand(N, 3, H, W).contiguous()             # the black-background branch broadcasts [1,3,1,1] (autoencoder.py:268-270)
        dev = rayrgba.device
        with torch.cuda.device(dev):
            irgbrec = torch.empty((N, 3, H, W), device=dev)
            rayalpha = torch.empty((N, 1, H, W), device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.LIB.mvp_composite_forward(N, H, W, _ptr(rayrgba), _ptr(ccw), _ptr(ccb), _ptr(bg), _ptr(irgbrec),
                                                      _ptr(rayalpha), stream))
        ctx.save_for_backward(rayrgba, ccw, bg)
        return irgbrec, rayalpha

    @staticmethod
    def backward(ctx, grad_irgbrec, grad_rayalpha):
        rayrgba, ccw, bg = ctx.saved_tensors
        N, H, W, _ = rayrgba.shape
        need_rgba, need_ccw, need_ccb, need_bg = ctx.needs_input_grad
        dev = rayrgba.device
        grad_irgbrec = grad_irgbrec.contiguous()
        grad_rayalpha = None if grad_rayalpha is None else grad_rayalpha.contiguous()
        want_cc = ccw is not None and (need_ccw or need_ccb)
        want_bg = bg is not None and need_bg
        with torch.cuda.device(dev):
            grad_rayrgba = torch.empty_like(rayrgba)
            grad_ccw = torch.zeros((N, 3), device=dev) if want_cc else None
            grad_ccb = torch.zeros((N, 3), device=dev) if want_cc else None
            grad_bg = torch.empty((N, 3, H, W), device=dev) if want_bg else None
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.LIB.mvp_composite_backward(N, H, W, _ptr(rayrgba), _ptr(ccw), _ptr(bg), _ptr(grad_irgbrec),
                                                       _ptr(grad_rayalpha), _ptr(grad_rayrgba), _ptr(grad_ccw), _ptr(grad_ccb),
                                                       _ptr(grad_bg), stream))
        return (grad_rayrgba if need_rgba else None, grad_ccw if need_ccw else None, grad_ccb if need_ccb else None,
                grad_bg if need_bg else None)


def composite(rayrgba, ccw=None, ccb=None, bg=None):
    """irgbrec, rayalpha = composite(rayrgba, w, b, bg).

    `ccw`, `ccb` are the per-sample colour-calibration weight and bias the reference's `Colorcal.forward` forms as
    `wcam[camindex] + wident[idindex]` and `bcam[camindex] + bident[idindex]` (colorcal.py:26-28); `bg` is the
    background image [N,3,H,W] (anything broadcastable to it), or None for black."""
    if bg is not None and tuple(bg.shape) != (rayrgba.size(0), 3, rayrgba.size(1), rayrgba.size(2)):
        bg = bg.expand(rayrgba.size(0), 3, rayrgba.size(1), rayrgba.size(2))   # differentiable broadcast
    return Composite.apply(rayrgba, ccw, ccb, bg)


def split_rgba(rayrgba):
    """rayrgb [N,3,H,W], rayalpha [N,1,H,W] -- the two `.contiguous()` copies of mvpraymarcher.py:50-51 in one pass."""
    return Composite.apply(rayrgba, None, None, None)
