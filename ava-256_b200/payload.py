"""Payload hand-off from the decoders to the raymarcher on top of the C-ABI
(`mvp_assemble_payload_*`).

The reference turns the colour decoder's image [N, B*3, h*B, w*B] and the geometry decoder's opacity image
[N, B, h*B, w*B] into the raymarcher's channels-last slabs with a chain of eager ops,

    rgb.view(N, B, 3, h, B, w, B).permute(0, 3, 5, 1, 4, 6, 2).reshape(N, h*w, B, B, B, 3)   models/decoders/rgb.py:128-143
    opacity.view(N, B, 1, h, B, w, B).permute(0, 3, 5, 1, 4, 6, 2).reshape(N, h*w, B, B, B, 1)  models/decoders/geometry.py:180-185
    template = cat([relu(primrgb * 25.0 + 100.0), relu(primalpha)], dim=-1)                     models/decoders/assembler.py:261

i.e. five passes over 134 MB per view at K = 16384, B = 8.  `assemble_payload` does it in one pass (16 B read + 16 B
written per texel) and its adjoint in one more; values are bit-identical to the eager chain."""
import ctypes

import torch
from torch.autograd import Function

from . import lib as _lib
from .composite import _check_f32_cuda, _ptr


class AssemblePayload(Function):
    @staticmethod
    def forward(ctx, tex, opacity, boxsize, rgb_scale, rgb_bias):
        _check_f32_cuda("tex", tex)
        _check_f32_cuda("opacity", opacity)
        B = int(boxsize)
        assert tex.dim() == 4 and opacity.dim() == 4
        N, C, Himg, Wimg = tex.shape
        assert C == 3 * B and Himg % B == 0 and Wimg % B == 0, "tex must be [N, B*3, h*B, w*B]"
        assert tuple(opacity.shape) == (N, B, Himg, Wimg), "opacity must be [N, B, h*B, w*B]"
        hb, wb = Himg // B, Wimg // B
        tex, opacity = tex.contiguous(), opacity.contiguous()
        dev = tex.device
        with torch.cuda.device(dev):
            tplate = torch.empty((N, hb * wb, B, B, B, 4), device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.LIB.mvp_assemble_payload_forward(N, hb, wb, B, _ptr(tex), _ptr(opacity), float(rgb_scale),
                                                             float(rgb_bias), _ptr(tplate), stream))
        ctx.save_for_backward(tplate)
        ctx.dims = (N, hb, wb, B, float(rgb_scale))
        return tplate

    @staticmethod
    def backward(ctx, grad_tplate):
        (tplate,) = ctx.saved_tensors
        N, hb, wb, B, rgb_scale = ctx.dims
        grad_tplate = grad_tplate.contiguous()
        dev = tplate.device
        with torch.cuda.device(dev):
            grad_tex = torch.empty((N, 3 * B, hb * B, wb * B), device=dev)
            grad_opacity = torch.empty((N, B, hb * B, wb * B), device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.LIB.mvp_assemble_payload_backward(N, hb, wb, B, _ptr(tplate), _ptr(grad_tplate), rgb_scale,
                                                              _ptr(grad_tex), _ptr(grad_opacity), stream))
        return grad_tex, grad_opacity, None, None, None


def assemble_payload(tex, opacity, boxsize=8, rgb_scale=25.0, rgb_bias=100.0):
    """template [N, h*w, B, B, B, 4] from the decoders' images; defaults are the reference's hard-coded
    de-normalisation (assembler.py:261)."""
    return AssemblePayload.apply(tex, opacity, boxsize, rgb_scale, rgb_bias)


class ExpandViews(Function):
    @staticmethod
    def forward(ctx, x, n_views):
        _check_f32_cuda("x", x)
        x = x.contiguous()
        dev = x.device
        with torch.cuda.device(dev):
            out = torch.empty((int(n_views),) + tuple(x.shape), device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.LIB.mvp_expand_views(_ptr(x), _ptr(out), x.numel(), int(n_views), stream))
        return out

    @staticmethod
    def backward(ctx, grad):
        return sum_views(grad), None


def sum_views(x, out=None):
    """x [n_views, ...] -> sum over the views [...] in one pass (`mvp_sum_views`; view order, deterministic).  `out`: optional
    contiguous destination with x[0].numel() elements (e.g. a slice of a flat all-reduce buffer)."""
    _check_f32_cuda("x", x)
    x = x.contiguous()
    n, count = x.shape[0], x[0].numel()
    dev = x.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(tuple(x.shape[1:]), device=dev)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and out.numel() == count
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.LIB.mvp_sum_views(_ptr(x), _ptr(out), count, int(n), stream))
    return out


def expand_views(x, n_views):
    """[n_views, *x.shape] copies of one subject's tensor (`x[None].expand(n_views, ...).contiguous()` in one pass of
    streaming stores, `mvp_expand_views`); the adjoint is the sum over the views."""
    return ExpandViews.apply(x, n_views)
