"""compute_raydirs: host-side mirror of /root/reference/extensions/utils/utils.py:21-51 on top of the C-ABI
(`mvp_compute_raydirs`).  Same call signature and the same (non-)gradient behaviour: the reference's backward returns
None for every input (utils.py:45-46)."""
import ctypes

import torch
from torch.autograd import Function

from . import lib as _lib


class ComputeRaydirs(Function):
    @staticmethod
    def forward(ctx, viewpos, viewrot, focal, princpt, pixelcoords, volradius):
        for name, t in (("viewpos", viewpos), ("viewrot", viewrot), ("focal", focal), ("princpt", princpt)):
            if not t.is_cuda:
                raise RuntimeError("%s must be a CUDA tensor" % name)            # utils.cpp CHECK_CUDA
            assert t.is_contiguous() and t.dtype == torch.float32                 # utils.py:24-25
        N = viewpos.size(0)
        if isinstance(pixelcoords, tuple):                                        # utils.py:28-30
            W, H = pixelcoords
            pixelcoords = None
        else:
            assert pixelcoords.is_cuda and pixelcoords.is_contiguous() and pixelcoords.dtype == torch.float32
            H, W = pixelcoords.size(1), pixelcoords.size(2)
        dev = viewpos.device
        with torch.cuda.device(dev):
            raypos = torch.empty((N, H, W, 3), device=dev)
            raydirs = torch.empty((N, H, W, 3), device=dev)
            tminmax = torch.empty((N, H, W, 2), device=dev)
            P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.LIB.mvp_compute_raydirs(N, H, W, P(viewpos), P(viewrot), P(focal), P(princpt), P(pixelcoords),
                                                    float(volradius), P(raypos), P(raydirs), P(tminmax), stream))
        return raypos, raydirs, tminmax

    @staticmethod
    def backward(ctx, grad_raypos, grad_raydirs, grad_tminmax):
        return None, None, None, None, None, None


def compute_raydirs(viewpos, viewrot, focal, princpt, pixelcoords, volradius):
    raypos, raydirs, tminmax = ComputeRaydirs.apply(viewpos, viewrot, focal, princpt, pixelcoords, volradius)
    return raypos, raydirs, tminmax
