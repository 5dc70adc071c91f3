"""TEST INFRASTRUCTURE -- driver for the UNMODIFIED reference CUDA extension built by oracle/build_ref.py
(oracle/_ref/mvpraymarchlib.so).  It calls the reference's pybind functions (mvpraymarch.cpp:398-405) with the
fixed-order tree the reference's Python builds (mvpraymarch.py:58-75), written here from its closed form:
children(i) = (2i+1, 2i+2), parent(i) = floor((i-1)/2), leaf K-1+k <-> slab k.

Never imported by the product.  /root/reference is not needed at run time, only the prebuilt .so.
"""
import importlib.util
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "mvpraymarchlib.so")
_mod = None


def available():
    return os.path.exists(_SO) and torch.cuda.is_available()


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("mvpraymarchlib", _SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def _tree(N, K, dev):
    i = torch.arange(2 * K - 1, dtype=torch.int32, device=dev)
    children = torch.stack([2 * i + 1, 2 * i + 2], dim=-1)
    children[K - 1:] = -1
    parent = torch.div(i - 1, 2, rounding_mode="floor")
    sortedobjid = (torch.arange(N * K, dtype=torch.int32, device=dev) % K).view(N, K)
    return (sortedobjid.contiguous(), children[None].repeat(N, 1, 1).contiguous(), parent[None].repeat(N, 1).contiguous())


def forward(raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template, fadescale=8.0, fadeexp=8.0,
            blocksize=(8, 16), warp=None):
    """Returns (rayrgba, raysat, state) from the reference kernels (one call; N*K*T^3*4 must stay < 2^31)."""
    m = module()
    N, H, W = raypos.shape[:3]
    K = primpos.shape[1]
    assert N * K * template[0, 0].numel() < 2 ** 31, "reference int32 stride overflow (primsampler.h:31-36): chunk the views"
    dev = raypos.device
    sortedobjid, nodechildren, nodeparent = _tree(N, K, dev)
    nodeaabb = torch.empty((N, 2 * K - 1, 2, 3), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    m.compute_aabb(primpos, primrot, primscale, sortedobjid, nodechildren, nodeparent, nodeaabb, 0)
    rayrgba = torch.empty((N, H, W, 4), device=dev)
    raysat = torch.full((N, H, W, 3), -1.0, device=dev)
    m.raymarch_forward(raypos, raydir, stepsize, tminmax, sortedobjid, nodechildren, nodeaabb, primpos, primrot, primscale,
                       template, warp, rayrgba, raysat, None, 1 if warp is not None else 0, False, 512, True, True, fadescale,
                       fadeexp, 0, 0.0, 3, blocksize[0], blocksize[1])
    torch.cuda.synchronize()
    return rayrgba, raysat, (sortedobjid, nodechildren, nodeaabb)


def backward(raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template, rayrgba, raysat, state,
             grad_rayrgba, fadescale=8.0, fadeexp=8.0, blocksize=(8, 16), warp=None):
    m = module()
    sortedobjid, nodechildren, nodeaabb = state
    g = [torch.zeros_like(t) for t in (primpos, primrot, primscale, template)]
    gw = torch.zeros_like(warp) if warp is not None else None
    torch.cuda.synchronize()
    m.raymarch_backward(raypos, raydir, stepsize, tminmax, sortedobjid, nodechildren, nodeaabb, primpos, g[0], primrot,
                        g[1], primscale, g[2], template, g[3], warp, gw, rayrgba, grad_rayrgba.contiguous(), raysat, None,
                        1 if warp is not None else 0, False, 512, True, True, fadescale, fadeexp, 0, 0.0, 3, blocksize[0],
                        blocksize[1])
    torch.cuda.synchronize()
    return tuple(g) + ((gw,) if warp is not None else ())


# ---- reference ray generator (extensions/utils -> utilslib), for the compute_raydirs parity test ----
_USO = os.path.join(os.path.dirname(_SO), "utilslib", "utilslib.so")
_umod = None


def utils_available():
    return os.path.exists(_USO) and torch.cuda.is_available()


def compute_raydirs(viewpos, viewrot, focal, princpt, pixelcoords, H, W, volradius):
    """Calls the reference's compute_raydirs_forward (utils.cpp:46-82); pixelcoords may be None."""
    global _umod
    if _umod is None:
        spec = importlib.util.spec_from_file_location("utilslib", _USO)
        _umod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_umod)
    N = viewpos.shape[0]
    dev = viewpos.device
    raypos = torch.empty((N, H, W, 3), device=dev)
    raydir = torch.empty((N, H, W, 3), device=dev)
    tminmax = torch.empty((N, H, W, 2), device=dev)
    torch.cuda.synchronize()
    _umod.compute_raydirs_forward(viewpos, viewrot, focal, princpt, pixelcoords, W, H, volradius, raypos, raydir, tminmax)
    torch.cuda.synchronize()
    return raypos, raydir, tminmax
