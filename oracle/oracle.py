"""TEST INFRASTRUCTURE -- ctypes front-end of oracle/mvp_oracle.c (numpy in, numpy out).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mvp_oracle.c")
LIB = os.path.join(HERE, "libmvp_oracle.so")


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off, compiled twice (float / double) into one shared object."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    objs = []
    for real, suf in (("float", "f32"), ("double", "f64")):
        obj = os.path.join(HERE, "mvp_oracle_%s.o" % suf)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-std=c11",
             "-DREAL=%s" % real, "-DSUFFIX=%s" % suf, "-c", SRC, "-o", obj]
        )
        objs.append(obj)
    subprocess.check_call(["gcc", "-shared", "-fopenmp", "-o", LIB] + objs + ["-lm"])
    return LIB


class _Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "H", "W", "K", "TD", "TH", "TW", "bsx", "bsy", "maxhitboxes")] + [
        (n, ctypes.c_double) for n in ("stepsize", "fadescale", "fadeexp")
    ] + [(n, ctypes.c_int32) for n in ("WD", "WH", "WW")]


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _cfg(raypos, primpos, template, stepsize, fadescale, fadeexp, blocksize, maxhitboxes, warp=None):
    N, H, W = raypos.shape[:3]
    K = primpos.shape[1]
    TD, TH, TW = template.shape[2:5]
    WD, WH, WW = warp.shape[2:5] if warp is not None else (0, 0, 0)
    return _Cfg(N, H, W, K, TD, TH, TW, blocksize[0], blocksize[1], maxhitboxes, stepsize, fadescale, fadeexp, WD, WH, WW)


def _prep(dtype, *arrs):
    return [np.ascontiguousarray(a, dtype=dtype) for a in arrs]


def forward(raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template, fadescale=8.0, fadeexp=8.0,
            blocksize=(8, 16), maxhitboxes=512, dtype=np.float32, want_raysat=True, return_stats=False, warp=None):
    """Reference-semantics forward on the CPU.  Returns (rayrgba, raysat[, stats]).
    warp: optional [N,K,WD,WH,WW,3] channels-last warp field (algo 1)."""
    lib = _load()
    suf = "f32" if dtype == np.float32 else "f64"
    raypos, raydir, tminmax, primpos, primrot, primscale, template = _prep(
        dtype, raypos, raydir, tminmax, primpos, primrot, primscale, template)
    if warp is not None:
        warp = np.ascontiguousarray(warp, dtype=dtype)
    cfg = _cfg(raypos, primpos, template, stepsize, fadescale, fadeexp, blocksize, maxhitboxes, warp)
    N, H, W = raypos.shape[:3]
    rayrgba = np.empty((N, H, W, 4), dtype)
    raysat = np.full((N, H, W, 3), -1, dtype) if want_raysat else None
    stats = np.zeros(4, np.int64)
    fn = getattr(lib, "mvp_oracle_forward_" + suf)
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(cfg), _p(raypos), _p(raydir), _p(tminmax), _p(primpos), _p(primrot), _p(primscale),
            _p(template), _p(warp), _p(rayrgba), _p(raysat), _p(stats))
    assert rc == 0
    if return_stats:
        return rayrgba, raysat, dict(samples=int(stats[0]), warp_steps=int(stats[1]), list_entries=int(stats[2]),
                                     capped_warps=int(stats[3]))
    return rayrgba, raysat


def backward(raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template, grad_rayrgba, raysat,
             fadescale=8.0, fadeexp=8.0, blocksize=(8, 16), maxhitboxes=512, dtype=np.float32, warp=None):
    """Reference-semantics backward.  Returns float64 (grad_primpos, grad_primrot, grad_primscale, grad_template)
    and, when `warp` is given, grad_warp as a fifth element."""
    lib = _load()
    suf = "f32" if dtype == np.float32 else "f64"
    raypos, raydir, tminmax, primpos, primrot, primscale, template, grad_rayrgba, raysat = _prep(
        dtype, raypos, raydir, tminmax, primpos, primrot, primscale, template, grad_rayrgba, raysat)
    if warp is not None:
        warp = np.ascontiguousarray(warp, dtype=dtype)
    cfg = _cfg(raypos, primpos, template, stepsize, fadescale, fadeexp, blocksize, maxhitboxes, warp)
    g = [np.zeros(a.shape, np.float64) for a in (primpos, primrot, primscale, template)]
    gw = np.zeros(warp.shape, np.float64) if warp is not None else None
    fn = getattr(lib, "mvp_oracle_backward_" + suf)
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(cfg), _p(raypos), _p(raydir), _p(tminmax), _p(primpos), _p(primrot), _p(primscale),
            _p(template), _p(warp), _p(grad_rayrgba), _p(raysat), _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(gw))
    assert rc == 0
    return tuple(g) + ((gw,) if warp is not None else ())


def aabb(primpos, primrot, primscale, dtype=np.float32):
    """Node AABBs [2K-1,2,3] of ONE view (bvh.cu:157-201 on the fixed-order heap)."""
    lib = _load()
    suf = "f32" if dtype == np.float32 else "f64"
    primpos, primrot, primscale = _prep(dtype, primpos, primrot, primscale)
    K = primpos.shape[0]
    out = np.empty((2 * K - 1, 2, 3), dtype)
    getattr(lib, "mvp_oracle_aabb_" + suf)(ctypes.c_int(K), _p(primpos), _p(primrot), _p(primscale), _p(out))
    return out
