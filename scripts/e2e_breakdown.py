"""What the end-to-end step (bench.py `e2e`) pays on top of the device-resident step: the same chunked, double-buffered
pipeline with parts switched off.  usage: e2e_breakdown.py [chunk]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ava256_b200 import parallel, scene
from ava256_b200.op import mvpraymarch
from ava256_b200.payload import expand_views

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nv, h, w, k, t = 80, 1024, 667, 16384, 8
dev = torch.device("cuda", 0)
s = scene.make_scene(nv, h, w, k, t, seed=1112, view_ids=list(range(nv)), device=dev, alpha_mu=17.0, alpha_sigma=6.0)
stepsize = s["stepsize"]
grad_out = torch.randn(nv, h, w, 4, device=dev)
names = ("primpos", "primrot", "primscale", "template")
host_in = {n: s[n].detach().cpu().pin_memory() for n in ("raypos", "raydir", "tminmax")}
host_prim = {n: s[n].detach()[0].cpu().pin_memory() for n in names}
host_grad = grad_out.cpu().pin_memory()
host_out = torch.empty(nv, h, w, 4).pin_memory()
red = parallel.GradReducer(k, t, t, t, dev)
flats = red.bufs
host_flat = torch.empty(flats[0].numel()).pin_memory()
dev_in = [{n: s[n] for n in host_in}, {n: torch.empty_like(s[n]) for n in host_in}]
dev_grad = [grad_out, torch.empty_like(grad_out)]
for n in names:
    del s[n]
torch.cuda.empty_cache()
dev_prim = [{n: host_prim[n].to(dev) for n in names} for _ in range(2)]
bounds = [(i, min(i + chunk, nv)) for i in range(0, nv, chunk)]
s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
cmp_done, flat_out_done = [None, None], [None, None]


def e2e_step(i, h2d=True, d2h=True, expand=True, persist=None):
    b_ = i & 1
    ev_in = []
    with torch.cuda.stream(s_in):
        if cmp_done[b_] is not None:
            s_in.wait_event(cmp_done[b_])
        if h2d:
            for n in names:
                dev_prim[b_][n].copy_(host_prim[n], non_blocking=True)
        for (a0, a1) in bounds:
            if h2d:
                for n in host_in:
                    dev_in[b_][n][a0:a1].copy_(host_in[n][a0:a1], non_blocking=True)
                dev_grad[b_][a0:a1].copy_(host_grad[a0:a1], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s_in)
            ev_in.append(e)
    flat_ = flats[b_]
    if flat_out_done[b_] is not None:
        s_cmp.wait_event(flat_out_done[b_])
    flat_.zero_()
    di, pr = dev_in[b_], dev_prim[b_]
    for ci, (a0, a1) in enumerate(bounds):
        s_cmp.wait_event(ev_in[ci])
        nvc = a1 - a0
        if expand or persist is None:
            if os.environ.get('TORCH_EXPAND'):
                lv = [pr[n][None].expand(nvc, *pr[n].shape).contiguous().requires_grad_(True) for n in names]
            else:
                lv = [expand_views(pr[n], nvc).requires_grad_(True) for n in names]
        else:
            lv = persist
            for x in lv:
                x.grad = None
        o_ = mvpraymarch(di["raypos"][a0:a1], di["raydir"][a0:a1], stepsize, di["tminmax"][a0:a1], (lv[0], lv[1], lv[2]), lv[3], None)
        o_.backward(dev_grad[b_][a0:a1])
        off = 0
        for x in (lv[3], lv[0], lv[1], lv[2]):
            n_ = x[0].numel()
            flat_[off:off + n_] += x.grad.view(nvc, n_).sum(dim=0)
            off += n_
        od = o_.detach()
        e = torch.cuda.Event()
        e.record(s_cmp)
        if d2h:
            with torch.cuda.stream(s_out):
                s_out.wait_event(e)
                host_out[a0:a1].copy_(od, non_blocking=True)
                od.record_stream(s_out)
    e = torch.cuda.Event()
    e.record(s_cmp)
    cmp_done[b_] = e
    with torch.cuda.stream(s_out):
        s_out.wait_event(e)
        if d2h:
            host_flat.copy_(flat_, non_blocking=True)
        e2 = torch.cuda.Event()
        e2.record(s_out)
        flat_out_done[b_] = e2


def drain():
    s_cmp.wait_stream(s_in)
    s_cmp.wait_stream(s_out)


def timeit(label, nrep=10, **kw):
    for i in range(2):
        e2e_step(i, **kw)
    drain()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(nrep):
        e2e_step(i, **kw)
    drain()
    b.record()
    torch.cuda.synchronize()
    print("%-44s %.2f ms per step" % (label, a.elapsed_time(b) / nrep), flush=True)


print("chunk %d views" % chunk)
timeit("full e2e step")
timeit("no uploads", h2d=False)
timeit("no downloads", d2h=False)
timeit("no copies at all", h2d=False, d2h=False)
pers = [dev_prim[0][n][None].expand(chunk, *dev_prim[0][n].shape).contiguous().requires_grad_(True) for n in names]
timeit("no copies, no per-view expand", h2d=False, d2h=False, expand=False, persist=pers)
# raw copy rates
for label, fn in (("H2D 2.76 GB", lambda: [dev_in[1][n].copy_(host_in[n], non_blocking=True) for n in host_in] + [dev_grad[1].copy_(host_grad, non_blocking=True)]),
                  ("D2H 0.87 GB", lambda: host_out.copy_(dev_grad[1], non_blocking=True))):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    print("%s alone: %.2f ms" % (label, a.elapsed_time(b)))
# glue rates on their own: per-view expand of the template (134 MB -> 20 x 134 MB) and the view-sum of its gradient
src = dev_prim[0]["template"]
def rate(label, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-34s %.3f ms  %.2f TB/s" % (label, ms, nbytes / ms / 1e9))
big = torch.empty((chunk,) + tuple(src.shape), device=dev)
import ctypes
from ava256_b200 import lib
rate("expand().contiguous() x%d" % chunk, lambda: big.copy_(src[None].expand_as(big)), big.numel() * 4)
rate("mvp_expand_views x%d" % chunk, lambda: expand_views(src, chunk), big.numel() * 4)
outv = torch.empty(src.numel(), device=dev)
rate("torch.sum over %d views" % chunk, lambda: torch.sum(big.view(chunk, -1), dim=0, out=outv), big.numel() * 4)
rate("zero-fill", lambda: big.zero_(), big.numel() * 4)
