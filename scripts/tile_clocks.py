"""Diagnostics (needs a library built with -DMVP_TILE_CLOCKS=1): per-tile start / end times of one forward launch."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ava256_b200 import lib, scene
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
H, W, K, T = 1024, 667, 16384, 8
s = scene.make_scene(N, H, W, K, T, alpha_mu=17.0, alpha_sigma=6.0, device="cuda")
wsb = lib.workspace_bytes(N, H, W, K, T, T, T)
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
rgba = torch.empty(N, H, W, 4, device="cuda"); rsat = torch.empty(N, H, W, 3, device="cuda"); raux = torch.empty(N, H, W, 4, dtype=torch.int32, device="cuda")
P = lambda x: ctypes.c_void_p(x.data_ptr())
fa = lib.ForwardArgs(); fa.shape = lib.Shape(N, H, W, K, T, T, T)
fa.stepsize, fa.fadescale, fa.fadeexp, fa.flags = s["stepsize"], 8.0, 8.0, 0
fa.raypos, fa.raydir, fa.tminmax = P(s["raypos"]), P(s["raydir"]), P(s["tminmax"])
fa.primpos, fa.primrot, fa.primscale, fa.tplate = P(s["primpos"]), P(s["primrot"]), P(s["primscale"]), P(s["template"])
fa.rayrgba, fa.raysat, fa.rayaux, fa.workspace, fa.workspace_bytes = P(rgba), P(rsat), P(raux), P(ws), wsb
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    lib.check(lib.LIB.mvp_raymarch_forward(ctypes.byref(fa), st)); fa.flags = 1
torch.cuda.synchronize()
f = lib.LIB.mvp_debug_tileclk_offset; f.restype = ctypes.c_size_t; f.argtypes = [ctypes.POINTER(lib.Shape)]
off = f(ctypes.byref(fa.shape))
tiles = N * ((H + 3) // 4) * ((W + 7) // 8)
clk = ws[off:off + tiles * 32].cpu().numpy().view(np.int64).reshape(tiles, 4)
g0, g1, cyc, meta = clk[:, 0], clk[:, 1], clk[:, 2], clk[:, 3]
t0 = g0.min()
start = (g0 - t0) / 1e3; end = (g1 - t0) / 1e3; dur = end - start
print("N=%d tiles=%d kernel span %.1f us" % (N, tiles, end.max()))
print("tile duration us: mean %.1f p50 %.1f p90 %.1f p99 %.1f p99.9 %.1f max %.1f" % (dur.mean(), *np.percentile(dur, [50, 90, 99, 99.9]), dur.max()))
order = np.argsort(-dur)[:8]
for i in order:
    print("  tile %7d view %d: start %.1f end %.1f dur %.1f us  cta %d sm %d" % (i, i // (tiles // N), start[i], end[i], dur[i], meta[i] >> 32, meta[i] & 0xffff))
# how many SMs are busy over time
edges = np.linspace(0, end.max(), 41)
busy = [(np.minimum(end, b) - np.maximum(start, a)).clip(min=0).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
print("avg concurrently running tiles (of %d slots) per 1/40 of the span:" % (148 * 28))
print(" ".join("%d" % x for x in busy))
late = start > 0.9 * end.max()
print("tiles started in the last 10%% of the span: %d, their max duration %.1f us" % (late.sum(), dur[late].max() if late.any() else 0))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "tile_dur_us.npy"), dur.astype(np.float32).reshape(N, (H + 3) // 4, (W + 7) // 8))
np.save(os.path.join(ROOT, "gpurun_out", "tile_start_us.npy"), start.astype(np.float32).reshape(N, (H + 3) // 4, (W + 7) // 8))
