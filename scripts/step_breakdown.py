"""Where a small-launch render step goes (one rank of an 8-GPU run = 10 views): CUDA-event sections of the public-op step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ava256_b200 import parallel, scene
from ava256_b200.op import mvpraymarch
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 10
h, w, k, t = 1024, 667, 16384, 8
dev = torch.device("cuda", 0)
s = scene.make_scene(nv, h, w, k, t, view_ids=list(range(0, 80, 80 // nv)), device=dev, alpha_mu=17.0, alpha_sigma=6.0)
grad_out = torch.randn(nv, h, w, 4, device=dev)
leaves = [s[n].requires_grad_(True) for n in ("primpos", "primrot", "primscale", "template")]
red = parallel.GradReducer(k, t, t, t, dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
def step(marks=None):
    for x in leaves: x.grad = None
    if marks is not None: marks.append(ev()); marks[-1].record()
    out = mvpraymarch(s["raypos"], s["raydir"], s["stepsize"], s["tminmax"], (leaves[0], leaves[1], leaves[2]), leaves[3], None)
    if marks is not None: marks.append(ev()); marks[-1].record()
    out.backward(grad_out)
    if marks is not None: marks.append(ev()); marks[-1].record()
    red.reduce(leaves[3].grad, leaves[0].grad, leaves[1].grad, leaves[2].grad)
    if marks is not None: marks.append(ev()); marks[-1].record()
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
t0 = time.perf_counter()
for _ in range(20): step()
cpu = (time.perf_counter() - t0) / 20
e1.record(); torch.cuda.synchronize()
print("views %d: %.3f ms per step on the device, %.3f ms of CPU time to enqueue a step" % (nv, e0.elapsed_time(e1) / 20, cpu * 1e3))
acc = [0.0, 0.0, 0.0]
for _ in range(10):
    m = []; step(m); torch.cuda.synchronize()
    for i in range(3): acc[i] += m[i].elapsed_time(m[i + 1]) / 10
print("sections (events between calls, include launch gaps): forward call %.3f  backward call %.3f  view-sum %.3f" % tuple(acc))
