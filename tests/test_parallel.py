"""World-size-2 gloo test (CPU) of the multi-GPU host logic: view sharding + local view-sum + one all-reduce of the
flat primitive-gradient buffer equals the single-process sum over all views."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import scene_args_np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _per_view_grads(n_views, view_lo, view_hi):
    """Oracle gradients of views [lo,hi) of a 4-view shared-subject scene (CPU checker; stands in for the GPU op)."""
    from ava256_b200 import scene
    from oracle import oracle
    s = scene.make_scene(n_views, 24, 16, 16, 4, alpha_mu=2.0, alpha_sigma=2.0)
    s = {k: (v[view_lo:view_hi].contiguous() if torch.is_tensor(v) else v) for k, v in s.items()}
    s["stepsize"] = 1.0 / 32
    a, kw = scene_args_np(s)
    rgba, raysat = oracle.forward(*a, **kw)
    g = np.random.default_rng(5).standard_normal((n_views,) + rgba.shape[1:]).astype(np.float32)[view_lo:view_hi]
    gp, gr, gs, gt = oracle.backward(*a, g, raysat, **kw)
    f32 = lambda x: torch.from_numpy(x.astype(np.float32))  # noqa: E731
    return f32(gt), f32(gp), f32(gr), f32(gs)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ava256_b200 import parallel
    lo, hi = parallel.shard_views(4, rank, world)
    g = _per_view_grads(4, lo, hi)
    flat, views = parallel.reduce_primitive_grads(*g)
    # the asynchronous, double-buffered reducer the benchmark step uses: three steps in flight over two buffers, scaled
    # gradients so that a buffer reused too early (before its all-reduce finished) would show
    red = parallel.GradReducer(16, 4, 4, 4, "cpu")
    flats = []
    for step in range(3):
        f_ = red.reduce(*[x * float(step + 1) for x in g])
        if step == 1:
            red.wait(f_)
            flats.append(f_.clone())
    red.finish()
    flats.append(red.bufs[0].clone())          # step 2 landed in buffer 0
    if rank == 0:
        torch.save((flat, flats), out)
    dist.destroy_process_group()


def test_view_sharding_and_gradient_allreduce(tmp_path):
    from ava256_b200 import parallel
    assert parallel.shard_views(80, 3, 8) == (30, 40)
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    flat2, flats = torch.load(out)
    flat1, views = parallel.reduce_primitive_grads(*_per_view_grads(4, 0, 4))
    assert parallel.rank_views(80, 3, 8) == list(range(3, 80, 8)) and parallel.rank_views(80, 3, 8, interleave=False) == list(range(30, 40))
    assert sorted(v for r in range(8) for v in parallel.rank_views(80, r, 8)) == list(range(80))
    tol = 1e-6 * float(flat1.abs().max())
    assert torch.allclose(flats[0], 2.0 * flat1, rtol=1e-5, atol=2 * tol) and torch.allclose(flats[1], 3.0 * flat1, rtol=1e-5, atol=3 * tol)
    assert flat1.numel() == parallel.flat_grad_numel(16, 4, 4, 4)
    assert views[0].shape == (16, 4, 4, 4, 4) and views[2].shape == (16, 3, 3)
    assert torch.allclose(flat1, flat2, rtol=1e-5, atol=1e-6 * float(flat1.abs().max()))
    assert float(flat1.abs().max()) > 0
