"""Multi-GPU plumbing of the path: views shard over ranks with no data-path collective in
forward; when the views of a step share one subject's primitives, the per-view primitive gradients are summed locally
and all-reduced once per step (one flat fp32 buffer: template, primpos, primrot, primscale)."""
import torch
import torch.distributed as dist


def shard_views(n_views, rank, world):
    """Contiguous block of views of `rank`: [start, stop)."""
    if n_views % world:
        raise ValueError("n_views (%d) must be divisible by world size (%d)" % (n_views, world))
    per = n_views // world
    return rank * per, (rank + 1) * per


def flat_grad_numel(K, TD, TH, TW):
    return K * TD * TH * TW * 4 + K * 15


def reduce_primitive_grads(grad_template, grad_primpos, grad_primrot, grad_primscale, flat=None, group=None):
    """Sum per-view gradients [nv, ...] over the local views into `flat` and all-reduce it (if a process group is up).

    Returns (flat, views) where views = (g_template [K,TD,TH,TW,4], g_primpos [K,3], g_primrot [K,3,3], g_primscale [K,3])
    are views into flat."""
    parts = (grad_template, grad_primpos, grad_primrot, grad_primscale)
    nv = grad_template.shape[0]
    total = sum(p[0].numel() for p in parts)
    if flat is None:
        flat = torch.empty(total, dtype=grad_template.dtype, device=grad_template.device)
    assert flat.numel() == total
    o = 0
    views = []
    for p in parts:
        n = p[0].numel()
        torch.sum(p.reshape(nv, n), dim=0, out=flat[o:o + n])
        views.append(flat[o:o + n].view(p.shape[1:]))
        o += n
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    return flat, tuple(views)
