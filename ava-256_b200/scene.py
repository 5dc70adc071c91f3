"""Synthetic inputs for the mvpraymarch path: cameras, rays, primitives, payload.

Host-side input generation only -- it produces exactly the tensors `models/autoencoder.py:240-252` hands to
the raymarcher: rays from the `compute_raydirs` formula (/root/reference/extensions/utils/utils_kernel.cu:32-50),
slab placement as `models/decoders/assembler.py:180-253` lays slabs on the tracked mesh (recipe B of the survey:
a sqrt(K) x sqrt(K) UV grid mapped equal-area onto an ellipsoid with the asset head's extent), and a
non-negative RGBA payload like the decoder's relu output (`assembler.py:261`).

Everything is generated on the CPU from a seeded torch.Generator (identical on every machine), then moved.
"""
import math

import torch

VOLRADIUS = 256.0
CAM_DIST_MM = 1428.0          # |campos| of assets/camera_calibration.json (survey probe)
FOCAL_FULLRES = 10355.0       # px at 4096 x 2668
FULLRES_H = 4096
ELLIPSOID_MM = (95.0, 178.0, 112.0)   # asset mesh half-extent (x, y, z)


def fibonacci_dirs(n, offset=0, ids=None):
    """n well-spread unit vectors (Fibonacci lattice), restricted to the frontal 60% of the sphere like the rig.
    `ids` (optional) selects arbitrary lattice points of the 80-camera dome instead of the block [offset, offset + n)."""
    if ids is not None:
        i = torch.as_tensor(list(ids), dtype=torch.float64) + 0.5
        total = max(80, int(max(ids)) + 1)
    else:
        i = torch.arange(offset, offset + n, dtype=torch.float64) + 0.5
        total = max(80, n + offset)
    z = 1.0 - 1.2 * i / total          # z in (1, -0.2]: front and sides, like a capture dome
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    phi = i * math.pi * (3.0 - math.sqrt(5.0))
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=-1)


def look_at_cameras(n, offset=0, ids=None):
    """Returns (campos [n,3] in mm, camrot [n,3,3] rows = camera x,y,z axes in world coords)."""
    d = fibonacci_dirs(n, offset, ids)
    campos = d * CAM_DIST_MM
    zc = -d                                                    # camera looks at the origin
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64).expand_as(zc)
    xc = torch.linalg.cross(up, zc)
    bad = xc.norm(dim=-1) < 1e-6
    if bad.any():
        xc[bad] = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64)
    xc = xc / xc.norm(dim=-1, keepdim=True)
    yc = torch.linalg.cross(zc, xc)
    return campos, torch.stack([xc, yc, zc], dim=1)


def compute_raydirs_host(campos, camrot, focal, princpt, H, W, volradius=VOLRADIUS, dtype=torch.float32):
    """fp32 restatement of compute_raydirs_forward_kernel (utils_kernel.cu:32-50) on an integer pixel grid."""
    n = campos.shape[0]
    campos = campos.to(dtype)
    camrot = camrot.to(dtype)
    py, px = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    pc = torch.stack([px, py], dim=-1)[None].expand(n, H, W, 2)
    pc = (pc - princpt.to(dtype)[:, None, None, :]) / focal.to(dtype)[:, None, None, :]
    rd = torch.cat([pc, torch.ones_like(pc[..., :1])], dim=-1)
    rd = (camrot[:, None, None, 0, :] * rd[..., 0:1] + camrot[:, None, None, 1, :] * rd[..., 1:2]
          + camrot[:, None, None, 2, :] * rd[..., 2:3])
    rd = rd / rd.norm(dim=-1, keepdim=True)
    rp = (campos / volradius)[:, None, None, :].expand(n, H, W, 3).contiguous()
    t1 = (-1.0 - rp) / rd
    t2 = (1.0 - rp) / rd
    tmin = torch.minimum(t1, t2).amax(dim=-1)
    tmax = torch.maximum(t1, t2).amin(dim=-1)
    tminmax = torch.stack([tmin.clamp(min=0.0), tmax], dim=-1)
    return rp.contiguous(), rd.contiguous(), tminmax.contiguous()


def make_cameras(n_views, H, W, view_offset=0, view_ids=None):
    """Camera parameters of the dome views in the form `compute_raydirs` takes them (fp32): viewpos [n,3] (mm), viewrot [n,3,3],
    focal [n,2], princpt [n,2] -- what `make_rays` turns into rays on the host."""
    campos, camrot = look_at_cameras(n_views, view_offset, view_ids)
    ds = FULLRES_H / H
    focal = torch.full((n_views, 2), FOCAL_FULLRES / ds, dtype=torch.float64)
    princpt = torch.tensor([[W / 2.0, H / 2.0]], dtype=torch.float64).expand(n_views, 2)
    return tuple(t.to(torch.float32).contiguous() for t in (campos, camrot, focal, princpt))


def make_rays(n_views, H, W, view_offset=0, dtype=torch.float32, view_ids=None):
    """Rays of `n_views` dome cameras looking at the head: raypos, raydir [n,H,W,3], tminmax [n,H,W,2]."""
    campos, camrot = look_at_cameras(n_views, view_offset, view_ids)
    ds = FULLRES_H / H
    focal = torch.full((n_views, 2), FOCAL_FULLRES / ds, dtype=torch.float64)
    princpt = torch.tensor([[W / 2.0, H / 2.0]], dtype=torch.float64).expand(n_views, 2)
    return compute_raydirs_host(campos, camrot, focal, princpt, H, W, dtype=dtype)


def _rodrigues(rvec):
    theta = torch.sqrt(1e-12 + (rvec ** 2).sum(-1, keepdim=True))
    k = rvec / theta
    K = torch.zeros(rvec.shape[0], 3, 3, dtype=rvec.dtype)
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    eye = torch.eye(3, dtype=rvec.dtype)[None]
    s, c = torch.sin(theta)[..., None], torch.cos(theta)[..., None]
    return eye + s * K + (1 - c) * (K @ K)


def make_primitives(K, seed=1112, dtype=torch.float32):
    """One subject's slabs: primpos [K,3], primrot [K,3,3], primscale [K,3] (inverse half-extents)."""
    g = int(round(math.sqrt(K)))
    assert g * g == K, "K must be a square number (UV grid)"
    gen = torch.Generator().manual_seed(seed)
    a, b, c = (e / VOLRADIUS for e in ELLIPSOID_MM)
    v, u = torch.meshgrid(torch.arange(g, dtype=torch.float64), torch.arange(g, dtype=torch.float64), indexing="ij")

    def surf(uu, vv):
        ct = 1.0 - 2.0 * (vv + 0.5) / g
        st = torch.sqrt(torch.clamp(1.0 - ct * ct, min=0.0))
        ph = 2.0 * math.pi * (uu + 0.5) / g
        return torch.stack([a * st * torch.cos(ph), b * ct, c * st * torch.sin(ph)], dim=-1)

    p = surf(u, v)                                      # [g,g,3], k = v*g + u  (assembler.py:180 row-major UV)
    du = surf(u + 1.0, v) - p
    dv = surf(u, v + 1.0) - p
    t = du / du.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    nrm = torch.linalg.cross(du, dv)
    nrm = nrm / nrm.norm(dim=-1, keepdim=True).clamp(min=1e-12)
    bt = torch.linalg.cross(nrm, t)
    tbn = torch.stack([t, bt, nrm], dim=-1).reshape(K, 3, 3)   # columns = tangent, bitangent, normal
    spacing = torch.maximum(du.norm(dim=-1), dv.norm(dim=-1)).reshape(K)
    pos = p.reshape(K, 3) + 0.002 * torch.randn(K, 3, generator=gen, dtype=torch.float64)
    rot = tbn @ _rodrigues(0.01 * torch.randn(K, 3, generator=gen, dtype=torch.float64))
    scale = (0.8 * 2.0 / spacing) * torch.exp(0.05 * torch.randn(K, generator=gen, dtype=torch.float64))
    scale = scale[:, None].expand(K, 3).contiguous()
    return pos.to(dtype).contiguous(), rot.to(dtype).contiguous(), scale.to(dtype).contiguous()


def make_payload(K, T, seed=1112, alpha_mu=6.0, alpha_sigma=6.0, dtype=torch.float32, device="cpu"):
    """template [K,T,T,T,4] channels-last: rgb ~ U(0,255), alpha = relu(N(mu, sigma)) (1/unit length)."""
    gen = torch.Generator(device=device).manual_seed(seed + 7)
    tpl = torch.empty(K, T, T, T, 4, dtype=dtype, device=device)
    tpl[..., :3] = torch.rand(K, T, T, T, 3, generator=gen, dtype=dtype, device=device) * 255.0
    tpl[..., 3] = torch.relu(alpha_mu + alpha_sigma * torch.randn(K, T, T, T, generator=gen, dtype=dtype, device=device))
    return tpl


def make_scene(n_views, H, W, K, T, seed=1112, view_offset=0, device="cpu", alpha_mu=6.0, alpha_sigma=6.0,
               share_primitives=True, view_ids=None):
    """Full op inputs for `n_views` views of one subject, laid out as the reference op takes them.

    Returns dict(raypos, raydir, tminmax, primpos [N,K,3], primrot [N,K,3,3], primscale [N,K,3],
    template [N,K,T,T,T,4], stepsize).  With share_primitives the per-view tensors are materialised copies of
    one subject's primitives (SURVEY 8d "template materialised per view").
    """
    rp, rd, tmm = make_rays(n_views, H, W, view_offset, view_ids=view_ids)
    pos, rot, scale = make_primitives(K, seed)
    dev = torch.device(device)
    tpl = make_payload(K, T, seed, alpha_mu, alpha_sigma, device="cpu" if dev.type == "cpu" else device)
    out = dict(raypos=rp.to(dev), raydir=rd.to(dev), tminmax=tmm.to(dev), stepsize=1.0 / VOLRADIUS)
    pos, rot, scale, tpl = pos.to(dev), rot.to(dev), scale.to(dev), tpl.to(dev)
    if not share_primitives:
        gen = torch.Generator().manual_seed(seed + 13)
        jit = (0.001 * torch.randn(n_views, K, 3, generator=gen)).to(dev)
        out["primpos"] = (pos[None] + jit).contiguous()
    else:
        out["primpos"] = pos[None].expand(n_views, K, 3).contiguous()
    out["primrot"] = rot[None].expand(n_views, K, 3, 3).contiguous()
    out["primscale"] = scale[None].expand(n_views, K, 3).contiguous()
    out["template"] = tpl[None].expand(n_views, K, T, T, T, 4).contiguous()
    return out
