"""`Raymarcher`: host-side mirror of /root/reference/models/raymarchers/mvpraymarcher.py:17-54 (same constructor arguments,
same forward signature and return tuple).  The image planes `rayrgb` / `rayalpha` come straight out of the render kernel's
epilogue and their gradients go straight into the backward's prologue (`op.mvpraymarch_planes`); the reference makes them
with a permute + two `.contiguous()` copies (:50-51)."""
import torch.nn as nn

from .composite import split_rgba
from .op import mvpraymarch, mvpraymarch_camera, mvpraymarch_planes


class Raymarcher(nn.Module):
    def __init__(self, volradius, dt: float = 1.0, with_rgba: bool = False):
        """with_rgba: also return the [N,4,H,W] view of the channels-last `rayrgba` as the third item, like the reference
        (:50).  No caller in ava-256 reads it (models/autoencoder.py:246 drops it), so by default the channels-last image
        is never written and the third item is None."""
        super().__init__()
        self.volume_radius = volradius
        self.dt = dt / self.volume_radius                       # mvpraymarcher.py:24
        self.with_rgba = with_rgba

    def forward(self, raypos, raydir, tminmax, decout, renderoptions={}, rayterm=None, with_pos_img=None):
        args = (raypos, raydir, self.dt, tminmax, (decout["primpos"], decout["primrot"], decout["primscale"]))
        kwargs = dict(template=decout["template"], warp=decout["warp"] if "warp" in decout else None, rayterm=rayterm,
                      **{k: v for k, v in renderoptions.items() if k in mvpraymarch.__code__.co_varnames})   # :45
        if self.with_rgba:
            rayrgba = mvpraymarch(*args, **kwargs)
            rayrgb, rayalpha = split_rgba(rayrgba)              # one kernel instead of the permute + two copies
            return rayrgb, rayalpha, rayrgba.permute(0, 3, 1, 2), None
        rayrgb, rayalpha = mvpraymarch_planes(*args, **kwargs)
        return rayrgb, rayalpha, None, None

    def forward_camera(self, viewpos, viewrot, focal, princpt, pixelcoords, decout, renderoptions={}, rayterm=None):
        """`compute_raydirs(viewpos, viewrot, focal, princpt, pixelcoords, volradius)` followed by `forward(raypos, raydir, tminmax,
        ...)` -- models/autoencoder.py:240-252 -- as one call.  With `pixelcoords = (W, H)` the rays are generated inside the render
        kernels and never stored (`op.mvpraymarch_camera`); same return tuple as `forward`."""
        kwargs = dict(rayterm=rayterm, **{k: v for k, v in renderoptions.items() if k in mvpraymarch.__code__.co_varnames})
        out = mvpraymarch_camera(viewpos, viewrot, focal, princpt, pixelcoords, self.volume_radius, self.dt,
                                 (decout["primpos"], decout["primrot"], decout["primscale"]), decout["template"],
                                 decout["warp"] if "warp" in decout else None, planes=not self.with_rgba, **kwargs)
        if self.with_rgba:
            rayrgb, rayalpha = split_rgba(out)
            return rayrgb, rayalpha, out.permute(0, 3, 1, 2), None
        return out[0], out[1], None, None
