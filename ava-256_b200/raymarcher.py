"""`Raymarcher`: host-side mirror of /root/reference/models/raymarchers/mvpraymarcher.py:17-54 (same constructor,
same forward signature and return tuple) with the output split done by one kernel (`composite.split_rgba`) instead of a
permute + two `.contiguous()` copies (:50-51)."""
import torch.nn as nn

from .composite import split_rgba
from .op import mvpraymarch


class Raymarcher(nn.Module):
    def __init__(self, volradius, dt: float = 1.0):
        super().__init__()
        self.volume_radius = volradius
        self.dt = dt / self.volume_radius                       # mvpraymarcher.py:24

    def forward(self, raypos, raydir, tminmax, decout, renderoptions={}, rayterm=None, with_pos_img=None):
        rayrgba = mvpraymarch(
            raypos, raydir, self.dt, tminmax,
            (decout["primpos"], decout["primrot"], decout["primscale"]),
            template=decout["template"],
            warp=decout["warp"] if "warp" in decout else None,
            rayterm=rayterm,
            **{k: v for k, v in renderoptions.items() if k in mvpraymarch.__code__.co_varnames},   # :45
        )
        assert rayrgba is not None
        rayrgb, rayalpha = split_rgba(rayrgba)
        return rayrgb, rayalpha, rayrgba.permute(0, 3, 1, 2), None   # third item is the permuted view, as in :50
