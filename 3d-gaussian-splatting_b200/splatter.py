"""Scene + per-frame pipeline: host-side mirror of the reference's `Splatter`
(reference splatter.py:323-655) on top of the fused B200 frame path.

What train.py / visergui.py touch is kept (train.py:57-61,:67,:99,:150,:198-199,
visergui.py:137-149): `gaussian_3ds.{pos,rgb,opa,quat,scale}` (nn.Parameters, same
conventions: opa / rgb logits, wxyz quaternion, raw scale), `forward(camera_id,
extrinsics, intrinsics)` -> clamped, centre-cropped HxWx3 image, `ground_truth`, `imgs`,
`culling_mask` (int64), `n_tile_gaussians`, `n_gaussians`, `device`,
`scale_activation`, `set_camera`.

The per-frame work of reference `project_and_culling` + `render` (:513-634: 4 boolean-mask
compactions, the dense [T, N/20] list, cumsum, two 4-tensor gathers, fp32-key sort,
>= 7 host syncs) is ONE autograd node here (`renderer.render_frame`).

Scenes come from tensors (`Splatter.from_tensors`, synthetic benchmarks / tests) or from a
COLMAP model through `colmap_io` (reference splatter.py:363-412).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

import gaussian
from renderer import render_frame, render_frame_final

EPS = 1e-4


class Gaussian3ds(nn.Module):
    """Parameter holder (reference splatter.py:39-58, init_values=True branch)."""

    def __init__(self, pos, rgb, opa, quat, scale):
        super().__init__()
        self.pos = nn.Parameter(pos)
        self.rgb = nn.Parameter(rgb)
        self.opa = nn.Parameter(opa)
        self.quat = nn.Parameter(quat)
        self.scale = nn.Parameter(scale)

    def reset_opa(self):                                        # reference splatter.py:119-120
        with torch.no_grad():
            self.opa.fill_(math.log(0.01 / 0.99))


class Tiles:
    """Padded render-target geometry (reference splatter.py:255-272)."""

    def __init__(self, width, height, focal_x, focal_y):
        self.width, self.height = int(width), int(height)
        self.padded_width = int(math.ceil(self.width / 16)) * 16
        self.padded_height = int(math.ceil(self.height / 16)) * 16
        self.focal_x, self.focal_y = float(focal_x), float(focal_y)
        self.n_tile_x = self.padded_width // 16
        self.n_tile_y = self.padded_height // 16

    def __len__(self):
        return self.n_tile_x * self.n_tile_y

    def crop(self, image):
        top = (self.padded_height - self.height) // 2
        left = (self.padded_width - self.width) // 2
        return image[top:top + self.height, left:left + self.width, :]


class Splatter(nn.Module):
    def __init__(self, gaussians: dict, views: Sequence[dict], images: Optional[List[torch.Tensor]] = None,
                 near=0.3, use_sh_coeff=False, tile_culling_prob_thresh=0.05, scale_activation="abs",
                 device=None, debug=0):
        """gaussians: dict(pos, rgb, opa, quat, scale) raw parameter tensors.
        views: list of dict(width, height, focal_x, focal_y, rot[3,3], tran[3]) (world->camera).
        images: optional per-view uint8 HxWx3 ground truth (reference keeps them on the GPU)."""
        super().__init__()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.use_sh_coeff = use_sh_coeff          # rgb is [N,27] (degree 2) / [N,48] (degree 3) raw SH coefficients
        if bool(use_sh_coeff) != (gaussians["rgb"].shape[1] != 3):
            raise ValueError("use_sh_coeff must match the colour width (3 = RGB logits, 27 / 48 = SH)")
        self.near = near
        self.tile_culling_prob_thresh = tile_culling_prob_thresh
        self.scale_activation = scale_activation
        self.debug = debug
        to = dict(device=self.device, dtype=torch.float32)
        self.gaussian_3ds = Gaussian3ds(*(gaussians[k].detach().to(**to).contiguous()
                                         for k in ("pos", "rgb", "opa", "quat", "scale")))
        self.views = [dict(v) for v in views]
        for v in self.views:                     # camera is host data; device copies only for users
            v["rot"] = torch.as_tensor(v["rot"], dtype=torch.float32).cpu().contiguous()
            v["tran"] = torch.as_tensor(v["tran"], dtype=torch.float32).cpu().contiguous()
        self.imgs = [] if images is None else [im.to(self.device) for im in images]
        self._rctx = gaussian.RenderContext()
        self.ground_truth = None
        self.culling_mask = None
        self.n_tile_gaussians = 0
        self.n_gaussians = self.gaussian_3ds.pos.shape[0]
        self.current_view = None
        self.tile_info = None
        if self.views:
            self.set_camera(0)

    @classmethod
    def from_tensors(cls, gaussians, views, **kw):
        return cls(gaussians, views, **kw)

    # -- camera -----------------------------------------------------------------------------
    def set_camera(self, idx, extrinsics=None, intrinsics=None):
        """reference splatter.py:465-511 (idx=None: free camera from the GUI)."""
        if idx is None:
            v = dict(width=int(math.ceil(intrinsics["width"])), height=int(math.ceil(intrinsics["height"])),
                     focal_x=float(intrinsics["focal_x"]), focal_y=float(intrinsics["focal_y"]),
                     rot=torch.as_tensor(extrinsics["rot"], dtype=torch.float32).cpu().contiguous(),
                     tran=torch.as_tensor(extrinsics["tran"], dtype=torch.float32).cpu().contiguous())
            self.ground_truth = None
        else:
            v = self.views[idx]
            self.ground_truth = (self.imgs[idx].to(torch.float16) / 255.) if idx < len(self.imgs) else None
        self.current_view = v
        self.current_w2c_rot = v["rot"]
        self.current_w2c_tran = v["tran"]
        self.tile_info = Tiles(v["width"], v["height"], v["focal_x"], v["focal_y"])

    # -- frame ------------------------------------------------------------------------------
    def render_padded(self):
        """Padded, un-clamped image (what reference `render` returns, splatter.py:563-634)."""
        g, v = self.gaussian_3ds, self.current_view
        image, mask = render_frame(self._rctx, g.pos, g.rgb, g.opa, g.quat, g.scale, v["width"], v["height"],
                                   v["focal_x"], v["focal_y"], v["rot"], v["tran"], self.near,
                                   self.tile_culling_prob_thresh, self.scale_activation)
        self.culling_mask = mask
        self.n_gaussians = g.pos.shape[0]
        return image

    def forward(self, camera_id=None, extrinsics=None, intrinsics=None):
        """reference splatter.py:643-655; clamp(0,1) + centre crop (:652-653) run inside the blend
        kernels (`render_frame_final`)."""
        self.set_camera(camera_id, extrinsics, intrinsics)
        g, v = self.gaussian_3ds, self.current_view
        image, mask = render_frame_final(self._rctx, g.pos, g.rgb, g.opa, g.quat, g.scale, v["width"], v["height"],
                                         v["focal_x"], v["focal_y"], v["rot"], v["tran"], self.near,
                                         self.tile_culling_prob_thresh, self.scale_activation)
        self.culling_mask = mask
        self.n_gaussians = g.pos.shape[0]
        return image

    def forward_unfused_post(self, camera_id=None, extrinsics=None, intrinsics=None):
        """Same image through the padded raw render + torch clamp/crop (kept for cross-checks)."""
        self.set_camera(camera_id, extrinsics, intrinsics)
        padded = self.render_padded()
        return self.tile_info.crop(torch.clamp(padded, 0, 1))

    def frame_stats(self):
        s = self._rctx.stats()
        self.n_tile_gaussians = int(s["n_instances"])
        return s
