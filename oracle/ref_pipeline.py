"""Replays the reference's per-frame call sequence against any `gaussian`-compatible
extension module.  TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

reference splatter.py is not importable here (pykdtree / kornia / viser / torchmetrics are
missing) and needs a COLMAP dataset, so this driver issues — on seeded
synthetic tensors — exactly the operator sequence of `Splatter.project_and_culling` +
`Splatter.render` + `forward` (splatter.py:513-655): torch pre-activations, `global_culling`,
4 boolean-mask compactions, dense [T, Nc//20] list fill, `calc_tile_list` (with the 4 clones of
`_tocpp`), clamp, sum/cumsum/max syncs, `gather_gaussians`, 4-tensor gather, fp32 composite key
`torch.sort`, second 4-tensor gather, `draw`, clamp + centre crop.  Used with
  * the reference's own CUDA build (oracle/_ref/gaussian_ref*.so + its renderer.py) as the
    `--impl reference` arm of bench.py and as the live parity partner in tests/, and
  * our `gaussian` module, to prove the drop-in boundary.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import sysconfig

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def load_reference():
    """(gaussian_ref module, reference renderer module bound to it) or (None, None)."""
    so = os.path.join(REF_DIR, "gaussian_ref" + sysconfig.get_config_var("EXT_SUFFIX"))
    rpy = os.path.join(REF_DIR, "renderer.py")
    if not (os.path.exists(so) and os.path.exists(rpy)):
        return None, None
    if "gaussian_ref" in sys.modules and "renderer_ref" in sys.modules:
        return sys.modules["gaussian_ref"], sys.modules["renderer_ref"]
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("gaussian_ref", so)
    gref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gref)
    sys.modules["gaussian_ref"] = gref
    # the reference's renderer.py does `import gaussian`: bind that name to the reference build
    # only while it is being imported
    saved = sys.modules.get("gaussian")
    sys.modules["gaussian"] = gref
    try:
        spec = importlib.util.spec_from_file_location("renderer_ref", rpy)
        rref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rref)
        sys.modules["renderer_ref"] = rref
    finally:
        if saved is None:
            del sys.modules["gaussian"]
        else:
            sys.modules["gaussian"] = saved
    return gref, rref


class LegacyFrame:
    """One view rendered through the legacy stage-by-stage boundary of module `gmod`
    (+ the autograd Functions of `rmod`, a renderer.py-compatible module)."""

    def __init__(self, gmod, rmod, width, height, fx, fy, rot, tran, near=0.3, thresh=0.05,
                 scale_activation="abs", use_sh_coeff=False, fast=True):
        self.g, self.r = gmod, rmod
        self.width, self.height, self.fx, self.fy = int(width), int(height), float(fx), float(fy)
        self.rot, self.tran = rot, tran                      # CUDA tensors
        self.near, self.thresh = near, thresh
        self.scale_activation, self.use_sh, self.fast = scale_activation, use_sh_coeff, fast
        self.Wp = int(math.ceil(self.width / 16)) * 16
        self.Hp = int(math.ceil(self.height / 16)) * 16
        self.ntx, self.nty = self.Wp // 16, self.Hp // 16
        dev = rot.device
        # Tiles.create_tiles (splatter.py:274-301)
        left = torch.linspace(-self.Wp / 2, self.Wp / 2, self.ntx + 1, device=dev)[:-1]
        top = torch.linspace(-self.Hp / 2, self.Hp / 2, self.nty + 1, device=dev)[:-1]
        t = gmod.Tiles()
        t.left = (left / self.fx).repeat(self.nty).contiguous()
        t.right = ((left + 16) / self.fx).repeat(self.nty).contiguous()
        t.top = (top / self.fy).repeat_interleave(self.ntx).contiguous()
        t.bottom = ((top + 16) / self.fy).repeat_interleave(self.ntx).contiguous()
        self.tiles_cpp = t
        self.lx, self.ly = 16 / self.fx, 16 / self.fy
        self.leftmost, self.topmost = -self.Wp / 2 / self.fx, -self.Hp / 2 / self.fy
        # RayInfo (splatter.py:305-321)
        c2w = torch.inverse(rot)
        self.rays_o = -c2w @ tran
        lt = torch.tensor([(-self.Wp / 2 + 0.5) / self.fx, (-self.Hp / 2 + 0.5) / self.fy, 1.0], device=dev)
        self.lefttop = c2w @ (lt - tran)
        self.dx = c2w @ torch.tensor([1.0 / self.fx, 0, 0], device=dev)
        self.dy = c2w @ torch.tensor([0, 1.0 / self.fy, 0], device=dev)
        self.aux = {}

    def __call__(self, pos, rgb, opa, quat, scale):
        g, r = self.g, self.r
        # project_and_culling (splatter.py:513-542)
        nq = quat / quat.norm(dim=1, keepdim=True)
        ns = scale.abs() + 1e-4 if self.scale_activation == "abs" else r.trunc_exp(scale)
        _pos, _cov, mask = r.global_culling(pos, nq, ns, self.rot.detach(), self.tran.detach(), self.near,
                                            self.width * 1.2 / 2 / self.fx, self.height * 1.2 / 2 / self.fy)
        mb = mask.bool()
        c_pos, c_cov = _pos[mb], _cov[mb]
        c_rgb = rgb[mb] if self.use_sh else rgb[mb].sigmoid()
        c_opa = opa[mb].sigmoid()
        self.aux = dict(mask=mask)
        dev = pos.device
        zero_img = lambda: torch.zeros(self.Hp, self.Wp, 3, device=dev, dtype=torch.float32)
        if c_pos.shape[0] == 0:
            return self._finish(zero_img())
        # render (splatter.py:563-634)
        T = self.ntx * self.nty
        tile_n_point = torch.zeros(T, device=dev, dtype=torch.int32)
        MAXP = c_pos.shape[0] // 20
        tile_list = torch.ones(T, MAXP, device=dev, dtype=torch.int32) * -1
        cobj = g.Gaussian3ds()
        cobj.pos, cobj.rgb, cobj.opa, cobj.cov = c_pos.clone(), c_rgb.clone(), c_opa.clone(), c_cov.clone()
        g.calc_tile_list(cobj, self.tiles_cpp, tile_n_point, tile_list, self.thresh, 2, self.lx, self.ly,
                         self.ntx, self.nty, self.leftmost, self.topmost)
        tile_n_point = torch.min(tile_n_point, torch.ones_like(tile_n_point) * MAXP)
        if tile_n_point.sum() == 0:
            return self._finish(zero_img())
        M = tile_n_point.sum()
        gathered = torch.empty(M, dtype=torch.int32, device=dev)
        tile_ids = torch.empty(M, dtype=torch.int32, device=dev)
        accum = torch.cat([torch.Tensor([0]).to(dev), torch.cumsum(tile_n_point, 0)]).to(tile_n_point)
        maxcnt = tile_n_point.max().item()
        g.gather_gaussians(accum, tile_list, gathered, tile_ids, int(maxcnt))
        gi = gathered.long()
        t_pos, t_rgb, t_opa, t_cov = c_pos[gi], c_rgb[gi], c_opa[gi], c_cov[gi]
        BASE = t_pos[..., 2].max()
        key = t_pos[..., 2].to(torch.float32) + tile_ids.to(torch.float32) * (BASE + 1)
        _, order = torch.sort(key)
        t_pos, t_rgb, t_opa, t_cov = t_pos[order], t_rgb[order], t_opa[order], t_cov[order]
        self.aux.update(accum=accum, n_instances=int(M), max_tile=int(maxcnt), MAXP=MAXP,
                        sorted=(t_pos, t_rgb, t_opa, t_cov))
        img = r.draw(t_pos, t_rgb, t_opa, t_cov, accum, self.Hp, self.Wp, self.fx, self.fy, False, False,
                     self.use_sh, self.fast, self.rays_o, self.lefttop, self.dx, self.dy)
        return self._finish(img)

    def _finish(self, padded):
        self.aux["padded"] = padded
        top = (self.Hp - self.height) // 2
        left = (self.Wp - self.width) // 2
        return torch.clamp(padded, 0, 1)[top:top + self.height, left:left + self.width, :]   # splatter.py:652-653
