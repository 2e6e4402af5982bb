"""Seeded synthetic scenes for parity tests and benchmarks (no datasets are available offline).

Follows SURVEY.md section 8(d): a pin-hole camera built with the reference's own conventions
(src/diff_recon/utils/camera.py:6-35,109-117 -- world_view_transform is the TRANSPOSE of the W2C
matrix, full_proj = view @ proj^T, znear=1, zfar=1000) looking down +z from (0,0,-4), and triangles
whose centres fill the frustum slab z_view in [2,8] (1.1x over-scan so ~17% are culled) with
log-normal pixel footprints.  Everything is generated on the CPU with a seeded torch.Generator, so
the GPU box and this container produce bit-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

ZNEAR, ZFAR = 1.0, 1000.0  # camera.py:109-110


def make_camera(width: int, height: int, focal_px: Optional[float] = None, cam_z: float = -4.0) -> Dict[str, object]:
    """Camera tensors in the layout TriangleRasterizationSettings expects (all CPU float32)."""
    f = 1.2 * width if focal_px is None else float(focal_px)
    fovx = 2.0 * math.atan(width / (2.0 * f))
    fovy = 2.0 * math.atan(height / (2.0 * f))
    tanx, tany = math.tan(fovx / 2), math.tan(fovy / 2)
    # W2C = [R^T | t] with R = I, t = (0, 0, -cam_z): p_view = p_world + t
    w2c = torch.eye(4, dtype=torch.float32)
    w2c[2, 3] = -cam_z
    view = w2c.transpose(0, 1).contiguous()  # "world_view_transform"
    top, right = tany * ZNEAR, tanx * ZNEAR
    proj = torch.zeros(4, 4, dtype=torch.float32)
    proj[0, 0] = 2.0 * ZNEAR / (2 * right)
    proj[1, 1] = 2.0 * ZNEAR / (2 * top)
    proj[3, 2] = 1.0
    proj[2, 2] = ZFAR / (ZFAR - ZNEAR)
    proj[2, 3] = -(ZFAR * ZNEAR) / (ZFAR - ZNEAR)
    projT = proj.transpose(0, 1)
    full = (view.unsqueeze(0).bmm(projT.unsqueeze(0))).squeeze(0).contiguous()
    campos = view.inverse()[3, :3].contiguous()
    return dict(image_width=int(width), image_height=int(height), tanfovx=tanx, tanfovy=tany, viewmatrix=view, projmatrix=full,
                campos=campos)


@dataclass
class Scene:
    name: str
    cam: Dict[str, object]
    vertex: torch.Tensor          # (P,3,3)
    opacity: torch.Tensor         # (P,1) post-sigmoid
    shs: Optional[torch.Tensor]   # (P,M,3) or None
    feature: Optional[torch.Tensor]  # (P,C) or None
    sh_degree: int
    gamma: float
    background: torch.Tensor      # (C,)
    background_depth: float
    rich_info: bool
    back_culling: bool = False
    grads: Dict[str, torch.Tensor] = field(default_factory=dict)  # upstream dL_dout_*

    @property
    def P(self) -> int:
        return int(self.vertex.shape[0])

    def to(self, device) -> "Scene":
        mv = lambda t: None if t is None else t.to(device)
        cam = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.cam.items()}
        return Scene(self.name, cam, mv(self.vertex), mv(self.opacity), mv(self.shs), mv(self.feature), self.sh_degree, self.gamma,
                     mv(self.background), self.background_depth, self.rich_info, self.back_culling,
                     {k: v.to(device) for k, v in self.grads.items()})

    def settings_kwargs(self, debug: bool = False) -> Dict[str, object]:
        return dict(image_width=self.cam["image_width"], image_height=self.cam["image_height"], tanfovx=self.cam["tanfovx"],
                    tanfovy=self.cam["tanfovy"], viewmatrix=self.cam["viewmatrix"], projmatrix=self.cam["projmatrix"],
                    campos=self.cam["campos"], sh_degree=self.sh_degree, gamma=self.gamma, scale_modifier=1.0,
                    background_depth=self.background_depth, background=self.background, back_culling=self.back_culling,
                    rich_info=self.rich_info, debug=debug)


def make_scene(name: str, P: int, width: int, height: int, sh_degree: int = 0, M: Optional[int] = None, rich_info: bool = False,
               gamma: float = 1.0, seed: int = 0, rho_px: float = 3.0, rho_sigma: float = 0.6, use_feature: bool = False, channels: int = 3,
               back_culling: bool = False, geometry_grads: bool = False, opacity_ste: Optional[float] = None) -> Scene:
    g = torch.Generator().manual_seed(seed)
    cam = make_camera(width, height)
    tanx, tany = cam["tanfovx"], cam["tanfovy"]
    zv = 2.0 + 6.0 * torch.rand(P, generator=g)
    cx = (2 * torch.rand(P, generator=g) - 1) * 1.1 * tanx * zv
    cy = (2 * torch.rand(P, generator=g) - 1) * 1.1 * tany * zv
    centre = torch.stack([cx, cy, zv - 4.0], dim=1)  # world z = z_view + cam_z
    rho = torch.exp(math.log(rho_px) + rho_sigma * torch.randn(P, generator=g))
    s = rho * zv * 2.0 * tanx / width
    r = torch.randn(P, 3, 3, generator=g) * s[:, None, None]
    r = r - r.mean(dim=1, keepdim=True)
    vertex = (centre[:, None, :] + r).float().contiguous()
    opacity = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g)).float()
    if opacity_ste is not None:  # mesh configs binarise opacity (VanillaTS_model.py:620-621)
        opacity = (opacity > opacity_ste).float()
    shs = feature = None
    if use_feature:
        feature = torch.rand(P, channels, generator=g).float()
        C = channels
    else:
        M = (sh_degree + 1) ** 2 if M is None else M
        f_dc = 3.0 * torch.rand(P, 1, 3, generator=g) - 1.5
        f_rest = 0.1 * torch.randn(P, M - 1, 3, generator=g)
        shs = torch.cat([f_dc, f_rest], dim=1).float().contiguous()
        C = 3
    background = torch.rand(C, generator=g).float()
    bg_depth = float((vertex - cam["campos"][None, None, :]).norm(dim=-1).max()) if P > 0 else 5000.0
    grads = {"dL_dout_feature": (torch.rand(C, height, width, generator=g) / (height * width)).float()}
    if rich_info:
        if geometry_grads:
            grads["dL_dout_depth"] = (torch.rand(height, width, generator=g) / (height * width)).float()
            grads["dL_dout_normal"] = (torch.rand(3, height, width, generator=g) / (height * width)).float()
        else:
            grads["dL_dout_depth"] = torch.zeros(height, width)
            grads["dL_dout_normal"] = torch.zeros(3, height, width)
    return Scene(name, cam, vertex, opacity, shs, feature, sh_degree, gamma, background, bg_depth, rich_info, back_culling, grads)


# BASELINE.json configs (SURVEY.md section 8 header / BASELINE.md section 3)
CONFIGS = {
    "C1": dict(P=10_000, width=256, height=256, sh_degree=0, rich_info=False),
    "C2": dict(P=300_000, width=800, height=800, sh_degree=3, rich_info=True),
    "C3": dict(P=1_500_000, width=1920, height=1080, sh_degree=3, rich_info=True),
    "C4": dict(P=100_000, width=1600, height=1600, sh_degree=0, rich_info=True, gamma=7.0, opacity_ste=0.3),
    "C5": dict(P=5_000_000, width=1920, height=1080, sh_degree=0, rich_info=True, geometry_grads=True),
}


def make_config(name: str, seed: int = 0, **overrides) -> Scene:
    kw = dict(CONFIGS[name])
    kw.update(overrides)
    return make_scene(name, seed=seed, **kw)
