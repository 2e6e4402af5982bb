#!/usr/bin/env python3
"""bench.py -- rasterizer fwd+bwd frames/s on the BASELINE.json workload (default C3: 1.5 M triangles, 1920x1080,
SH degree 3, rich_info=True -- what training runs), synthetic seeded data (triangle_splatting_b200/scenes.py).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one forward + backward of the rasterizer through the reference-shaped public API
(TriangleRasterizer / _RasterizeTriangles autograd.Function -> C ABI of libts2d.so).
  value    whole-job frames/s, parameters and upstream gradients resident in HBM, CUDA events, max over ranks.
  e2e      the same metric for a training-style step driven from HOST buffers: per step the camera tensors and a
           ground-truth image are copied from pinned host memory, an L1 loss against it drives backward, and the
           scalar loss is read back (bytes counted from the tensors copied).  `e2e_all_host` additionally moves
           every parameter tensor H2D and every gradient tensor D2H inside the timed region.
  roofline dominant kernel (largest per-stage device time, CUDA events recorded inside libts2d around each stage).
  cpu_baseline  the CPU oracle port (oracle/, OpenMP) timed on a bounded sample of the same workload.
--impl reference times the UNMODIFIED reference CUDA extension (oracle/_ref, built from /root/reference by
oracle/build_ref.py) on the same scene through its own pybind entry points; if that module cannot be loaded
it falls back to the CPU oracle port on a bounded sample and says so.
N > 1: image-space tile sharding of ONE render (tile % N == rank); pixels and contrib statistics travel through NVLink peer memory
inside the forward kernel when the fabric is available (else one NCCL all-reduce), the per-triangle gradient accumulators are summed
with one all-reduce; strong scaling.  `--check` (default on for N > 1) compares the sharded frame and gradients with the single-GPU
result on every rank, outside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


# ----------------------------------------------------------------------------------------------- utilities
class ClockSampler:
    """nvidia-smi clocks/throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def algorithmic_bytes(P, V, R, N, T, K, M, rich, rows=0):
    """bytes each stage must touch once (no temp / zero-fill / sort-pass amplification)."""
    rho = 1 if rich else 0
    rec = 44 + 24 * rho
    return {
        "preprocess": 36 * P + 12 * K * V + (rec + 28) * V + 8 * P,
        "order_scan": 8 * P + 8 * P,
        "binning": 28 * V + 12 * R + 24 * R + 8 * R + 8 * T,
        "render_fwd": (4 + rec) * R + (20 + 16 * rho) * N + 8 * rho * V,
        "render_bwd": (4 + rec) * R + (20 + 16 * rho) * N + (40 + 24 * rho) * V,
        "preprocess_bwd": (40 + 24 * rho) * V + (36 + 12 * K + 3) * V + 4 * P + (36 + 8 + 4 + 12 + 12 * M) * P,
        # not in SURVEY 8(d) (the reference has no counterpart): bookkeeping of the atomics-free gradient write-back, counted as what
        # it must touch once -- a key, a list entry and an index per instance / one 64 B row per (sub-tile, entry) pair in and out
        "bwd_prepare": 8 * R + 9 * R,
        "bwd_reduce": 64 * rows + 64 * P,
    }


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def scene_for(config: str):
    from triangle_splatting_b200.scenes import make_config

    return make_config(config)


# --------------------------------------------------------------------------------------------- step builders
_COPY_STREAMS = {}


def step_start_mark(dev):
    """Event on the current stream at the start of a step: everything the previous step queued lies in front of it."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    return ev


def prefetch_gt(host, dev, after):
    """H2D copy of this step's ground-truth image on a side stream (inside the timed region, both arms): the image is not
    needed before the loss, so the PCIe transfer overlaps the forward pass.  Called AFTER the forward pass has been queued (the
    device idles until the step's first kernel arrives, so nothing is put in front of that); `after` = step_start_mark() keeps the
    copy ordered behind the previous step only, not behind this step's forward.  Returns a callable that makes the current stream
    wait for the copy and hands back the device tensor."""
    cs = _COPY_STREAMS.get(dev)
    if cs is None:
        cs = _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    cs.wait_event(after)  # ordered after the previous step's consumers of the buffer the allocator may hand back
    with torch.cuda.stream(cs):
        gt = host["gt"].to(dev, non_blocking=True)
    done = torch.cuda.Event()
    done.record(cs)

    def wait():
        cur.wait_event(done)
        gt.record_stream(cur)
        return gt

    return wait


class OursStep:
    """fwd+bwd through TriangleRasterizer (autograd) exactly as diff_recon's TriangleRenderer calls it."""

    def __init__(self, sc, dev, primitive="2D"):
        from triangle_splatting_b200 import TriangleRasterizationSettings
        if primitive == "3D":
            from triangle_splatting_b200 import TriangleRasterizer3D as TriangleRasterizer
        else:
            from triangle_splatting_b200 import TriangleRasterizer

        self.sc = sc.to(dev)
        self.dev = dev
        self.vertex = self.sc.vertex.clone().requires_grad_(True)
        self.shs = self.sc.shs.clone().requires_grad_(True)
        self.opacity = self.sc.opacity.clone().requires_grad_(True)
        self.settings_cls, self.rast_cls = TriangleRasterizationSettings, TriangleRasterizer
        self.rast = TriangleRasterizer(raster_settings=TriangleRasterizationSettings(**self.sc.settings_kwargs()))
        self.g = [self.sc.grads["dL_dout_feature"], self.sc.grads["dL_dout_depth"], self.sc.grads["dL_dout_normal"]]

    def forward(self, rast=None):
        center2D = torch.zeros((self.sc.P, 2), device=self.dev, requires_grad=True)
        return (rast or self.rast).forward(vertex=self.vertex, center2D=center2D, opacity=self.opacity, shs=self.shs, feature=None)

    def __call__(self):
        self.vertex.grad = self.shs.grad = self.opacity.grad = None
        out = self.forward()
        torch.autograd.backward([out[0], out[2], out[3]], self.g)
        return out

    def e2e_step(self, host):
        """camera + GT image from pinned host memory -> forward -> L1 loss -> backward -> loss to host."""
        self.vertex.grad = self.shs.grad = self.opacity.grad = None
        mark = step_start_mark(self.dev)
        cam = {k: host[k].to(self.dev, non_blocking=True) for k in ("viewmatrix", "projmatrix", "campos", "background")}
        kw = self.sc.settings_kwargs()
        kw.update(cam)
        out = self.forward(self.rast_cls(raster_settings=self.settings_cls(**kw)))
        gt = prefetch_gt(host, self.dev, mark)()
        loss = (out[0] - gt).abs().mean()
        loss.backward()
        return float(loss.item())

    def e2e_all_host_step(self, host):
        cam = {k: host[k].to(self.dev, non_blocking=True) for k in ("viewmatrix", "projmatrix", "campos", "background")}
        v = host["vertex"].to(self.dev, non_blocking=True).requires_grad_(True)
        s = host["shs"].to(self.dev, non_blocking=True).requires_grad_(True)
        o = host["opacity"].to(self.dev, non_blocking=True).requires_grad_(True)
        g = [host[k].to(self.dev, non_blocking=True) for k in ("g_feature", "g_depth", "g_normal")]
        kw = self.sc.settings_kwargs()
        kw.update(cam)
        c2d = torch.zeros((self.sc.P, 2), device=self.dev, requires_grad=True)
        out = self.rast_cls(raster_settings=self.settings_cls(**kw)).forward(vertex=v, center2D=c2d, opacity=o, shs=s, feature=None)
        torch.autograd.backward([out[0], out[2], out[3]], g)
        host["o_image"].copy_(out[0], non_blocking=True)
        host["o_gv"].copy_(v.grad, non_blocking=True)
        host["o_gs"].copy_(s.grad, non_blocking=True)
        host["o_go"].copy_(o.grad, non_blocking=True)
        host["o_gc"].copy_(c2d.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()


class ReferenceStep:
    """The unmodified reference extension (oracle/_ref) driven through its own pybind entry points, wrapped in the
    same autograd.Function shape as the reference's Python package (R2D/diff_triangle_rasterization_2D/__init__.py)."""

    def __init__(self, sc, dev, ref):
        self.sc = sc.to(dev)
        self.dev, self.ref = dev, ref
        self.vertex = self.sc.vertex.clone().requires_grad_(True)
        self.shs = self.sc.shs.clone().requires_grad_(True)
        self.opacity = self.sc.opacity.clone().requires_grad_(True)
        self.g = [self.sc.grads["dL_dout_feature"], self.sc.grads["dL_dout_depth"], self.sc.grads["dL_dout_normal"]]
        refmod = ref

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, vertex, center2D, shs, feature, opacity, s):
                args = (s["image_width"], s["image_height"], s["tanfovx"], s["tanfovy"], s["viewmatrix"].contiguous(),
                        s["projmatrix"].contiguous(), s["campos"].contiguous(), s["sh_degree"], s["gamma"], s["scale_modifier"],
                        float(s["background_depth"]), s["background"].contiguous(), vertex, shs, feature, opacity, s["back_culling"],
                        s["rich_info"], s["debug"])
                (R, out_feature, radii, depth, normal, csum, cmax, gb, bb, ib) = refmod.rasterize_triangles(*args)
                ctx.s, ctx.R = s, R
                ctx.save_for_backward(vertex, shs, feature, opacity, radii, gb, bb, ib)
                return out_feature, radii, depth, normal, csum, cmax

            @staticmethod
            def backward(ctx, g_feat, _r, g_depth, g_normal, _a, _b):
                s = ctx.s
                vertex, shs, feature, opacity, radii, gb, bb, ib = ctx.saved_tensors
                args = (s["tanfovx"], s["tanfovy"], s["viewmatrix"].contiguous(), s["projmatrix"].contiguous(), s["campos"].contiguous(),
                        s["sh_degree"], s["gamma"], s["scale_modifier"], float(s["background_depth"]), s["background"].contiguous(), vertex,
                        shs, feature, opacity, ctx.R, radii, gb, bb, ib, g_feat.contiguous(), g_depth.contiguous(), g_normal.contiguous(),
                        s["rich_info"], s["debug"])
                gv, gc, gs, gf, go = refmod.rasterize_triangles_backward(*args)
                return gv, gc, gs, None, go, None

        self.fn = _Fn
        self.empty = torch.empty(0, device=dev)

    def forward(self, kw=None):
        center2D = torch.zeros((self.sc.P, 2), device=self.dev, requires_grad=True)
        return self.fn.apply(self.vertex, center2D, self.shs, self.empty, self.opacity, kw or self.sc.settings_kwargs())

    def __call__(self):
        self.vertex.grad = self.shs.grad = self.opacity.grad = None
        out = self.forward()
        torch.autograd.backward([out[0], out[2], out[3]], self.g)
        return out

    def e2e_step(self, host):
        self.vertex.grad = self.shs.grad = self.opacity.grad = None
        mark = step_start_mark(self.dev)
        cam = {k: host[k].to(self.dev, non_blocking=True) for k in ("viewmatrix", "projmatrix", "campos", "background")}
        kw = self.sc.settings_kwargs()
        kw.update(cam)
        out = self.forward(kw)
        gt = prefetch_gt(host, self.dev, mark)()  # same order as our arm
        loss = (out[0] - gt).abs().mean()
        loss.backward()
        return float(loss.item())


def model_params(sc, dev):
    """Raw model parameters for a scene (what VanillaTSModel holds): _vertex, _f_dc, _f_rest, _opacity (logits)."""
    s = sc.to(dev)
    leaf = lambda t: t.contiguous().clone().requires_grad_(True)
    return dict(vertex=leaf(s.vertex), f_dc=leaf(s.shs[:, :1, :]), f_rest=leaf(s.shs[:, 1:, :]),
                opacity=leaf(torch.logit(s.opacity.clamp(1e-6, 1 - 1e-6))))


class ModelStepOurs:
    """Model-level step through the fused front-end: raw parameters in, gradients w.r.t. them out;
    sigmoid, SH concat, background depth and the training statistics happen inside K1 / K9 -- no host sync besides num_rendered."""

    def __init__(self, sc, dev, primitive):
        from triangle_splatting_b200 import TrainingStatistics, TriangleModelRasterizer, TriangleRasterizationSettings

        self.sc, self.dev = sc.to(dev), dev
        self.p = model_params(sc, dev)
        self.stats = TrainingStatistics(sc.P, dev)
        self.rast = TriangleModelRasterizer(TriangleRasterizationSettings(**self.sc.settings_kwargs()), primitive=primitive, statistics=self.stats)
        self.g = [self.sc.grads["dL_dout_feature"], self.sc.grads["dL_dout_depth"], self.sc.grads["dL_dout_normal"]]

    def __call__(self):
        p = self.p
        for t in p.values():
            t.grad = None
        c2d = torch.zeros((self.sc.P, 2), device=self.dev, requires_grad=True)
        out = self.rast.forward(p["vertex"], c2d, p["opacity"], p["f_dc"], p["f_rest"] if p["f_rest"].shape[1] else None)
        torch.autograd.backward([out[0], out[2], out[3]], self.g)


class ModelStepReference:
    """The reference's own flow for the same step: VanillaTS_model.py:608-647 preamble in torch (cat, sigmoid, bg_depth with its
    implicit .item()), the unmodified reference extension, and _training_statistic (:347-363)."""

    def __init__(self, sc, dev, ref, primitive):
        self.inner = ReferenceStep(sc, dev, ref)
        self.sc, self.dev = self.inner.sc, dev
        self.p = model_params(sc, dev)
        self.st = {k: torch.zeros(sc.P, device=dev) for k in ("gradient_accum", "gradient_denom", "max_radii2D", "contrib_sum", "contrib_max",
                                                              "contrib_denom")}
        self.g = self.inner.g

    def __call__(self):
        p, st = self.p, self.st
        for t in p.values():
            t.grad = None
        vertex = p["vertex"]
        shs = torch.cat((p["f_dc"], p["f_rest"]), dim=1)
        opacity = torch.sigmoid(p["opacity"])
        bg_depth = (self.sc.cam["campos"].view(1, 1, 3) - vertex).norm(dim=-1).max()
        kw = self.sc.settings_kwargs()
        kw["background_depth"] = bg_depth  # float(...) inside the op wrapper: the host sync the reference has
        center2D = torch.zeros((self.sc.P, 2), device=self.dev, requires_grad=True)
        out = self.inner.fn.apply(vertex, center2D, shs, self.inner.empty, opacity, kw)
        torch.autograd.backward([out[0], out[2], out[3]], self.g)
        radii, contrib_sum, contrib_max = out[1], out[4], out[5]
        visible_mask = radii > 0
        st["gradient_accum"][visible_mask] += torch.norm(center2D.grad[visible_mask, :2], dim=-1)
        st["gradient_denom"][visible_mask] += 1
        st["contrib_sum"][visible_mask] = torch.max(st["contrib_sum"][visible_mask], contrib_sum[visible_mask])
        st["contrib_max"][visible_mask] = torch.max(st["contrib_max"][visible_mask], contrib_max[visible_mask])
        st["contrib_denom"][visible_mask] += 1
        st["max_radii2D"][visible_mask] = torch.max(st["max_radii2D"][visible_mask], radii[visible_mask])


def reference_image_loss(image, gt, w_ssim, kernel_cache={}):
    """The trainer's pixel-wise loss as the reference computes it (trainer_utils.py:9-103 SSIMLoss with its 11x11 Gaussian window,
    :323-324 L1; weights as VanillaTS_trainer.py:71-75,108), torch ops with torch's default TF32 setting for convolutions."""
    import torch.nn.functional as F

    ch = image.shape[0]
    key = (ch, image.device)
    if key not in kernel_cache:
        x_grid = torch.arange(11).unsqueeze(0).repeat(11, 1)
        xy = torch.stack([x_grid, x_grid.T], dim=-1).float()
        k = torch.exp(-(xy - 5.0).pow(2).sum(dim=-1) / (2 * 1.5**2.0))
        kernel_cache[key] = (k / k.sum()).unsqueeze(0).unsqueeze(0).float().repeat(ch, 1, 1, 1).to(image.device)
    kernel = kernel_cache[key]
    img1, img2 = image.unsqueeze(0), gt.unsqueeze(0)
    window = lambda t: F.conv2d(t, kernel, padding=5, groups=ch)
    mu1, mu2 = window(img1), window(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = window(img1 * img1) - mu1_sq
    sigma2_sq = window(img2 * img2) - mu2_sq
    sigma12 = window(img1 * img2) - mu1_mu2
    C1, C2 = 0.01**2, 0.03**2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return (1.0 - w_ssim) * torch.abs(image - gt).mean() + w_ssim * (1 - ssim_map.mean())


def image_loss_ms(sc, dev, fused: bool, steps=10, w_ssim=0.2):
    """fwd + bwd of the pixel-wise image loss on a frame of the workload's size (median of per-step CUDA-event times)."""
    g = torch.Generator().manual_seed(4321)
    h, w = sc.cam["image_height"], sc.cam["image_width"]
    gt = torch.rand(3, h, w, generator=g).to(dev)
    img = (gt + 0.1 * torch.randn(3, h, w, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
    if fused:
        from triangle_splatting_b200 import image_loss

        fn = lambda: image_loss(img, gt, w_ssim)
    else:
        fn = lambda: reference_image_loss(img, gt, w_ssim)

    def step():
        img.grad = None
        fn().backward()

    med, mean = median_step_ms(step, steps, 3, dev, 1)
    return {"ms": med, "mean_ms": mean, "steps": steps, "w_ssim": w_ssim,
            "what": ("fused L1 + SSIM kernels of libts2d (ts2d_image_loss_forward / _backward)" if fused else
                     "the reference's torch composition (L1 + SSIMLoss, five depth-wise 11x11 convolutions; cuDNN, TF32 allowed)") +
                    f", fwd + bwd on a 3x{h}x{w} frame"}


def reference_geometry_loss(depth, normal, tan_fovx, tan_fovy, scale_factor=0.5, q=0.9):
    """The trainer's geometry term as the reference computes it (trainer_utils.py:159-185 ScharrFilter, :212-255 DepthNormalLoss; the
    shipped MatrixCity config: scale_factor 0.5), torch ops."""
    import torch.nn.functional as F

    kx = torch.tensor([[-3, 0, 3], [-10, 0, 10], [-3, 0, 3]], dtype=torch.float32, device=depth.device).view(1, 1, 3, 3) / 32
    ky = torch.tensor([[-3, -10, -3], [0, 0, 0], [3, 10, 3]], dtype=torch.float32, device=depth.device).view(1, 1, 3, 3) / 32
    W0, H0 = depth.shape[-1], depth.shape[-2]
    d = depth.unsqueeze(0).unsqueeze(0)
    if scale_factor is not None and scale_factor != 1:
        d = F.interpolate(d, scale_factor=scale_factor, mode="bilinear", align_corners=False)
    grad = torch.cat((F.conv2d(d, kx, padding=1), F.conv2d(d, ky, padding=1)), dim=1).squeeze(0)
    Dx, Dy = torch.unbind(grad / d.squeeze(0), 0)
    W, H = d.shape[-1], d.shape[-2]
    x, y = torch.meshgrid(torch.arange(W, dtype=torch.float32, device=depth.device), torch.arange(H, dtype=torch.float32, device=depth.device), indexing="xy")
    nrm = torch.stack([W * Dx / (2 * tan_fovx), H * Dy / (2 * tan_fovy), -(1 + (x - W / 2 + 0.5) * Dx + (y - H / 2 + 0.5) * Dy)], dim=0)
    gnorm = grad.norm(dim=0, keepdim=True)
    if W0 != W or H0 != H:
        nrm = F.interpolate(nrm.unsqueeze(0), size=(H0, W0), mode="bilinear", align_corners=False).squeeze(0)
        gnorm = F.interpolate(gnorm.unsqueeze(0), size=(H0, W0), mode="bilinear", align_corners=False).squeeze(0)
    nrm = nrm / nrm.norm(dim=0, keepdim=True)
    mask = (gnorm < torch.quantile(gnorm, q)).float().squeeze(0)
    normal = F.normalize(normal, p=2, dim=0, eps=1e-8)
    return ((1 - (normal * nrm).sum(dim=0)) * mask).mean()


def geometry_loss_ms(sc, dev, fused: bool, steps=10):
    """fwd + bwd of the depth-normal consistency loss on a frame of the workload's size (median of per-step CUDA-event times)."""
    g = torch.Generator().manual_seed(99)
    h, w = sc.cam["image_height"], sc.cam["image_width"]
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    depth = (4.0 + 0.002 * xx + 0.003 * yy + 0.5 * torch.sin(xx / 70.0) * torch.cos(yy / 50.0) + 0.01 * torch.rand(h, w, generator=g)).to(dev).requires_grad_(True)
    normal = (torch.randn(3, h, w, generator=g) * 0.3 + torch.tensor([0.1, -0.2, -1.0]).view(3, 1, 1)).to(dev).requires_grad_(True)
    tfx, tfy = float(sc.cam["tanfovx"]), float(sc.cam["tanfovy"])
    if fused:
        from triangle_splatting_b200 import depth_normal_loss

        fn = lambda: depth_normal_loss(depth, normal, tfx, tfy, 0.5)
    else:
        fn = lambda: reference_geometry_loss(depth, normal, tfx, tfy, 0.5)

    def step():
        depth.grad = normal.grad = None
        fn().backward()

    med, mean = median_step_ms(step, steps, 3, dev, 1)
    return {"ms": med, "mean_ms": mean, "steps": steps, "scale_factor": 0.5,
            "what": ("fused depth-normal loss kernels of libts2d (ts2d_depth_normal_loss_forward / _backward: radix select instead of a sort)" if fused else
                     "the reference's torch composition (DepthNormalLoss: interpolate, Scharr convolutions, torch.quantile, normalise, masked mean)") +
                    f", fwd + bwd on a {h}x{w} depth + normal frame"}


MODEL_STEP_WHAT = ("raw parameters (_vertex, _f_dc, _f_rest, _opacity logits) -> opacity activation, SH concat, background depth -> "
                   "rasterizer fwd + bwd -> gradients w.r.t. the raw parameters + _training_statistic; CUDA events, parameters resident")


def pinned_host_buffers(sc, all_host: bool):
    pin = lambda t: t.detach().cpu().contiguous().pin_memory()
    h = {k: pin(sc.cam[k]) for k in ("viewmatrix", "projmatrix", "campos")}
    h["background"] = pin(sc.background)
    g = torch.Generator().manual_seed(1234)
    h["gt"] = torch.rand(sc.background.numel(), sc.cam["image_height"], sc.cam["image_width"], generator=g).pin_memory()
    if all_host:
        h.update(vertex=pin(sc.vertex), shs=pin(sc.shs), opacity=pin(sc.opacity), g_feature=pin(sc.grads["dL_dout_feature"]),
                 g_depth=pin(sc.grads["dL_dout_depth"]), g_normal=pin(sc.grads["dL_dout_normal"]))
        P = sc.P
        h.update(o_image=torch.empty_like(h["gt"]).pin_memory(), o_gv=torch.empty(P, 3, 3).pin_memory(),
                 o_gs=torch.empty_like(h["shs"]).pin_memory(), o_go=torch.empty(P, 1).pin_memory(), o_gc=torch.empty(P, 2).pin_memory())
    return h


# --------------------------------------------------------------------------------------------- CPU baseline
def cpu_oracle_baseline(sc, tile_step=None, budget_s=20.0, primitive="2D"):
    """Oracle port (oracle/ts2d_oracle.c, OpenMP over tiles) on a bounded sample: full per-triangle stages + binning,
    composite fwd+bwd on every `tile_step`-th tile, extrapolated to the whole frame."""

    from oracle.oracle import Oracle

    o = Oracle("f32")
    kw = sc.settings_kwargs()
    kw.pop("debug")
    kw = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    arrs = dict(vertex=sc.vertex.cpu().numpy(), shs=sc.shs.cpu().numpy(), feature=None, opacity=sc.opacity.cpu().numpy(), primitive=primitive)
    t0 = time.perf_counter()
    st = o.forward(**kw, **arrs, stages="bin")
    t_geom = time.perf_counter() - t0
    R = st["num_rendered"]
    if tile_step is None:  # pilot on every 16th tile, then size the sample for ~budget_s of composite work
        PILOT = 16
        t0 = time.perf_counter()
        stp = o.forward(**kw, **arrs, tile_step=PILOT, tile_offset=0)
        o.backward(stp, sc.grads["dL_dout_feature"].cpu().numpy(), sc.grads["dL_dout_depth"].cpu().numpy() if sc.rich_info else None,
                   sc.grads["dL_dout_normal"].cpu().numpy() if sc.rich_info else None)
        t_pilot = max(1e-3, time.perf_counter() - t0 - t_geom)
        tile_step = int(min(256, max(1, round(PILOT * t_pilot / max(1.0, budget_s - 2 * t_geom)))))
    t0 = time.perf_counter()
    st = o.forward(**kw, **arrs, tile_step=tile_step, tile_offset=0)
    t_fwd_all = time.perf_counter() - t0  # includes the geometry stages again
    t0 = time.perf_counter()
    o.backward(st, sc.grads["dL_dout_feature"].cpu().numpy(), sc.grads["dL_dout_depth"].cpu().numpy() if sc.rich_info else None,
               sc.grads["dL_dout_normal"].cpu().numpy() if sc.rich_info else None)
    t_bwd = time.perf_counter() - t0
    t_comp_f = max(0.0, t_fwd_all - t_geom)
    frame_s = t_geom + tile_step * t_comp_f + tile_step * t_bwd
    return {"value": 1.0 / frame_s, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"CPU oracle port (C, OpenMP): per-triangle stages + binning of the full scene ({t_geom:.1f}s) + composite fwd+bwd on "
                      f"every {tile_step}-th tile ({t_comp_f:.1f}s + {t_bwd:.1f}s)" + (f", extrapolated x{tile_step}" if tile_step > 1 else " = the whole frame"),
            "seconds_measured": t_geom + t_fwd_all + t_bwd}


def sharded_vs_single_check(step, sc, dev, primitive, world):
    """Outside the timed region, on EVERY rank: the tile-sharded frame and gradients of `step` against the same step with sharding
    switched off (the whole frame on this GPU).  Returns the worst figures over the ranks and whether all ranks hold identical bits."""
    from triangle_splatting_b200 import distributed as tsd

    def run(st):
        out = st()
        res = {"image": out[0], "depth": out[2], "normal": out[3], "contrib_sum": out[4], "contrib_max": out[5], "radii": out[1].float(),
               "dL_dvertex": st.vertex.grad, "dL_dshs": st.shs.grad, "dL_dopacity": st.opacity.grad}
        return {k: v.detach().clone() for k, v in res.items()}

    sharded = run(step)
    tsd.disable_tile_sharding()
    single = run(OursStep(sc, dev, primitive))
    tsd.enable_tile_sharding()
    rel = {}
    for k, b in single.items():
        a = sharded[k]
        eps = 1e-3 * b.double().pow(2).mean().sqrt() + 1e-30
        rel[k] = float(((a.double() - b.double()).abs() / torch.maximum(b.double().abs(), eps)).max().item())
    keys = sorted(rel)
    worst = torch.tensor([rel[k] for k in keys], device=dev, dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    sums = torch.stack([sharded[k].double().sum() for k in keys])  # a checksum per tensor: equal on all ranks <=> (practically) identical bits
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return {"ranks": world, "metric": "max |sharded - single| / max(|single|, 1e-3 RMS), worst rank", "rel_err": {k: float(v) for k, v in zip(keys, worst.tolist())},
            "radii_equal": rel["radii"] == 0.0 and float(worst[keys.index("radii")]) == 0.0, "all_ranks_identical": bool(torch.equal(lo, hi))}


def work_statistics(fwd_state, sc, dev, world, stage_ms):
    """SURVEY 8(d) work figures of this rank's shard from the forward pass's own state: list lengths, sum of n_contrib (the pairs the
    reference's loops visit), the 256 R upper bound, the (sub-tile, entry) rows of the backward, and ns per pair of K7 / K8."""
    from triangle_splatting_b200 import _lib

    lib = _lib.load()
    W, H = sc.cam["image_width"], sc.cam["image_height"]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R, gb, bb, ib = int(fwd_state[0]), fwd_state[7], fwd_state[8], fwd_state[9]
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ranges = torch.zeros((gx * gy, 2), device=dev, dtype=torch.int32)
    ncon = torch.zeros((H, W), device=dev, dtype=torch.int32)
    _lib.check(lib.ts2d_export_binning(p(gb), p(bb), bb.numel(), p(ib), sc.P, R, W, H, None, None, p(ranges), stream), "export_binning")
    _lib.check(lib.ts2d_export_image(p(ib), W, H, p(ncon), None, stream), "export_image")
    ctr = getattr(fwd_state[0], "counters", None)
    rows = int(ctr.backward_rows()) if ctr is not None else 0
    lens = (ranges[:, 1] - ranges[:, 0]).long()
    lens = lens[lens > 0] if world > 1 else lens  # sharded: only the tiles this rank owns have lists
    pairs = int(ncon.long().sum().item())
    out = {"num_rendered": R, "tiles_with_lists": int(lens.numel()), "list_len_mean": round(float(lens.float().mean().item()), 2) if lens.numel() else 0.0,
           "list_len_max": int(lens.max().item()) if lens.numel() else 0, "pairs_sum_n_contrib": pairs,
           "n_contrib_mean": round(pairs / max(1, (W * H) // world), 2), "pairs_upper_256R": 256 * R, "backward_rows": rows}
    if pairs:
        out["ns_per_pair_render_fwd"] = round(stage_ms["render_fwd"] * 1e6 / pairs, 5)
        out["ns_per_pair_render_bwd"] = round(stage_ms["render_bwd"] * 1e6 / pairs, 5)
    return out


# ----------------------------------------------------------------------------------------------------- main
def sample_clocks(step, dev, index, timed_ms, steps, sampler_result):
    """nvidia-smi needs ~100 ms per sample; if the timed region was shorter than that, sample the same load in a ~1 s
    probe loop right after it (and say so), so a clock lock / throttle during the measurement is still visible."""
    # the decision and the loop length depend only on timed_ms (identical on every rank): the step may contain collectives
    if timed_ms >= 600.0:
        sampler_result["sampled"] = "during the timed region"
        return sampler_result
    smp = ClockSampler(index)
    smp.start()
    for _ in range(int(min(4000, 1200.0 / max(0.05, timed_ms / max(1, steps))) + 1)):
        step()
    torch.cuda.synchronize(dev)
    res = smp.stop()
    res["sampled"] = f"1.2 s probe loop of the same step right after the timed region (which lasted only {timed_ms:.0f} ms)"
    return res


def timed_region(step, steps, warmup, dev, world):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()  # per-step marks inside the same timed region (no synchronisation): median / p10 / p90 for free
    torch.cuda.synchronize(dev)
    ms = ev[0].elapsed_time(ev[steps])
    per = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)], device=dev)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(per, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    q = torch.quantile(per.float(), torch.tensor([0.1, 0.5, 0.9], device=dev)).tolist()
    timed_region.per_step = {"p10_ms": round(q[0], 4), "median_ms": round(q[1], 4), "p90_ms": round(q[2], 4)}
    return ms


def median_step_ms(step, steps, warmup, dev, world):
    """Per-step CUDA-event times (barrier-free between steps), max over ranks per step, -> (median, mean) in ms.  Used for the
    model-level step, whose torch preamble allocates: a single allocator hiccup must not decide a 5-step average."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(t.median().item()), float(t.mean().item())


def wall_region(fn, steps, warmup, dev):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--primitive", default="2D", choices=["2D", "3D"],
                    help="2D: diff_triangle_rasterization_2D (the north-star path); 3D: diff_triangle_rasterization_3D (the *_mesh configs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the sharded-vs-single-GPU comparison (outside the timed region)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model-step", action="store_true", help="skip the model-level step (fused parameter-space front-end)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "rasterizer fwd+bwd frames/sec @1080p, 1.5M tris; HBM GB/s vs roofline"
    base = {"metric": metric, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}

    sc = scene_for(a.config)
    cfg = {"workload": f"{a.config}: P={sc.P} triangles, {sc.cam['image_width']}x{sc.cam['image_height']}, SH degree {sc.sh_degree} "
                       f"(M={sc.shs.shape[1]}), rich_info={sc.rich_info}, gamma={sc.gamma}, fwd+bwd",
           "primitive": a.primitive, "P": sc.P, "width": sc.cam["image_width"], "height": sc.cam["image_height"], "sh_degree": sc.sh_degree,
           "l2_policy": "inputs larger than L2 (SH 288 MB + 120 MB raster records + instance lists >> 126 MB L2); no explicit flush",
           "parallelism": "single GPU"}  # N > 1: filled in below once the exchange mechanism that actually ran is known

    if not torch.cuda.is_available():
        if a.impl == "reference" and rank == 0:
            cb = cpu_oracle_baseline(sc, primitive=a.primitive)
            line = dict(base, impl="reference", value=cb["value"], ms_per_step=1e3 / cb["value"], config=cfg, cpu_baseline=cb,
                        e2e={"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0,
                        note="no CUDA device and no loadable reference extension: CPU oracle port on a bounded sample")
            print(json.dumps(line))
            return
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    # ------------------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        from oracle import build_ref

        ref = None
        try:
            ref = build_ref.load(a.primitive)
        except Exception as ex:  # noqa: BLE001
            print(f"[bench] reference extension failed to load: {ex}", file=sys.stderr)
        if ref is None:
            cb = cpu_oracle_baseline(sc, primitive=a.primitive)
            line = dict(base, impl="reference", value=cb["value"], ms_per_step=1e3 / cb["value"], config=cfg, cpu_baseline=cb, n_gpus=1,
                        e2e={"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0,
                        note="oracle/_ref not loadable: CPU oracle port on a bounded sample")
            print(json.dumps(line))
            return
        step = ReferenceStep(sc, dev, ref)
        smp = ClockSampler(local_rank)
        smp.start()
        ms = timed_region(step, a.steps, a.warmup, dev, 1)
        clocks = sample_clocks(step, dev, local_rank, ms, a.steps, smp.stop())
        fps = a.steps / (ms / 1e3)
        host = pinned_host_buffers(sc, all_host=False)
        ms_e = wall_region(lambda: step.e2e_step(host), max(3, a.steps // 2), 2, dev)
        e2e_fps = max(3, a.steps // 2) / (ms_e / 1e3)
        h2d = sum(host[k].numel() * 4 for k in ("viewmatrix", "projmatrix", "campos", "background", "gt"))
        model_step = None
        if sc.rich_info and sc.shs is not None and not a.no_model_step:
            mstep = ModelStepReference(sc, dev, ref, a.primitive)
            k_m = max(7, a.steps // 2)
            med, mean = median_step_ms(mstep, k_m, 5, dev, 1)
            model_step = {"value": 1e3 / med, "unit": "frames/s", "ms_per_step": med, "mean_ms_per_step": mean, "steps": k_m,
                          "statistic": "median of per-step CUDA-event times", "what": MODEL_STEP_WHAT}
            del mstep
        img_loss = None
        if not a.no_model_step:
            try:
                img_loss = image_loss_ms(sc, dev, fused=False)
            except Exception as ex:  # noqa: BLE001
                img_loss = {"ms": None, "what": f"failed: {ex}"}
        geo_loss = None
        if not a.no_model_step:
            try:
                geo_loss = geometry_loss_ms(sc, dev, fused=False)
            except Exception as ex:  # noqa: BLE001
                geo_loss = {"ms": None, "what": f"failed: {ex}"}
        line = dict(base, impl="reference", n_gpus=1, value=fps, ms_per_step=ms / a.steps, config=cfg, clocks=clocks, model_step=model_step,
                    image_loss=img_loss, geometry_loss=geo_loss,
                    e2e={"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                    cpu_baseline={"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference",
                                  "sample": "full workload on the reference's own CUDA build (oracle/_ref, sm_100): the reference has no CPU "
                                            "implementation of this path; 1 host thread drives the GPU"},
                    gpu_launches=0, note="unmodified reference extension through its pybind entry points; none of our kernels on this path")
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------ our arm
    if world > 1:
        # rank 0 must print exactly one JSON line on stdout: NCCL writes its version banner to fd 1 at communicator creation
        # whenever NCCL_DEBUG >= VERSION, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        from triangle_splatting_b200 import distributed as tsd

        tsd.enable_tile_sharding()
    from triangle_splatting_b200 import _lib

    lib = _lib.load()
    step = OursStep(sc, dev, a.primitive)
    out = step()  # first call: also gives V, R for the byte model
    torch.cuda.synchronize(dev)
    if world > 1:
        from triangle_splatting_b200 import distributed as tsd

        used_fabric = tsd.fabric(dev) is not None
        cfg["parallelism"] = (f"image-space tile sharding x{world} (tile % N == rank), strong scaling; exchange: "
                              + ("NVLink peer memory, kernels behind the composite kernels on the same stream: multimem.st of the owned tiles' 64-byte "
                                 "rows into every replica; contrib statistics and the 64 B/triangle gradient accumulators combined per home slice "
                                 "inside the switch (multimem.ld_reduce) and written back to every replica; two signal-pad rendezvous per pass"
                                 if used_fabric else "NCCL all-reduce(sum) of the zero-filled frame planes + contrib_sum, all-reduce(max) of contrib_max; "
                                 "NCCL all-reduce(sum) of the 64 B/triangle accumulators between K8 + row reduction and K9"))
        cfg["exchange"] = {"forward": "fabric" if used_fabric else "nccl", "backward": "fabric" if used_fabric else "nccl"}
    radii = out[1]
    V = int((radii > 0).sum().item())
    N = sc.cam["image_width"] * sc.cam["image_height"]
    T = ((sc.cam["image_width"] + 15) // 16) * ((sc.cam["image_height"] + 15) // 16)

    # R of the whole frame (each rank only knows its shard's R)
    from triangle_splatting_b200 import _C as tsC  # noqa: N811
    R_local = None
    smp = ClockSampler(local_rank)
    smp.start()
    ms = timed_region(step, a.steps, a.warmup, dev, world)
    clocks = sample_clocks(step, dev, local_rank, ms, a.steps, smp.stop())
    per_step = dict(getattr(timed_region, "per_step", {}))
    fps = a.steps / (ms / 1e3)
    # per-stage times: a second pass of the same steps with the library's stage events switched on, OUTSIDE the timed region -- the
    # events (two per stage) are stream operations of their own and sit between kernels that otherwise follow each other by
    # programmatic dependent launch, so the stage table is an explanation of the frame time, not a part of it
    lib.ts2d_profile_enable(1)
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize(dev)
    st_ms = (ctypes.c_float * len(_lib.STAGES))()
    st_n = (ctypes.c_int32 * len(_lib.STAGES))()
    lib.ts2d_profile_read(st_ms, st_n)
    lib.ts2d_profile_enable(0)
    stage_ms = {s: st_ms[i] / max(1, st_n[i]) for i, s in enumerate(_lib.STAGES)}

    # R (num_rendered) of this rank's shard: one forward through the raw _C API.  Every rank makes the call: a tile-sharded
    # forward is collective (barriers of the peer-memory fabric, or the NCCL assembly in the autograd wrapper)
    s, c = step.sc, step.sc.cam
    fa = (c["image_width"], c["image_height"], c["tanfovx"], c["tanfovy"], c["viewmatrix"], c["projmatrix"], c["campos"], s.sh_degree,
          s.gamma, 1.0, s.background_depth, s.background, s.vertex, s.shs, torch.Tensor([]), s.opacity, s.back_culling, s.rich_info, False)
    fwd_state = tsC.rasterize_triangles(*fa, shard=(rank, world) if world > 1 else (0, 1), primitive=a.primitive)
    R_local = int(fwd_state[0])
    R_total = R_local * world if R_local is not None else None  # interleaved tiles: shards are balanced to <1%
    work = work_statistics(fwd_state, sc, dev, world, stage_ms) if rank == 0 else None
    del fwd_state

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        K = (sc.sh_degree + 1) ** 2
        ab = algorithmic_bytes(sc.P, V, R_local, N // world, T // world, K, sc.shs.shape[1], sc.rich_info, rows=work.get("backward_rows", 0))
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        stages = {k: {"ms": round(stage_ms[k], 4), "alg_bytes": ab[k], "gbs": round(ab[k] / (stage_ms[k] * 1e-3) / 1e9, 1) if stage_ms[k] > 0 else None,
                      "share": round(stage_ms[k] / max(1e-9, sum(stage_ms.values())), 3)} for k in stage_ms}
        ach = ab[dom] / (stage_ms[dom] * 1e-3) / 1e9
        traffic, traffic_src, issue_pct = None, None, None  # DRAM bytes per launch / issue-slot use of the dominant kernel, from the committed ncu capture of this build
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if world == 1 and a.primitive == "2D" and cfg["workload"].startswith(tj["workload"]):
                traffic, traffic_src = tj["bytes_per_launch"].get(dom), tj["source"]
                issue_pct = tj.get("issue_active_pct", {}).get(dom)
        except (OSError, KeyError, ValueError):
            pass
        roof = {"bound": "hbm", "traffic_source": traffic_src, "kernel": {"render_fwd": "k_render_fwd_fast", "render_bwd": "k_render_bwd_fast", "preprocess": "k_preprocess",
                                            "preprocess_bwd": "k_preprocess_bwd", "binning": "k_emit_warp + k_tile_tables + k_radix_pass",
                                            "order_scan": "k_radix_hist + k_radix_pass + k_scan_sums / _block_sums / _apply", "bwd_prepare": "k_bwd_rows_mark + k_scan_sums / _block_sums / _apply",
                                            "bwd_reduce": "k_bwd_rows_reduce"}[dom],
                "achieved": round(ach, 2), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5), "traffic": traffic, "peak_source": peak_src,
                "issue_active_pct": issue_pct,
                "note": "per-pixel composite is FP32/MUFU-issue bound, not HBM bound (SURVEY.md section 8d): issue_active_pct (ncu, same capture as "
                        "`traffic`) is the fraction of its real ceiling; per-stage figures in `stages`",
                "whole_frame_gbs": round(sum(ab.values()) / (ms / a.steps * 1e-3) / 1e9, 1)}
        tile_bits = max(1, (T - 1).bit_length())
        # K1, depth histogram, 4 depth passes, 3 scan kernels | emit, tile tables (ranges + digit histograms), tile passes, K7, contrib
        # finish | row marking, 3 scan kernels, K8, row reduction, K9 -- every one of them a kernel of libts2d (no library kernels on
        # the path; memsets not counted)
        own_per_step = 9 + (3 + (tile_bits + 7) // 8 + (1 if sc.rich_info else 0)) + 7
        line = dict(base, impl="ours", value=fps, ms_per_step=ms / a.steps, per_step=per_step, config=cfg, clocks=clocks, roofline=roof, stages=stages,
                    gpu_launches=own_per_step * a.steps, work=work,
                    scene={"P": sc.P, "visible": V, "num_rendered_rank0": R_local, "num_rendered_est": R_total, "tiles": T, "pixels": N})

    # e2e (all ranks participate: the step contains collectives when sharded)
    if not a.no_e2e:
        host = pinned_host_buffers(sc, all_host=(world == 1))
        k_e = max(3, a.steps // 2)
        if world > 1:
            dist.barrier()
        ms_e = wall_region(lambda: step.e2e_step(host), k_e, 2, dev)
        if world > 1:
            t = torch.tensor([ms_e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        if rank == 0:
            h2d = sum(host[k].numel() * 4 for k in ("viewmatrix", "projmatrix", "campos", "background", "gt"))
            line["e2e"] = {"value": k_e / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                           "what": "camera H2D + GT image H2D (pinned, on a copy stream overlapping the forward) -> TriangleRasterizer fwd -> L1 loss -> backward -> loss.item()"}
        if world == 1:
            ms_a = wall_region(lambda: step.e2e_all_host_step(host), 3, 1, dev)
            h2d_all = sum(host[k].numel() * 4 for k in ("viewmatrix", "projmatrix", "campos", "background", "vertex", "shs", "opacity",
                                                         "g_feature", "g_depth", "g_normal"))
            d2h_all = sum(host[k].numel() * 4 for k in ("o_image", "o_gv", "o_gs", "o_go", "o_gc"))
            line["e2e_all_host"] = {"value": 3 / (ms_a / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                                    "what": "every parameter + upstream gradient H2D, image + all gradients D2H, per step (PCIe bound)"}

    # model-level step through the fused front-end (all ranks: sharded steps contain collectives)
    if sc.rich_info and sc.shs is not None and not a.no_model_step:
        mstep = ModelStepOurs(sc, dev, a.primitive)
        k_m = max(7, a.steps // 2)
        med, mean = median_step_ms(mstep, k_m, 5, dev, world)
        if rank == 0:
            line["model_step"] = {"value": 1e3 / med, "unit": "frames/s", "ms_per_step": med, "mean_ms_per_step": mean, "steps": k_m,
                                  "statistic": "median of per-step CUDA-event times", "what": MODEL_STEP_WHAT}
        del mstep

    if rank == 0 and world == 1 and not a.no_model_step:
        line["image_loss"] = image_loss_ms(sc, dev, fused=True)
        line["geometry_loss"] = geometry_loss_ms(sc, dev, fused=True)

    if world > 1 and not a.no_check:
        chk = sharded_vs_single_check(step, sc, dev, a.primitive, world)
        if rank == 0:
            line["check"] = chk

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_oracle_baseline(sc, primitive=a.primitive)
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # release the symmetric (peer-memory) buffers while the process group is still alive, then leave without running
        # interpreter-exit destructors: a rank that tears its CUDA context down early must not stall the others
        import gc

        from triangle_splatting_b200 import distributed as tsd

        del step
        tsd.disable_tile_sharding()
        gc.collect()
        torch.cuda.synchronize(dev)
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
