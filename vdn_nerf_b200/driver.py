"""Host-side mirrors of the pieces of ``dpt_runner.py`` / ``dpt_models/poses.py`` that sit directly around the rendering
path ("next" rows), so a training loop can stay on the device end to end:

* ``FusedAdam`` + ``lr_factor`` / ``cos_anneal_ratio``  - torch.optim.Adam's update for all parameter tensors in ONE kernel
  launch and the driver's learning-rate / anneal schedules (dpt_runner.py:88, 251-253, 296-319);
* ``color_loss``                                         - the driver's masked L1 colour loss and PSNR (dpt_runner.py:228-232)
  as one kernel + its gradient;
* ``LearnPose`` / ``GpuRaysGenerator``                   - learnable so(3) pose refinement and random-ray generation on the
  device (poses.py:16-47, 189-212; lie_group_helper.py:47-81): the reference assembles rays on the CPU and copies them
  to the GPU every step;
* ``save_checkpoint`` / ``load_checkpoint``              - the reference's checkpoint dictionary layout (dpt_runner.py:350-381);
* ``render_image``                                       - ``validate_image``'s chunked render (dpt_runner.py:520-560), optionally
  sharded over ranks (third row).

The driver itself (dataset, logging, HOCON) stays out of scope; these are the calls it makes.
"""
from __future__ import annotations

import ctypes
import math
from typing import Iterable, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from . import dist as vdist
from ._lib import check, int_array, ptr_array


# ----------------------------------------------------------------------------------------------------------------
# N1: schedules, loss, optimiser
# ----------------------------------------------------------------------------------------------------------------
def lr_factor(iter_step: int, warm_up_end: float, end_iter: float, alpha: float) -> float:
    """dpt_runner.py:303-309: linear warm-up, then cosine decay to `alpha`."""
    if iter_step < warm_up_end:
        return iter_step / warm_up_end
    progress = (iter_step - warm_up_end) / (end_iter - warm_up_end)
    return (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha


def cos_anneal_ratio(iter_step: int, anneal_end: float) -> float:
    """dpt_runner.py:296-300."""
    return 1.0 if anneal_end == 0.0 else min(1.0, iter_step / anneal_end)


class _ColorLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, true_rgb, mask):
        color = ops._prep(color)
        true_rgb = ops._prep(true_rgb)
        mask = ops._prep(mask.reshape(-1)) if mask is not None else None
        B = color.shape[0]
        sums = torch.empty(3, device=color.device, dtype=torch.float32)
        d_color = torch.empty_like(color)
        check(_lib.load().vdn_color_loss(ops._p(color), ops._p(true_rgb), ops._p(mask), B, ops._p(sums), ops._p(d_color),
                                         ops._stream()), "vdn_color_loss")
        mask_sum = sums[2] + 1e-5
        ctx.save_for_backward(d_color, mask_sum)
        ctx.mark_non_differentiable(sums)
        return sums[0] / mask_sum, sums
    @staticmethod
    def backward(ctx, g_loss, _g_sums):
        d_color, mask_sum = ctx.saved_tensors
        return d_color * (g_loss / mask_sum), None, None


def color_loss(color: torch.Tensor, true_rgb: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """(color_fine_loss, psnr) of dpt_runner.py:228-232: L1(color_error, 0, 'sum') / mask_sum with mask_sum = mask.sum() +
    1e-5, psnr = 20 log10(1 / sqrt(sum(((color - true) mask)^2) / (mask_sum * 3)))."""
    loss, sums = _ColorLossFn.apply(color, true_rgb, mask)
    psnr = 20.0 * torch.log10(1.0 / torch.sqrt(sums[1] / ((sums[2] + 1e-5) * 3.0)))
    return loss, psnr


class FusedAdam:
    """torch.optim.Adam(params, lr, betas, eps) (no weight decay, no amsgrad) with the whole update of all tensors in ONE
    kernel launch (`vdn_adam_step`).  The hyper-parameters live in a small device tensor, written from pinned host memory
    before every step, so a CUDA graph that captured `step()` follows the learning-rate schedule.  Parameters without a
    gradient are skipped like torch does."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if len(self.params) > 96:
            raise ValueError("FusedAdam handles at most 96 parameter tensors per instance")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.VdnLibraryError("FusedAdam needs contiguous fp32 CUDA parameters")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self._m = torch.zeros(total, device=dev, dtype=torch.float32)      # flat optimiser state
        self._v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg, self.exp_avg_sq = [], []
        off = 0
        for p in self.params:
            n = p.numel()
            self.exp_avg.append(self._m[off: off + n].view_as(p))
            self.exp_avg_sq.append(self._v[off: off + n].view_as(p))
            off += n
        nt = len(self.params)
        self.steps = [0] * nt                                               # torch keeps one step counter per parameter
        # ring of pinned staging buffers: the host may run several steps ahead of the asynchronous copies
        self._hyper_ring = [torch.zeros(6 + 2 * nt, dtype=torch.float32).pin_memory() for _ in range(16)]
        self._ring_pos = 0
        self._hyper = torch.zeros(6 + 2 * nt, device=dev, dtype=torch.float32)
        self.param_groups = [{"params": self.params, "lr": self.lr}]        # the driver writes g['lr'] (dpt_runner.py:311)

    @property
    def step_count(self):
        return max(self.steps) if self.steps else 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def upload_hyper(self):
        """Write lr / betas / eps and the per-parameter bias corrections of the CURRENT step counters to the device."""
        b1, b2 = self.betas
        h = self._hyper_ring[self._ring_pos]
        self._ring_pos = (self._ring_pos + 1) % len(self._hyper_ring)
        h[0], h[1], h[2], h[3] = float(self.param_groups[0]["lr"]), b1, b2, self.eps
        h[4], h[5] = 1.0 - b1, 1.0 - b2
        for i, t in enumerate(self.steps):
            t = max(t, 1)
            h[6 + 2 * i], h[7 + 2 * i] = 1.0 - b1 ** t, 1.0 - b2 ** t
        self._hyper.copy_(h, non_blocking=True)

    def step(self):
        """One Adam update.  Under CUDA-graph capture the step counters are not advanced and the hyper-parameters are not
        re-uploaded: advance `steps` and call `upload_hyper()` before each replay."""
        have = [p.grad is not None for p in self.params]
        if not any(have):
            return
        if not torch.cuda.is_current_stream_capturing():
            self.steps = [t + 1 if h else t for t, h in zip(self.steps, have)]
            self.upload_hyper()
        for p, h in zip(self.params, have):
            if h and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous()):
                p.grad = p.grad.float().contiguous()
        check(_lib.load().vdn_adam_step(len(self.params), ptr_array([p.data_ptr() for p in self.params]),
                                        ptr_array([p.grad.data_ptr() if h else 0 for p, h in zip(self.params, have)]),
                                        ptr_array([m.data_ptr() for m in self.exp_avg]),
                                        ptr_array([v.data_ptr() for v in self.exp_avg_sq]),
                                        int_array([p.numel() if h else 0 for p, h in zip(self.params, have)]),
                                        ops._p(self._hyper), ops._stream()), "vdn_adam_step")
        ops.bump_param_epoch()      # parameters changed behind autograd's version counters: packed weights are stale

    # torch-compatible state (dpt_runner.py:356, 368: the checkpoint stores optimizer.state_dict())
    def state_dict(self):
        return {"state": {i: {"step": torch.tensor(float(t)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
                          for i, (t, m, v) in enumerate(zip(self.steps, self.exp_avg, self.exp_avg_sq)) if t > 0},
                "param_groups": [{"lr": float(self.param_groups[0]["lr"]), "betas": self.betas, "eps": self.eps,
                                  "weight_decay": 0, "amsgrad": False, "params": list(range(len(self.params)))}]}

    def load_state_dict(self, sd):
        st = sd["state"]
        for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq)):
            if i in st:
                m.copy_(st[i]["exp_avg"])
                v.copy_(st[i]["exp_avg_sq"])
                self.steps[i] = int(float(st[i]["step"]))
        if sd.get("param_groups"):
            self.param_groups[0]["lr"] = float(sd["param_groups"][0]["lr"])


# ----------------------------------------------------------------------------------------------------------------
# N2: learnable poses and ray generation on the device
# ----------------------------------------------------------------------------------------------------------------
def so3_exp(r: torch.Tensor) -> torch.Tensor:
    """lie_group_helper.py:60-70: axis-angle (3,) -> rotation matrix (3,3) (Rodrigues, with the reference's 1e-15)."""
    zero = torch.zeros(1, dtype=torch.float32, device=r.device)
    skew = torch.stack([torch.cat([zero, -r[2:3], r[1:2]]), torch.cat([r[2:3], zero, -r[0:1]]),
                        torch.cat([-r[1:2], r[0:1], zero])], dim=0)
    norm_r = r.norm() + 1e-15
    eye = torch.eye(3, dtype=torch.float32, device=r.device)
    return eye + (torch.sin(norm_r) / norm_r) * skew + ((1 - torch.cos(norm_r)) / norm_r ** 2) * (skew @ skew)


class LearnPose(nn.Module):
    """poses.py:16-47: per-camera delta pose (axis-angle r, translation t) applied on the left of the initial pose."""

    def __init__(self, num_cams, learn_R=True, learn_t=True, init_c2w=None):
        super().__init__()
        self.num_cams = num_cams
        self.init_c2w = nn.Parameter(init_c2w.clone().float(), requires_grad=False) if init_c2w is not None else None
        self.r = nn.Parameter(torch.zeros(num_cams, 3, dtype=torch.float32), requires_grad=learn_R)
        self.t = nn.Parameter(torch.zeros(num_cams, 3, dtype=torch.float32), requires_grad=learn_t)

    def forward(self, cam_id):
        R = so3_exp(self.r[cam_id])
        c2w = torch.cat([torch.cat([R, self.t[cam_id].unsqueeze(1)], dim=1),
                         torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=R.device)], dim=0)
        if self.init_c2w is not None:
            c2w = c2w @ self.init_c2w[cam_id]
        return c2w


class _RayGenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, px, py, kinv, pose34):
        px, py = ops._prep(px), ops._prep(py)
        kinv, pose34 = ops._prep(kinv), ops._prep(pose34)
        B = px.shape[0]
        rays_o = torch.empty(B, 3, device=px.device, dtype=torch.float32)
        rays_d = torch.empty(B, 3, device=px.device, dtype=torch.float32)
        check(_lib.load().vdn_raygen_fwd(ops._p(px), ops._p(py), B, ops._p(kinv), ops._p(pose34), ops._p(rays_o),
                                         ops._p(rays_d), ops._stream()), "vdn_raygen_fwd")
        ctx.save_for_backward(px, py, kinv)
        return rays_o, rays_d

    @staticmethod
    def backward(ctx, d_o, d_d):
        px, py, kinv = ctx.saved_tensors
        d_o = ops._prep(d_o) if d_o is not None else None
        d_d = ops._prep(d_d) if d_d is not None else None
        d_pose = torch.empty(3, 4, device=px.device, dtype=torch.float32)
        check(_lib.load().vdn_raygen_bwd(ops._p(px), ops._p(py), px.shape[0], ops._p(kinv), ops._p(d_o), ops._p(d_d),
                                         ops._p(d_pose), ops._stream()), "vdn_raygen_bwd")
        return None, None, None, d_pose


def rays_from_pixels(px, py, intrinsic_inv, c2w):
    """(rays_o, rays_d) [B,3] of pixels (px, py) for a camera with inverse intrinsics [>=3,>=3] and pose c2w [>=3,4]
    (poses.py:199-207); differentiable with respect to the pose."""
    return _RayGenFn.apply(px.float(), py.float(), intrinsic_inv[:3, :3].contiguous(), c2w[:3, :4].contiguous())


def near_far_from_sphere(rays_o, rays_d):
    """dataset.py:111-118."""
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return mid - 1.0, mid + 1.0


class GpuRaysGenerator:
    """`RaysGenerator.gen_random_rays_at` (poses.py:189-212) with images, masks, pixel draws and the ray transform on the
    device: returns (rays_o, rays_d, mask[:, :1], color) without the reference's CPU gather and host round trip."""

    def __init__(self, images, masks, intrinsics, pose_net, learnable=True):
        self.images, self.masks = images, masks            # [n, H, W, 3] device tensors
        self.H, self.W = images.shape[1], images.shape[2]
        self.intrinsics_inv = torch.inverse(intrinsics)     # [n, 4, 4] or [4, 4]
        self.pose_net = pose_net                            # LearnPose, or a [n, 4, 4] tensor of fixed poses
        self.learnable = learnable

    def pose(self, img_idx):
        return self.pose_net(img_idx) if self.learnable else self.pose_net[img_idx]

    def gen_random_rays_at(self, img_idx, batch_size, generator=None):
        dev = self.images.device
        px = torch.randint(0, self.W, [batch_size], device=dev, generator=generator)
        py = torch.randint(0, self.H, [batch_size], device=dev, generator=generator)
        color = self.images[img_idx][(py, px)]
        mask = self.masks[img_idx][(py, px)]
        kinv = self.intrinsics_inv if self.intrinsics_inv.dim() == 2 else self.intrinsics_inv[img_idx]
        rays_o, rays_d = rays_from_pixels(px, py, kinv, self.pose(img_idx))
        return rays_o, rays_d, mask[:, :1], color


# ----------------------------------------------------------------------------------------------------------------
# N3: checkpoints with the reference's layout
# ----------------------------------------------------------------------------------------------------------------
def checkpoint_dict(nerf, sdf_network, deviation_network, color_network, depth_network, optimizer, iter_step):
    """dpt_runner.py:369-378."""
    return {"nerf": nerf.state_dict(), "sdf_network_fine": sdf_network.state_dict(),
            "variance_network_fine": deviation_network.state_dict(), "color_network_fine": color_network.state_dict(),
            "depth_network_fine": depth_network.state_dict() if depth_network is not None else None,
            "optimizer": optimizer.state_dict() if optimizer is not None else None, "iter_step": int(iter_step)}


def save_checkpoint(path, nerf, sdf_network, deviation_network, color_network, depth_network, optimizer, iter_step):
    torch.save(checkpoint_dict(nerf, sdf_network, deviation_network, color_network, depth_network, optimizer, iter_step), path)


def load_checkpoint(checkpoint, nerf, sdf_network, deviation_network, color_network, depth_network=None, optimizer=None,
                    map_location=None):
    """dpt_runner.py:350-361; `checkpoint` is a path or an already loaded dict.  Returns iter_step."""
    if not isinstance(checkpoint, dict):
        checkpoint = torch.load(checkpoint, map_location=map_location)
    nerf.load_state_dict(checkpoint["nerf"], strict=False)
    sdf_network.load_state_dict(checkpoint["sdf_network_fine"])
    deviation_network.load_state_dict(checkpoint["variance_network_fine"])
    color_network.load_state_dict(checkpoint["color_network_fine"])
    if depth_network is not None and checkpoint.get("depth_network_fine") is not None:
        depth_network.load_state_dict(checkpoint["depth_network_fine"])
    if optimizer is not None and checkpoint.get("optimizer") is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
    return checkpoint["iter_step"]


# ----------------------------------------------------------------------------------------------------------------
# e3: validation-image rendering, sharded over ranks
# ----------------------------------------------------------------------------------------------------------------
def render_image(renderer, rays_o, rays_d, batch_size, background_rgb=None, cos_anneal_ratio=1.0,
                 depth_before_color=False, group=None, gather=True):
    """`validate_image`'s loop (dpt_runner.py:528-558): rays [H, W, 3] are rendered in chunks of `batch_size` rays, colour
    and weight-averaged normals are collected.  With an initialised process group the CHUNKS are dealt round-robin to
    the ranks (chunk c -> rank c % world) and, when `gather`, assembled on every rank.  Returns (rgb [H,W,3],
    normals [H,W,3]) numpy arrays (None on ranks that do not gather)."""
    H, W, _ = rays_o.shape
    rank, world = vdist.world()
    o_chunks = rays_o.reshape(-1, 3).split(batch_size)
    d_chunks = rays_d.reshape(-1, 3).split(batch_size)
    n_s = renderer.n_samples + renderer.n_importance
    rgb = torch.zeros(H * W, 3, device=rays_o.device)
    nrm = torch.zeros(H * W, 3, device=rays_o.device)
    off = 0
    for c, (ob, db) in enumerate(zip(o_chunks, d_chunks)):
        n = ob.shape[0]
        if c % world == rank:
            near, far = near_far_from_sphere(ob, db)
            out = renderer.render(ob, db, near, far, cos_anneal_ratio=cos_anneal_ratio, background_rgb=background_rgb,
                                  depth_before_color=depth_before_color)
            rgb[off: off + n] = out["color_fine"].detach()
            normals = out["gradients"] * out["weights"][:, :n_s, None]
            if out.get("inside_sphere") is not None:
                normals = normals * out["inside_sphere"][..., None]
            nrm[off: off + n] = normals.sum(dim=1).detach()
            del out
        off += n
    if world > 1 and gather:
        import torch.distributed as dist
        dist.all_reduce(rgb, op=dist.ReduceOp.SUM, group=group)     # chunks are disjoint: the sum assembles the image
        dist.all_reduce(nrm, op=dist.ReduceOp.SUM, group=group)
    return rgb.reshape(H, W, 3).cpu().numpy(), nrm.reshape(H, W, 3).cpu().numpy()
