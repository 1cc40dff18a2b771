"""One optimiser-ready training step of the path, as ``dpt_runner.py:214-253`` forms it.

The reference driver stays the caller of ``NeuSRenderer.render``; this module restates only its loss
(dpt_runner.py:228-243) so that tests and ``bench.py`` can run the measured unit of work - render forward,
loss, backward, gradient all-reduce - without the driver's dataset / logging machinery.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import dist as vdist


def driver_loss(render_out, true_rgb, mask=None, igr_weight=0.1, mask_weight=0.0, gt_feats=None, depth_weight=1.0,
                global_batch: Optional[int] = None, data_parallel: bool = False):
    """color L1 / mask_sum + igr * eikonal + mask_weight * BCE [+ depth-feature L1 / mask_sum].

    With `data_parallel` the two batch-global normalisers (mask_sum and the Eikonal denominator) are formed
    over all ranks, so the sum over ranks of the returned losses - and hence the all-reduced gradient - equals
    the single-GPU value on the concatenated batch.
    """
    color = render_out["color_fine"]
    mask_given = mask is not None
    fused = color.is_cuda and not data_parallel
    need_mask = (not fused) or mask_weight != 0.0 or (gt_feats is not None and render_out.get("render_feats") is not None)
    mask_sum = None
    if need_mask:
        if mask is None:
            mask = torch.ones_like(render_out["weight_sum"])
        if data_parallel and global_batch is not None and not mask_given:
            mask_sum = float(global_batch) + 1e-5      # use_mask=False: mask == 1 (dpt_runner.py:209); host scalar
        elif data_parallel:
            mask_sum = vdist.global_sum(mask.sum()) + 1e-5     # use_mask=True: the normaliser is the global mask count
        elif mask_given:
            mask_sum = mask.sum() + 1e-5
        else:
            mask_sum = float(mask.numel()) + 1e-5      # mask == 1: its sum is the ray count
    if fused:
        # one kernel for the masked L1 sum, the mask count and the gradient sign (driver.color_loss, vdn_color_loss)
        from .driver import color_loss
        loss, _ = color_loss(color, true_rgb, mask if mask_given else None)
    else:
        color_error = (color - true_rgb) * mask
        loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    if data_parallel:
        eik = vdist.global_eikonal(render_out["_eik_num"], render_out["_eik_den"])
    else:
        eik = render_out["gradient_error"]
    loss = loss + eik * igr_weight
    if mask_weight != 0.0:
        bce = F.binary_cross_entropy(render_out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask,
                                     reduction="sum" if data_parallel else "mean")
        if data_parallel:
            bce = bce / float(global_batch if global_batch is not None else mask.numel())
        loss = loss + bce * mask_weight
    if gt_feats is not None and render_out.get("render_feats") is not None:
        err = (render_out["render_feats"] - gt_feats) * mask
        loss = loss + F.l1_loss(err, torch.zeros_like(err), reduction="sum") / mask_sum * depth_weight
    return loss


def train_step(renderer, params, rays_o, rays_d, near, far, true_rgb, gt_feats=None, background_rgb=None,
               cos_anneal_ratio=1.0, perturb_overwrite=-1, igr_weight=0.1, grad_sync=None, global_batch=None):
    """render + loss + backward (+ gradient all-reduce): the unit `train rays/s` counts.  Returns the loss.
    A barrier time-out of a tensor-core kernel in an earlier step raises here (asynchronous flag poll, no sync)."""
    from . import ops
    ops.poll_fault()
    for p in params:
        p.grad = None
    out = renderer.render(rays_o, rays_d, near, far, perturb_overwrite=perturb_overwrite,
                          background_rgb=background_rgb, cos_anneal_ratio=cos_anneal_ratio)
    loss = driver_loss(out, true_rgb, igr_weight=igr_weight, gt_feats=gt_feats, global_batch=global_batch,
                       data_parallel=grad_sync is not None)
    loss.backward()
    if grad_sync is not None:
        grad_sync()
    return loss.detach(), out


class GraphedTrainStep:
    """`train_step` captured once into a CUDA graph and replayed: the ~250 launches of a step (a good third of them
    few-microsecond PyTorch element-wise kernels of the sampler and the loss) are issued by one graph launch, which
    removes the gaps between them.  Shapes are fixed at capture.

    Single process only in practice: `grad_sync` / `global_batch` are accepted so that the two NCCL all-reduces of a
    data-parallel step could be captured as well, but that path is EXPERIMENTAL - the one 2-GPU trial of round 1 hung
    during warm-up / capture - and `bench.py` launches multi-GPU steps eagerly.

    The inputs of every call are copied into static device buffers; parameter tensors must keep their storage
    (in-place optimiser updates); `param.grad` tensors are allocated once, inside the graph's memory pool, and
    overwritten by each replay (re-attached if `optimizer.zero_grad(set_to_none=True)` dropped them).  The weight
    packing launches are part of the graph (`ops.force_repack`)."""

    def __init__(self, renderer, params, rays_o, rays_d, near, far, true_rgb, gt_feats=None, background_rgb=None,
                 cos_anneal_ratio=1.0, perturb_overwrite=-1, igr_weight=0.1, warmup=3, grad_sync=None, global_batch=None,
                 ray_grads=False):
        from . import ops
        self.params = list(params)
        self._static = [None if t is None else t.detach().clone()
                        for t in (rays_o, rays_d, near, far, true_rgb, gt_feats, background_rgb)]
        if ray_grads:       # learnable poses (BASELINE cfg 5): the step also back-propagates to the rays
            self._static[0].requires_grad_(True)
            self._static[1].requires_grad_(True)
        kw = dict(cos_anneal_ratio=cos_anneal_ratio, perturb_overwrite=perturb_overwrite, igr_weight=igr_weight,
                  grad_sync=grad_sync, global_batch=global_batch)

        def run():
            o, d, n, f, rgb, gt, bg = self._static
            if ray_grads:
                o.grad = d.grad = None
            return train_step(renderer, self.params, o, d, n, f, rgb, gt_feats=gt, background_rgb=bg, **kw)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        ops.force_repack(True)
        c0 = ops.launch_count()
        try:
            # with NCCL collectives inside, the capture must not poison the process-wide stream state: the process
            # group's watchdog thread queries events concurrently, which is illegal under the default "global" mode
            mode = "thread_local" if grad_sync is not None else "global"
            with torch.cuda.graph(self.graph, capture_error_mode=mode):
                self.loss, self.out = run()
        finally:
            ops.force_repack(False)
        self.launches_per_replay = ops.launch_count() - c0     # kernels of this library inside the graph
        self._grads = [p.grad for p in self.params]            # graph-pool tensors every replay writes into
        self.ray_grads = (self._static[0].grad, self._static[1].grad) if ray_grads else None

    def __call__(self, rays_o, rays_d, near, far, true_rgb, gt_feats=None, background_rgb=None):
        with torch.no_grad():
            for dst, src in zip(self._static, (rays_o, rays_d, near, far, true_rgb, gt_feats, background_rgb)):
                if dst is not None and src is not None and dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
        # an optimiser's zero_grad(set_to_none=True) (the default, as in dpt_runner.py:251) detaches the graph's gradient
        # tensors from the parameters: put them back, the replay writes into exactly these
        for p, g in zip(self.params, self._grads):
            if p.grad is not g:
                p.grad = g
        self.graph.replay()
        from . import ops
        ops.poll_fault()
        return self.loss, self.out
