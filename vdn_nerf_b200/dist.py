"""Data-parallel plumbing for the path: one process per GPU, torch.distributed (NCCL on B200, gloo in CPU tests).

The reference is single-GPU; rays and grid blocks are independent, so the path shards with
no data-path collective:

* training / rendering: rank r takes a contiguous slice of the global ray batch; parameters are replicated;
  one all-reduce(sum) of a flat fp32 gradient buffer per step, plus one scalar all-reduce for the batch-global
  Eikonal normaliser of renderer.py:315 so the result equals the single-GPU gradient exactly;
* extract_fields: rank r owns a contiguous slab of x-planes of the grid; slabs are gathered on rank 0.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most one, earlier ranks larger."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rank: int, world_size: int, *tensors: torch.Tensor) -> List[torch.Tensor]:
    n = tensors[0].shape[0]
    lo, hi = shard_range(n, rank, world_size)
    return [t[lo:hi] for t in tensors]


def global_eikonal(eik_num: torch.Tensor, eik_den: torch.Tensor, group=None) -> torch.Tensor:
    """gradient_error with the batch-global normaliser: sum_r num_r is formed by the gradient all-reduce, the
    denominator sum_r den_r is all-reduced here (no gradient flows through it, as in the reference where the
    sphere mask is detached).  Returns this rank's additive share num_r / (den_global + 1e-5)."""
    den = eik_den.detach().sum().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(den, op=dist.ReduceOp.SUM, group=group)
    return eik_num.sum() / (den[0] + 1e-5)


def global_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of a (detached) scalar over all ranks; identity without an initialised process group."""
    t = x.detach().reshape(1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0]


class FlatGradAllReduce:
    """All-reduce(sum) of every parameter gradient through one flat fp32 buffer (1.41 M floats = 5.6 MB for
    the womsk_white networks): a single latency-bound collective per step, no bucketing."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None
        self._views = None

    def __call__(self, scale: float = 1.0):
        if not self.params:
            return
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
            self._views = None
        if self._views is None:
            off = 0
            self._views = []
            for p in self.params:
                n = p.numel()
                self._views.append(self.flat[off: off + n].view(p.shape))
                off += n
            if dev.type == "cuda":
                # let the weight-norm backward write straight into this buffer (ops.set_grad_arena): a gradient that
                # already lives in its slot needs neither the pack nor the unpack copy
                from . import ops
                ops.set_grad_arena(self.params, self._views)
        # pack: one multi-tensor copy (instead of one launch per parameter) for the gradients that are not in place
        inplace = [p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self._views)]
        missing = [v for p, v in zip(self.params, self._views) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        have = [(v, p.grad) for p, v, ip in zip(self.params, self._views, inplace) if p.grad is not None and not ip]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if scale != 1.0:
            self.flat.mul_(scale)
        # unpack: one multi-tensor copy back into the (existing or freshly created) .grad tensors that are not views
        out_g, out_v = [], []
        for p, v, ip in zip(self.params, self._views, inplace):
            if ip:
                continue
            if p.grad is None:
                p.grad = torch.empty_like(v)
            out_g.append(p.grad)
            out_v.append(v)
        if out_g:
            torch._foreach_copy_(out_g, out_v)
        if dev.type == "cuda":
            from . import ops
            ops.reset_grad_arena_use()


def gather_grid(u_local: torch.Tensor, resolution: int, group=None) -> Optional[torch.Tensor]:
    """Gather the per-rank x-slabs of extract_fields on rank 0 (returns None elsewhere)."""
    rank, ws = world()
    if ws == 1:
        return u_local
    sizes = [shard_range(resolution, r, ws) for r in range(ws)]
    if rank == 0:
        out = torch.empty(resolution, resolution, resolution, device=u_local.device, dtype=u_local.dtype)
        out[sizes[0][0]: sizes[0][1]] = u_local
        for r in range(1, ws):
            buf = out[sizes[r][0]: sizes[r][1]]
            dist.recv(buf, src=r, group=group)
        return out
    dist.send(u_local.contiguous(), dst=0, group=group)
    return None
