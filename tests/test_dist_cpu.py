"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: ray sharding, flat gradient all-reduce, the
batch-global Eikonal normaliser and the slab gather of the grid query."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vdn_nerf_b200 import dist as vdist
from vdn_nerf_b200.training import driver_loss


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_render(theta, color_in, eik_in, relax):
    """A stand-in for render(): differentiable in `theta`, with the same dict entries driver_loss reads."""
    color = torch.sigmoid(color_in * theta[0] + theta[1])
    num = (relax * (eik_in * theta[2] - 1.0) ** 2).sum(-1)
    den = relax.sum(-1)
    return {"color_fine": color, "weight_sum": color.mean(-1, keepdim=True), "render_feats": None,
            "gradient_error": num.sum() / (den.sum() + 1e-5), "_eik_num": num, "_eik_den": den}


def _worker(rank, world_size, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(0)
        B = 10
        color_in, eik_in = torch.randn(B, 3, generator=g), torch.rand(B, 8, generator=g) * 2
        relax = (torch.rand(B, 8, generator=g) > 0.3).float()
        rgb = torch.full((B, 3), 0.5)
        theta = torch.nn.Parameter(torch.tensor([0.7, -0.2, 1.3]))
        # single-process reference on the full batch
        loss_full = driver_loss(_fake_render(theta, color_in, eik_in, relax), rgb)
        (g_full,) = torch.autograd.grad(loss_full, theta)
        # data-parallel: each rank renders its slice, normalisers are global, gradients are all-reduced
        ci, ei, rl, tg = vdist.shard_rays(rank, world_size, color_in, eik_in, relax, rgb)
        loss = driver_loss(_fake_render(theta, ci, ei, rl), tg, global_batch=B, data_parallel=True)
        loss.backward()
        sync = vdist.FlatGradAllReduce([theta])
        sync()
        total = loss.detach().clone()
        dist.all_reduce(total)
        ok = torch.allclose(theta.grad, g_full, rtol=1e-6, atol=1e-7) and torch.allclose(total, loss_full.detach(), rtol=1e-6)
        # slab gather of the grid query
        res = 7
        lo, hi = vdist.shard_range(res, rank, world_size)
        u_local = torch.arange(res ** 3, dtype=torch.float32).reshape(res, res, res)[lo:hi].clone()
        u = vdist.gather_grid(u_local, res)
        if rank == 0:
            ok = ok and torch.equal(u, torch.arange(res ** 3, dtype=torch.float32).reshape(res, res, res))
        else:
            ok = ok and u is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_equals_single_process():
    world_size = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world_size, _free_port(), ret), nprocs=world_size, join=True)
    assert all(ret.get(r) for r in range(world_size)), dict(ret)


def test_flat_grad_allreduce_single_process_paths():
    """Without a process group the reducer is a scaled identity; parameters that received no gradient get zeros, the
    flat buffer and its views are reused across calls."""
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2, 2))]
    ps[0].grad = torch.full((3, 4), 2.0)
    ps[2].grad = torch.arange(4.0).reshape(2, 2)
    sync = vdist.FlatGradAllReduce(ps)
    sync(scale=0.5)
    assert torch.equal(ps[0].grad, torch.full((3, 4), 1.0))
    assert torch.equal(ps[1].grad, torch.zeros(5))
    assert torch.equal(ps[2].grad, 0.5 * torch.arange(4.0).reshape(2, 2))
    flat = sync.flat
    ps[1].grad = torch.ones(5)
    sync()
    assert sync.flat is flat and torch.equal(ps[1].grad, torch.ones(5))
    assert torch.equal(sync.flat, torch.cat([p.grad.reshape(-1) for p in ps]))


def test_flat_grad_allreduce_leaves_in_place_gradients_alone():
    """A gradient that already lives in its slot of the flat buffer (the weight-norm backward wrote it there, ops.set_grad_arena)
    is neither packed nor unpacked: the collective runs on the buffer, `.grad` keeps aliasing it, the scale still applies."""
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    ps[0].grad = torch.full((3, 4), 2.0)
    ps[1].grad = torch.ones(5)
    sync = vdist.FlatGradAllReduce(ps)
    sync()
    v0 = sync._views[0]
    v0.fill_(3.0)                       # what the backward would do: write the new gradient into the slot ...
    ps[0].grad = v0.view_as(v0)         # ... and autograd adopts the alias
    ps[1].grad = torch.full((5,), 7.0)  # a gradient produced elsewhere still goes through the copies
    sync(scale=0.5)
    assert ps[0].grad.data_ptr() == v0.data_ptr()
    assert torch.equal(ps[0].grad, torch.full((3, 4), 1.5)) and torch.equal(ps[1].grad, torch.full((5,), 3.5))
    assert torch.equal(sync.flat, torch.cat([p.grad.reshape(-1) for p in ps]))
