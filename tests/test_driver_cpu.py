"""CPU tests of the host-side driver mirrors (vdn_nerf_b200/driver.py): schedules, so(3) pose, checkpoint layout and the
rank-sharded validation-image render (world_size-2 gloo, stand-in renderer)."""
import math
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vdn_nerf_b200 import driver


def test_schedules_match_the_driver_formulas():
    # dpt_runner.py:303-309 / 296-300 with the shipped conf values
    lr, alpha, warm, end = 5e-4, 0.05, 5000, 300000
    assert driver.lr_factor(0, warm, end, alpha) == 0.0
    assert abs(driver.lr_factor(2500, warm, end, alpha) - 0.5) < 1e-12
    assert abs(driver.lr_factor(warm, warm, end, alpha) - 1.0) < 1e-12
    assert abs(driver.lr_factor(end, warm, end, alpha) - alpha) < 1e-12
    mid = (warm + end) // 2
    want = (np.cos(np.pi * (mid - warm) / (end - warm)) + 1.0) * 0.5 * (1 - alpha) + alpha
    assert abs(driver.lr_factor(mid, warm, end, alpha) - want) < 1e-12
    assert driver.cos_anneal_ratio(100, 0.0) == 1.0 and driver.cos_anneal_ratio(25000, 50000) == 0.5
    assert driver.cos_anneal_ratio(10 ** 6, 50000) == 1.0


def test_so3_exp_is_a_rotation_and_matches_rodrigues():
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        r = torch.randn(3, generator=g)
        R = driver.so3_exp(r)
        assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5) and abs(float(torch.det(R)) - 1.0) < 1e-5
        th = float(r.norm())
        k = (r / th).double()
        K = torch.tensor([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=torch.float64)
        want = torch.eye(3, dtype=torch.float64) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)
        assert torch.allclose(R.double(), want, atol=1e-5)
    assert torch.allclose(driver.so3_exp(torch.zeros(3)), torch.eye(3))      # the reference's 1e-15 keeps r = 0 finite


def test_learn_pose_is_delta_times_init_and_differentiable():
    g = torch.Generator().manual_seed(1)
    init = torch.eye(4).repeat(3, 1, 1)
    init[:, :3, 3] = torch.randn(3, 3, generator=g)
    net = driver.LearnPose(3, True, True, init)
    with torch.no_grad():
        net.r[1] = torch.tensor([0.1, -0.2, 0.05])
        net.t[1] = torch.tensor([0.3, 0.0, -0.1])
    c2w = net(1)
    want = torch.eye(4)
    want[:3, :3] = driver.so3_exp(net.r[1])
    want[:3, 3] = net.t[1]
    assert torch.allclose(c2w, want @ init[1], atol=1e-6)
    c2w.sum().backward()
    assert net.r.grad is not None and net.t.grad is not None and net.init_c2w.grad is None


def test_checkpoint_dict_has_the_reference_layout():
    from vdn_nerf_b200 import configs
    import types
    import torch.nn as nn

    class Stub(nn.Module):          # key layout only: the real modules need the CUDA library for nothing here either
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.zeros(2))
    mods = [Stub() for _ in range(4)]
    ck = driver.checkpoint_dict(mods[0], mods[1], mods[2], mods[3], None, None, 123)
    assert list(ck.keys()) == ["nerf", "sdf_network_fine", "variance_network_fine", "color_network_fine",
                               "depth_network_fine", "optimizer", "iter_step"]     # dpt_runner.py:369-378
    assert ck["depth_network_fine"] is None and ck["iter_step"] == 123
    with torch.no_grad():
        mods[1].w.fill_(3.0)
    ck = driver.checkpoint_dict(*mods, None, None, 7)
    fresh = [Stub() for _ in range(4)]
    assert driver.load_checkpoint(ck, *fresh) == 7
    assert torch.equal(fresh[1].w, mods[1].w)


class _FakeRenderer:
    n_samples, n_importance = 2, 2

    def render(self, o, d, near, far, **kw):
        B = o.shape[0]
        col = torch.sigmoid(o + d)
        g = torch.stack([d, d * 2, d * 3, d * 4], dim=1)
        w = torch.softmax(torch.cat([near, far, near + far, far - near, near * 0], dim=1), dim=1)
        return {"color_fine": col, "gradients": g, "weights": w, "inside_sphere": torch.ones(B, 4)}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(3)
        H, W = 5, 7
        o = torch.randn(H, W, 3, generator=g)
        d = torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g), dim=-1)
        rgb, nrm = driver.render_image(_FakeRenderer(), o, d, batch_size=8)
        ret[rank] = (rgb, nrm)
    finally:
        dist.destroy_process_group()


def test_render_image_sharded_over_two_ranks_equals_single_process():
    g = torch.Generator().manual_seed(3)
    H, W = 5, 7
    o = torch.randn(H, W, 3, generator=g)
    d = torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g), dim=-1)
    want_rgb, want_nrm = driver.render_image(_FakeRenderer(), o, d, batch_size=8)
    assert want_rgb.shape == (H, W, 3)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        assert np.allclose(ret[r][0], want_rgb, atol=1e-6) and np.allclose(ret[r][1], want_nrm, atol=1e-6)
