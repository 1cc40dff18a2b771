"""Run the UNMODIFIED reference (staged under oracle/_ref by oracle/stage_ref.py) on the benchmark workload.

BENCH / TEST INFRASTRUCTURE ONLY - nothing under vdn_nerf_b200/ imports this.  Two uses (bench.py):

  * `bench.py --impl reference`: the reference's own CPU implementation of the training step (PyTorch CPU autograd through
    NeuSRenderer.render, dpt_models/renderer.py:332-439, and the driver's loss, dpt_runner.py:228-243) on the box's host
    cores, same rays / weights / config as the CUDA arm;
  * the `reference_cuda_eager` leg: the same classes on the GPU through the reference's own switch
    `torch.set_default_tensor_type('torch.cuda.FloatTensor')` (dpt_runner.py:744) - the honest "reference on the same
    box" number.  It changes a process-wide default, so it runs in a subprocess:

        python -m oracle.ref_runner --device cuda --rays 512 --steps 10 --warmup 3 [--depth]

    prints one JSON line {"rays_per_s", "ms_per_step", "launch_bound_note", ...}.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import stage_ref  # noqa: E402
from oracle import vdn_oracle as vo  # noqa: E402


def build(conf, device="cpu"):
    """Reference networks (driver's construction order, seed 0, built on the CPU then moved) + NeuSRenderer."""
    from vdn_nerf_b200 import configs
    rf, rr, _ = stage_ref.import_reference()
    mods = configs.build_networks(conf, rf, seed=0, device=None)
    mods = tuple(m.to(device) if m is not None else None for m in mods)
    rend = rr.NeuSRenderer(*mods, **conf["neus_renderer"])
    params = [p for m in mods if m is not None for p in m.parameters()]
    return mods, rend, params


def driver_loss(out, true_rgb, gt_feats=None, igr_weight=0.1):
    """dpt_runner.py:228-243 with mask == 1 (use_mask False), mask_weight 0, depth weight 1."""
    mask_sum = float(true_rgb.shape[0]) + 1e-5
    err = out["color_fine"] - true_rgb
    loss = F.l1_loss(err, torch.zeros_like(err), reduction="sum") / mask_sum
    loss = loss + out["gradient_error"] * igr_weight
    if gt_feats is not None and out.get("render_feats") is not None:
        e2 = out["render_feats"] - gt_feats
        loss = loss + F.l1_loss(e2, torch.zeros_like(e2), reduction="sum") / mask_sum
    return loss


def step(rend, params, o, d, near, far, rgb, bg, gt=None):
    for p in params:
        p.grad = None
    out = rend.render(o, d, near, far, background_rgb=bg, cos_anneal_ratio=1.0)
    loss = driver_loss(out, rgb, gt)
    loss.backward()
    return loss


def run(device: str, rays: int, steps: int, warmup: int, depth: bool, threads: int = 0):
    from vdn_nerf_b200 import configs
    conf = configs.CONFIGS["womsk_white_wdepth" if depth else "womsk_white"]
    if device == "cuda":
        torch.set_default_tensor_type("torch.cuda.FloatTensor")     # the reference's own device switch
    else:
        torch.set_num_threads(threads or (os.cpu_count() or 1))
    # build on the CPU default so the initial parameters equal the CUDA arm's bit for bit
    if device == "cuda":
        torch.set_default_tensor_type("torch.FloatTensor")
    mods, rend, params = build(conf, "cpu")
    rays_cpu = vo.synthetic_rays(rays)               # seeded CPU generator: the same rays as the CUDA arm of bench.py
    if device == "cuda":
        torch.set_default_tensor_type("torch.cuda.FloatTensor")
        mods = tuple(m.cuda() if m is not None else None for m in mods)
        rf, rr, _ = stage_ref.import_reference()
        rend = rr.NeuSRenderer(*mods, **conf["neus_renderer"])
        params = [p for m in mods if m is not None for p in m.parameters()]
    o, d, near, far = (t.to(device) for t in rays_cpu)
    rgb = torch.full((rays, 3), 0.5, device=device)
    gt = torch.full((rays, 96), 0.5, device=device) if depth else None
    bg = torch.ones(1, 3, device=device)
    times = []
    for i in range(warmup + steps):
        torch.manual_seed(2 + i)
        if device == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = step(rend, params, o, d, near, far, rgb, bg, gt)
        float(loss)
        if device == "cuda":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return {"device": device, "rays_per_step": rays, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * mean,
            "rays_per_s": rays / mean, "threads": torch.get_num_threads(), "depth": depth,
            "times_ms": [round(1e3 * t, 3) for t in times]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--rays", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--depth", action="store_true")
    a = ap.parse_args()
    import warnings
    warnings.simplefilter("ignore")
    print(json.dumps(run(a.device, a.rays, a.steps, a.warmup, a.depth)))


if __name__ == "__main__":
    main()
