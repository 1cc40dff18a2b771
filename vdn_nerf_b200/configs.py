"""Hyper-parameters of the reference's shipped configurations (the values are the contract).

``confs/womsk_white.conf`` (lines 41-91) and ``confs/womsk_white_wdepth.conf`` (lines 47-72) of the reference,
as plain dicts with the exact constructor kwargs ``dpt_runner.py:117-142`` splats into the classes.
"""
from __future__ import annotations

import copy

import torch

WOMSK_WHITE = {
    "nerf": dict(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4, skips=[4], rgb_dims=3,
                 use_viewdirs=True),
    "sdf_network": dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True),
    "variance_network": dict(init_val=0.3),
    "rendering_network": dict(d_feature=256, mode="idr", d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                              multires_view=4, squeeze_out=True),
    "depth_extract_network": None,
    "neus_renderer": dict(n_samples=64, n_importance=64, n_outside=32, up_sample_steps=4, perturb=1.0),
    "train": dict(batch_size=512, igr_weight=0.1, mask_weight=0.0, use_white_bkgd=True, anneal_end=50000),
}

WOMSK_WHITE_WDEPTH = copy.deepcopy(WOMSK_WHITE)
WOMSK_WHITE_WDEPTH["nerf"].update(gen_depth_feats=True, dpt_dim=96)
WOMSK_WHITE_WDEPTH["depth_extract_network"] = dict(d_feature=256, mode="idr", d_in=9, d_out=96, d_hidden=256,
                                                   n_layers=4, weight_norm=True, multires_view=4, squeeze_out=True)

CONFIGS = {"womsk_white": WOMSK_WHITE, "womsk_white_wdepth": WOMSK_WHITE_WDEPTH}


def build_networks(conf, classes, seed=0, device=None):
    """Construct (nerf, sdf, variance, colour, depth|None) in the driver's order (dpt_runner.py:117-129) after
    ``torch.manual_seed(seed)`` on the CPU, then move them to `device`.  `classes` is any namespace providing
    NeRF / SDFNetwork / SingleVarianceNetwork / RenderingNetwork (this package's or the reference's)."""
    if seed is not None:
        torch.manual_seed(seed)
    nerf = classes.NeRF(**conf["nerf"])
    sdf = classes.SDFNetwork(**conf["sdf_network"])
    var = classes.SingleVarianceNetwork(**conf["variance_network"])
    col = classes.RenderingNetwork(**conf["rendering_network"])
    dep = classes.RenderingNetwork(**conf["depth_extract_network"]) if conf["depth_extract_network"] else None
    nets = [nerf, sdf, var, col, dep]
    if device is not None:
        nets = [n.to(device) if n is not None else None for n in nets]
    return tuple(nets)
