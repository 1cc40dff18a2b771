"""Shared helpers for the test-suite (fixtures, network construction, error metrics)."""
import os

import numpy as np
import torch

from oracle import vdn_oracle as vo
from vdn_nerf_b200 import configs, fields

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = {"womsk_white": "womsk_white_b16.npz", "womsk_white_wdepth": "womsk_white_wdepth_b8.npz"}


def load_fixture(name):
    return np.load(os.path.join(GOLDEN, FIXTURES[name]))


def digest(t: torch.Tensor) -> np.ndarray:
    a = t.detach().cpu().double().reshape(-1).numpy()
    idx = (np.arange(16, dtype=np.int64) * 2654435761) % a.size
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def build(name, device=None, seed=0):
    """(modules tuple, conf) of this package's classes, seeded like the golden script."""
    conf = configs.CONFIGS[name]
    return configs.build_networks(conf, fields, seed=seed, device=device), conf


def oracle_nets(mods, conf, dtype=None, requires_grad=False):
    nets = vo.nets_from_modules(*mods, conf, dtype=dtype)
    if requires_grad:
        for _, t in nets.leaves():
            t.requires_grad_(True)
    return nets


def relerr(got, want):
    """max |got - want| / max |want| (inf-norm relative error)."""
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    den = want.abs().max().item()
    return (got - want).abs().max().item() / (den if den > 0 else 1.0)


def relerr_l2(got, want):
    """||got - want||_2 / ||want||_2."""
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    den = want.norm().item()
    return (got - want).norm().item() / (den if den > 0 else 1.0)


def t(a, device=None):
    return torch.from_numpy(np.asarray(a)).to(device) if device else torch.from_numpy(np.asarray(a))


def module_param_map(mods):
    """{'sdf.lin0.weight_v': param, ...} with the oracle's leaf naming."""
    nerf, sdf, var, col, dep = mods
    out = {}
    for tag, m in (("nerf", nerf), ("sdf", sdf), ("color", col), ("depth", dep)):
        if m is not None:
            out.update({f"{tag}.{k}": p for k, p in m.named_parameters()})
    out["variance"] = var.variance
    return out
