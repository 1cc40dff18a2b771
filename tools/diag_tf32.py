"""Diagnostics: tf32 vs fp32 mode of the colour head (forward + backward) for several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vdn_nerf_b200 import configs, fields, ops

dev = "cuda"
conf = configs.CONFIGS["womsk_white"]
mods = configs.build_networks(conf, fields, seed=0, device=dev)
col = mods[3]


def run(n, mode, scale):
    ops.set_precision(mode)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
    v = torch.randn(n, 3, generator=g)
    v = (v / v.norm(dim=-1, keepdim=True)).to(dev)
    nrm = torch.randn(n, 3, generator=g).to(dev)
    feat = (torch.randn(n, 256, generator=g) * scale).to(dev).requires_grad_(True)
    cot = torch.randn(n, 3, generator=g).to(dev)
    col.zero_grad()
    out = col(x, nrm, v, feat)
    (out * cot).sum().backward()
    return out.detach().clone(), {k: p.grad.clone() for k, p in col.named_parameters()}, feat.grad.clone()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


for scale in (1.0, 0.05):
    for n in (256, 700, 1024):
        o32, g32, f32 = run(n, "fp32", scale)
        o19, g19, f19 = run(n, "tf32", scale)
        worst = max((rel(g19[k], g32[k]), k) for k in g32)
        print(f"scale={scale} n={n}: out {rel(o19, o32):.2e}  dfeat {rel(f19, f32):.2e}  worst grad {worst[0]:.2e} ({worst[1]})")
        print("    per-layer bias grads:", " ".join(f"{rel(g19[f'lin{l}.bias'], g32[f'lin{l}.bias']):.1e}" for l in range(5)))
print("fault", ops.tc_fault())
