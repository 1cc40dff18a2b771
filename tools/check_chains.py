"""GPU diagnostic: fused training chains (chain engine + grouped TMA weight gradient) against the layer-wise tensor-core
path and the exact-fp32 path of the same library, network by network, tensor by tensor.

    python tools/check_chains.py [sdf] [color] [nerf] [--n 1000]

Prints inf-norm / L2 relative errors of outputs and of every parameter gradient; exits non-zero when a tensor is off by
more than the tensor-core tolerances of tests/test_gpu_tf32.py or the fault flag of a tcgen05 kernel was raised.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import util  # noqa: E402
from vdn_nerf_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)


def run_sdf(mod, x, cs, cf, cn, want_dx):
    mod.zero_grad()
    xx = x.clone().requires_grad_(want_dx)
    sdf, feat, nrm = mod.forward_split(xx)
    loss = (sdf * cs).sum() + (feat * cf).sum() + (nrm * cn).sum()
    loss.backward()
    out = {"sdf": sdf.detach(), "feat": feat.detach(), "normals": nrm.detach()}
    out.update({"d." + k: p.grad.detach().clone() for k, p in mod.named_parameters()})
    if want_dx:
        out["d.x"] = xx.grad.detach().clone()
    return out


def run_color(mod, pts, nrm, dirs, feat, co):
    mod.zero_grad()
    ins = [t.clone().requires_grad_(True) for t in (pts, nrm, dirs, feat)]
    out = mod(*ins)
    (out * co).sum().backward()
    res = {"out": out.detach()}
    res.update({"d." + k: p.grad.detach().clone() for k, p in mod.named_parameters()})
    for nm, t in zip(("pts", "normals", "dirs", "feat"), ins):
        res["d." + nm] = t.grad.detach().clone()
    return res


def run_nerf(mod, p4, dirs, cs, cr, cd, want_dx):
    mod.zero_grad()
    pp = p4.clone().requires_grad_(want_dx)
    dd = dirs.clone().requires_grad_(want_dx)
    s, r, d = mod(pp, dd)
    loss = (s * cs).sum() + (r * cr).sum()
    if d is not None:
        loss = loss + (d * cd).sum()
    loss.backward()
    res = {"sigma": s.detach(), "rgb": r.detach()}
    if d is not None:
        res["dpt"] = d.detach()
    res.update({"d." + k: p.grad.detach().clone() for k, p in mod.named_parameters()})
    if want_dx:
        res["d.pts"] = pp.grad.detach().clone()
        res["d.dirs"] = dd.grad.detach().clone()
    return res


def compare(title, ref, got, tol_out, tol_grad, l2_grad):
    bad = 0
    print(f"--- {title}")
    for k, w in ref.items():
        g = got[k]
        e, e2 = util.relerr(g, w), util.relerr_l2(g, w)
        is_grad = k.startswith("d.")
        lim = tol_grad if is_grad else tol_out
        ok = (e2 < lim) if (is_grad and l2_grad) else (e < lim)
        if not torch.isfinite(g).all():
            ok = False
        flag = "" if ok else "   <<<<<< FAIL"
        print(f"  {k:34s} inf {e:9.2e}  l2 {e2:9.2e}{flag}")
        bad += 0 if ok else 1
    return bad


def three_way(fn, tol_out=2e-3, tol_grad=2e-2, l2_grad=False, name=""):
    ops.set_precision("fp32")
    ref = fn()
    ops.set_precision("tf32")
    ops.set_chain(False)
    lw = fn()
    ops.set_chain(True)
    ch = fn()
    torch.cuda.synchronize()
    bad = compare(name + ": layer-wise tensor-core vs fp32", ref, lw, tol_out, tol_grad, l2_grad)
    bad_c = compare(name + ": fused chains vs fp32", ref, ch, tol_out, tol_grad, l2_grad)
    f = ops.tc_fault()
    if f:
        print("  TC FAULT FLAG RAISED")
    return bad_c + (1 if f else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["sdf", "color", "nerf"])
    ap.add_argument("--n", type=int, nargs="*", default=[1000])
    ap.add_argument("--conf", default="womsk_white_wdepth")
    ap.add_argument("--no-dx", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    mods, conf = util.build(args.conf, device=DEV)
    nerf, sdf, var, col, dep = mods
    bad = 0
    for n in args.n:
        g = torch.Generator().manual_seed(n)
        x = (torch.rand(n, 3, generator=g) * 2.4 - 1.2).to(DEV)
        r = lambda *s: torch.randn(*s, generator=g).to(DEV)
        if "sdf" in args.which:
            cs, cf, cn = r(n, 1), r(n, 256) * 0.1, r(n, 3)
            for want_dx in ([False] if args.no_dx else [False, True]):
                bad += three_way(lambda: run_sdf(sdf, x, cs, cf, cn, want_dx), name=f"SDF n={n} dx={want_dx}")
        if "color" in args.which:
            nrm, dirs, feat = r(n, 3), torch.nn.functional.normalize(r(n, 3), dim=-1), r(n, 256) * 0.3
            for tag, m in (("colour", col), ("depth", dep)):
                if m is None:
                    continue
                co = r(n, m.d_out)
                bad += three_way(lambda: run_color(m, x, nrm, dirs, feat, co), tol_grad=1e-1, l2_grad=True,
                                 name=f"{tag} n={n}")
        if "nerf" in args.which:
            p4, dirs = r(n, 4) * 0.5, r(n, 3)
            cs, cr, cd = r(n, 1), r(n, 3), r(n, 96)
            for want_dx in ([False] if args.no_dx else [False, True]):
                bad += three_way(lambda: run_nerf(nerf, p4, dirs, cs, cr, cd, want_dx), tol_grad=1e-1, l2_grad=True,
                                 name=f"NeRF n={n} dx={want_dx}")
    print("FAILURES:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
