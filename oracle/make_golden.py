"""Pin the oracle against the live reference and mint the golden fixtures under tests/golden/.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the unmodified reference is mounted at
/root/reference (it does not exist on the GPU box, so only the committed .npz files travel):

    python -m oracle.make_golden

For both shipped network configurations it (1) builds the reference's networks and this package's host-side
modules from the same seed and checks the initial parameters are bit-identical, (2) runs the reference's own
`NeuSRenderer.render`, `render_core`, `up_sample`, `cat_z_vals`, `SDFNetwork.gradient`, `extract_fields` ... on
seeded synthetic rays on the CPU, (3) asserts that oracle/vdn_oracle.py reproduces every one of them
bit-for-bit, and (4) writes inputs + reference outputs (+ digests of weights and of parameter gradients) as
small .npz fixtures.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))
    ic = types.ModuleType("icecream")
    ic.ic = lambda *a, **k: None
    sys.modules.setdefault("icecream", ic)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dpt_models.fields as rf
    import dpt_models.renderer as rr
    import dpt_models.embedder as re_
    return rf, rr, re_


def digest(t: torch.Tensor) -> np.ndarray:
    """[sum, l2, absmax, 16 sampled entries] of a tensor, as float64."""
    a = t.detach().cpu().double().reshape(-1).numpy()
    idx = (np.arange(16, dtype=np.int64) * 2654435761) % a.size
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def same(a, b, what):
    if a is None and b is None:
        return
    if not torch.equal(a, b):
        diff = (a.double() - b.double()).abs().max().item()
        raise AssertionError(f"oracle != reference for {what}: max |diff| = {diff:g}")


def main():
    import warnings
    warnings.simplefilter("ignore")
    from oracle import vdn_oracle as vo
    from vdn_nerf_b200 import configs
    from vdn_nerf_b200 import fields as my_fields
    rf, rr, re_ = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    for name, B in (("womsk_white", 16), ("womsk_white_wdepth", 8)):
        conf = configs.CONFIGS[name]
        ref_nets = configs.build_networks(conf, rf, seed=0)
        my_nets = configs.build_networks(conf, my_fields, seed=0)
        # (1) host-side init parity
        fx = {}
        for tag, rn, mn in zip(("nerf", "sdf", "variance", "color", "depth"), ref_nets, my_nets):
            if rn is None:
                continue
            rs, ms = rn.state_dict(), mn.state_dict()
            assert list(rs.keys()) == list(ms.keys()), (tag, list(rs.keys()), list(ms.keys()))
            for k in rs:
                assert torch.equal(rs[k], ms[k]), f"init mismatch {tag}.{k}"
                fx[f"wdigest/{tag}.{k}"] = digest(rs[k])
        nerf, sdf, var, col, dep = ref_nets
        nets = vo.nets_from_modules(nerf, sdf, var, col, dep, conf)
        rend = rr.NeuSRenderer(nerf, sdf, var, col, dep, **conf["neus_renderer"])
        o, d, near, far = vo.synthetic_rays(B)
        bg = torch.ones(1, 3)
        fx.update(rays_o=o.numpy(), rays_d=d.numpy(), near=near.numpy(), far=far.numpy())

        # (2) full render, deterministic sample placement
        ref_out = rend.render(o, d, near, far, perturb_overwrite=0, background_rgb=bg, cos_anneal_ratio=0.5)
        trace = []
        ora_out = vo.render(nets, o, d, near, far, perturb_overwrite=0, background_rgb=bg, cos_anneal_ratio=0.5,
                            trace=trace)
        for k, v in ref_out.items():
            same(v, ora_out[k], f"{name}/render/{k}")
            if v is not None:
                fx[f"render/{k}"] = v.detach().numpy()
        z_fine = ora_out["_fine_z_vals"]
        fx["render/fine_z_vals"] = z_fine.numpy()

        # (3) the up-sampling loop, stage by stage, against the reference's own methods
        with torch.no_grad():
            z = near + (far - near) * torch.linspace(0.0, 1.0, rend.n_samples)[None, :]
            s = sdf.sdf((o[:, None, :] + d[:, None, :] * z[..., :, None]).reshape(-1, 3)).reshape(B, -1)
            for i, tr in enumerate(trace):
                same(z, tr["z_in"], f"{name}/up{i}/z_in")
                same(s, tr["sdf_in"], f"{name}/up{i}/sdf_in")
                new_z = rend.up_sample(o, d, z, s, rend.n_importance // rend.up_sample_steps, 64 * 2 ** i)
                same(new_z, tr["new_z"], f"{name}/up{i}/new_z")
                z, s = rend.cat_z_vals(o, d, z, new_z, s, last=(i + 1 == rend.up_sample_steps))
                same(z, tr["z_out"], f"{name}/up{i}/z_out")
                for k in ("z_in", "sdf_in", "new_z", "inds", "z_out", "sort_index"):
                    fx[f"up{i}/{k}"] = tr[k].numpy()
            same(z, z_fine, f"{name}/final z")

        # (4) render_core alone (BASELINE cfg 1: no background), forward + backward of the driver loss
        for n_ in (nets,):
            for _, t in n_.leaves():
                t.requires_grad_(True)
        oo, dd = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
        core_ref = rend.render_core(o, d, z_fine, 2.0 / rend.n_samples, sdf, var, col, dep, background_rgb=bg,
                                    cos_anneal_ratio=0.5)
        core_ora = vo.render_core(nets, oo, dd, z_fine, 2.0 / nets.n_samples, background_rgb=bg, cos_anneal_ratio=0.5)
        for k in ("color", "sdf", "gradients", "weights", "cdf", "gradient_error", "inside_sphere", "d_feats", "s_val"):
            same(core_ref[k], core_ora[k], f"{name}/core/{k}")
            if core_ref[k] is not None:
                fx[f"core/{k}"] = core_ref[k].detach().numpy()

        def loss_of(core):
            ws = core["weights"].sum(dim=-1, keepdim=True)
            fake = {"color_fine": core["color"], "gradient_error": core["gradient_error"], "weight_sum": ws,
                    "render_feats": core["d_feats"]}
            gt = torch.full_like(core["d_feats"], 0.5) if core["d_feats"] is not None else None
            return vo.driver_loss(fake, torch.full((B, 3), 0.5), gt_feats=gt)
        loss_ref = loss_of(core_ref)
        for m in (sdf, var, col, dep):
            if m is not None:
                m.zero_grad()
        loss_ref.backward()
        loss_ora = loss_of(core_ora)
        leaves = [(k, t) for k, t in nets.leaves() if not k.startswith("nerf.")]
        grads = torch.autograd.grad(loss_ora, [t for _, t in leaves] + [oo, dd], allow_unused=True)
        same(loss_ref.detach(), loss_ora.detach(), f"{name}/core/loss")
        fx["core/loss"] = loss_ref.detach().numpy()
        ref_params = {}
        for tag, m in (("sdf", sdf), ("color", col), ("depth", dep)):
            if m is not None:
                ref_params.update({f"{tag}.{k}": p for k, p in m.named_parameters()})
        ref_params["variance"] = var.variance
        for (k, _), g in zip(leaves, grads):
            same(ref_params[k].grad, g, f"{name}/core/grad/{k}")
            fx[f"core_grad/{k}"] = digest(g)
        fx["core_grad/rays_o"] = grads[-2].numpy()
        fx["core_grad/rays_d"] = grads[-1].numpy()

        # (5) full training-step gradients (render + driver loss), digests only
        for m in (nerf, sdf, var, col, dep):
            if m is not None:
                m.zero_grad()
        full = rend.render(o, d, near, far, perturb_overwrite=0, background_rgb=bg, cos_anneal_ratio=0.5)
        gt = torch.full_like(full["render_feats"], 0.5) if full["render_feats"] is not None else None
        loss_full = vo.driver_loss(full, torch.full((B, 3), 0.5), gt_feats=gt)
        loss_full.backward()
        ora_full = vo.render(nets, o, d, near, far, perturb_overwrite=0, background_rgb=bg, cos_anneal_ratio=0.5)
        loss_full_ora = vo.driver_loss(ora_full, torch.full((B, 3), 0.5), gt_feats=gt)
        same(loss_full.detach(), loss_full_ora.detach(), f"{name}/step/loss")
        all_leaves = nets.leaves()
        g_full = torch.autograd.grad(loss_full_ora, [t for _, t in all_leaves], allow_unused=True)
        ref_params.update({f"nerf.{k}": p for k, p in nerf.named_parameters()})
        fx["step/loss"] = loss_full.detach().numpy()
        for (k, _), g in zip(all_leaves, g_full):
            rg = ref_params[k].grad
            if g is None:
                assert rg is None or float(rg.abs().max()) == 0.0, k
                continue
            same(rg, g, f"{name}/step/grad/{k}")
            fx[f"step_grad/{k}"] = digest(g)

        # (6) field-level known answers on a handful of points (white only carries them)
        if name == "womsk_white":
            g = torch.Generator().manual_seed(7)
            x = (torch.rand(48, 3, generator=g) * 2.4 - 1.2)
            v = torch.randn(48, 3, generator=g)
            v = v / v.norm(dim=-1, keepdim=True)
            p4 = torch.cat([v * 0.9, torch.rand(48, 1, generator=g)], -1)
            with torch.no_grad():
                e6 = re_.get_embedder(6, 3)[0](x)
                e4 = re_.get_embedder(4, 3)[0](v)
                e10 = re_.get_embedder(10, 4)[0](p4)
                so = sdf(x)
                sig, rgb, _ = nerf(p4, v)
            xg = x.clone()
            sg = sdf.gradient(xg).detach().squeeze(1)
            with torch.no_grad():
                co = col(x, sg, v, so[:, 1:])
            same(e6, vo.embed(x, 6), "embed6"); same(e4, vo.embed(v, 4), "embed4"); same(e10, vo.embed(p4, 10), "embed10")
            nets_ng = vo.nets_from_modules(nerf, sdf, var, col, dep, conf)
            same(so, vo.sdf_forward(nets_ng.sdf, x, nets_ng.sdf_spec), "sdf_forward")
            same(sg, vo.sdf_gradient(nets_ng.sdf, x.clone(), nets_ng.sdf_spec).detach().squeeze(1), "sdf_gradient")
            same(co, vo.rendering_forward(nets_ng.color, x, sg, v, so[:, 1:], nets_ng.color_spec), "color")
            s2, r2, _ = vo.nerf_forward(nets_ng.nerf, p4, v, nets_ng.nerf_spec)
            same(sig, s2, "nerf sigma"); same(rgb, r2, "nerf rgb")
            fx.update({"field/x": x.numpy(), "field/v": v.numpy(), "field/p4": p4.numpy(), "field/embed6": e6.numpy(),
                       "field/embed4": e4.numpy(), "field/embed10": e10.numpy(), "field/sdf_out": so.numpy(),
                       "field/sdf_grad": sg.numpy(), "field/color": co.numpy(), "field/nerf_sigma": sig.numpy(),
                       "field/nerf_rgb": rgb.numpy()})
            # (7) extract_fields on a ragged 72^3 grid (one full 64-block + an 8-wide remainder per axis)
            res = 72
            bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
            u_ref = rr.extract_fields(bmin, bmax, res, lambda pts: -sdf.sdf(pts))
            u_ora = vo.extract_fields(nets_ng, bmin, bmax, res)
            assert np.array_equal(u_ref, u_ora), "extract_fields"
            fx["grid/res"] = np.array(res)
            fx["grid/u_sub"] = u_ref[::3, ::3, ::3].copy()
            fx["grid/u_digest"] = digest(torch.from_numpy(u_ref))
        else:
            with torch.no_grad():
                g = torch.Generator().manual_seed(7)
                x = (torch.rand(32, 3, generator=g) * 2.0 - 1.0)
                v = torch.randn(32, 3, generator=g)
                v = v / v.norm(dim=-1, keepdim=True)
                p4 = torch.cat([v * 0.9, torch.rand(32, 1, generator=g)], -1)
                so = sdf(x)
                sig, rgb, dpt = nerf(p4, v)
            sg = sdf.gradient(x.clone()).detach().squeeze(1)
            with torch.no_grad():
                do = dep(x, sg, v, so[:, 1:])
            nets_ng = vo.nets_from_modules(nerf, sdf, var, col, dep, conf)
            same(do, vo.rendering_forward(nets_ng.depth, x, sg, v, so[:, 1:], nets_ng.depth_spec), "depth head")
            s2, r2, d2 = vo.nerf_forward(nets_ng.nerf, p4, v, nets_ng.nerf_spec)
            same(dpt, d2, "nerf dpt")
            fx.update({"field/x": x.numpy(), "field/v": v.numpy(), "field/p4": p4.numpy(), "field/sdf_grad": sg.numpy(),
                       "field/sdf_feat": so[:, 1:].numpy(), "field/depth_out": do.numpy(), "field/nerf_dpt": dpt.numpy(),
                       "field/nerf_sigma": sig.numpy(), "field/nerf_rgb": rgb.numpy()})

        path = os.path.join(out_dir, f"{name}_b{B}.npz")
        np.savez_compressed(path, **fx)
        print(f"wrote {path}: {len(fx)} arrays, {os.path.getsize(path) / 1024:.0f} KiB; oracle == reference on all of them")


if __name__ == "__main__":
    main()
