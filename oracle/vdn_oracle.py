"""CPU oracle for the VDN-NeRF neural-SDF volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the
reference algorithm, written as pure functions over ``state_dict``-style parameter dictionaries.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; nothing under ``vdn_nerf_b200/`` does.

Parity status: **pinned against the live reference**.  The reference ships no tests or golden
vectors, so ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (with ``mcubes``/``icecream`` stubbed), runs it on seeded inputs, checks that
this restatement reproduces it bit-for-bit on CPU and writes ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` re-checks the restatement against those fixtures everywhere.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# L0  embedder                                                    dpt_models/embedder.py:11-36
# --------------------------------------------------------------------------------------------
def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    """[x | sin(2^0 x) | cos(2^0 x) | ... ] (embedder.py:15-36; log-sampled bands, line 23)."""
    if multires <= 0:
        return x
    bands = 2.0 ** torch.linspace(0.0, multires - 1, multires)
    pieces = [x]
    for f in bands:
        f = f.to(x.dtype)
        pieces.append(torch.sin(x * f))
        pieces.append(torch.cos(x * f))
    return torch.cat(pieces, dim=-1)


# --------------------------------------------------------------------------------------------
# L1  fields                                                          dpt_models/fields.py
# --------------------------------------------------------------------------------------------
def effective_weight(p: Params, name: str) -> torch.Tensor:
    """Old-style weight_norm, dim=0: W = g * v / ||v||_row (fields.py:65-66, 141-142)."""
    if name + ".weight_v" in p:
        v = p[name + ".weight_v"]
        g = p[name + ".weight_g"]
        # the ATen op nn.utils.weight_norm's hook calls (its fused CPU kernel rounds the row norm differently
        # from v.norm(dim=1), so the composite formula would not be bit-identical to the reference)
        return torch._weight_norm(v, g, 0)
    return p[name + ".weight"]


def softplus100(x: torch.Tensor) -> torch.Tensor:
    """nn.Softplus(beta=100) with the default threshold 20 (fields.py:70)."""
    return F.softplus(x, beta=100.0, threshold=20.0)


@dataclass
class SDFSpec:
    n_lin: int = 9            # number of linear layers (n_layers + 1)
    skip_in: tuple = (4,)
    multires: int = 6
    scale: float = 1.0


def sdf_forward(p: Params, x: torch.Tensor, spec: SDFSpec = SDFSpec()) -> torch.Tensor:
    """SDFNetwork.forward (fields.py:72-89): returns [N, d_out] = [sdf | feature]."""
    inputs = embed(x * spec.scale, spec.multires)
    h = inputs
    for l in range(spec.n_lin):
        if l in spec.skip_in:
            h = torch.cat([h, inputs], dim=1) / math.sqrt(2)
        h = F.linear(h, effective_weight(p, f"lin{l}"), p[f"lin{l}.bias"])
        if l < spec.n_lin - 1:
            h = softplus100(h)
    return torch.cat([h[:, :1] / spec.scale, h[:, 1:]], dim=-1)


def sdf_value(p: Params, x: torch.Tensor, spec: SDFSpec = SDFSpec()) -> torch.Tensor:
    """SDFNetwork.sdf (fields.py:91-92)."""
    return sdf_forward(p, x, spec)[:, :1]


def sdf_gradient(p: Params, x: torch.Tensor, spec: SDFSpec = SDFSpec(), create_graph: bool = True):
    """SDFNetwork.gradient (fields.py:97-108): d sdf / d x by autograd, returns [N, 1, 3]."""
    if not x.requires_grad:
        x.requires_grad_(True)
    with torch.enable_grad():
        y = sdf_value(p, x, spec)
        g = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph,
                                retain_graph=True, only_inputs=True)[0]
    return g.unsqueeze(1)


@dataclass
class RenderNetSpec:
    n_lin: int = 5
    mode: str = "idr"
    multires_view: int = 4
    squeeze_out: bool = True


def rendering_forward(p: Params, points, normals, view_dirs, feats,
                      spec: RenderNetSpec = RenderNetSpec()) -> torch.Tensor:
    """RenderingNetwork.forward (fields.py:148-176)."""
    v = embed(view_dirs, spec.multires_view)
    if spec.mode == "idr":
        h = torch.cat([points, v, normals, feats], dim=-1)
    elif spec.mode == "no_view_dir":
        h = torch.cat([points, normals, feats], dim=-1)
    elif spec.mode == "no_normal":
        h = torch.cat([points, v, feats], dim=-1)
    else:
        raise ValueError(spec.mode)
    for l in range(spec.n_lin):
        h = F.linear(h, effective_weight(p, f"lin{l}"), p[f"lin{l}.bias"])
        if l < spec.n_lin - 1:
            h = torch.relu(h)
    return torch.sigmoid(h) if spec.squeeze_out else torch.relu(h)


@dataclass
class NeRFSpec:
    D: int = 8
    skips: tuple = (4,)
    multires: int = 10
    multires_view: int = 4
    gen_depth_feats: bool = False


def nerf_forward(p: Params, pts, views, spec: NeRFSpec = NeRFSpec()):
    """NeRF.forward with use_viewdirs=True (fields.py:324-353): (sigma_raw, rgb, depth_feat|None)."""
    e = embed(pts, spec.multires)
    ev = embed(views, spec.multires_view)
    h = e
    for i in range(spec.D):
        h = torch.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i in spec.skips:
            h = torch.cat([e, h], dim=-1)
    sigma = F.linear(h, p["alpha_linear.weight"], p["alpha_linear.bias"])
    feat = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])
    h = torch.cat([feat, ev], dim=-1)
    h = torch.relu(F.linear(h, p["views_linears.0.weight"], p["views_linears.0.bias"]))
    rgb = F.linear(h, p["rgb_linear.weight"], p["rgb_linear.bias"])
    dpt = None
    if spec.gen_depth_feats:
        dpt = F.linear(h, p["dpt_linear.weight"], p["dpt_linear.bias"])
    return sigma, rgb, dpt


def inv_s_from_variance(variance: torch.Tensor) -> torch.Tensor:
    """SingleVarianceNetwork.forward (fields.py:363-364) followed by the caller's clip
    (renderer.py:262): a [1,1] tensor."""
    return (torch.ones([1, 1], dtype=variance.dtype) * torch.exp(variance * 10.0)).clip(1e-6, 1e6)


# --------------------------------------------------------------------------------------------
# L2  renderer                                                      dpt_models/renderer.py
# --------------------------------------------------------------------------------------------
@dataclass
class Nets:
    """Everything the renderer needs, as parameter dictionaries."""
    sdf: Params
    color: Params
    variance: torch.Tensor
    nerf: Optional[Params] = None
    depth: Optional[Params] = None
    sdf_spec: SDFSpec = field(default_factory=SDFSpec)
    color_spec: RenderNetSpec = field(default_factory=RenderNetSpec)
    depth_spec: RenderNetSpec = field(default_factory=RenderNetSpec)
    nerf_spec: NeRFSpec = field(default_factory=NeRFSpec)
    n_samples: int = 64
    n_importance: int = 64
    n_outside: int = 32
    up_sample_steps: int = 4
    perturb: float = 1.0

    def leaves(self):
        """All trainable tensors in the driver's order nerf, sdf, variance, colour, depth
        (dpt_runner.py:117-129)."""
        out = []
        for name, d in (("nerf", self.nerf), ("sdf", self.sdf)):
            if d is not None:
                out += [(f"{name}.{k}", v) for k, v in d.items()]
        out.append(("variance", self.variance))
        for name, d in (("color", self.color), ("depth", self.depth)):
            if d is not None:
                out += [(f"{name}.{k}", v) for k, v in d.items()]
        return out


def nets_from_modules(nerf, sdf, variance_net, color, depth, conf, detach=True, dtype=None) -> "Nets":
    """Build the oracle's parameter dictionaries from constructed modules (this package's or the reference's:
    both expose the same named_parameters) and a conf dict of vdn_nerf_b200.configs."""
    def grab(m):
        if m is None:
            return None
        d = {}
        for k, v in m.named_parameters():
            t = v.detach().cpu().clone() if detach else v
            if dtype is not None:
                t = t.to(dtype)
            d[k] = t
        return d
    sc, rc, nc, rr = conf["sdf_network"], conf["rendering_network"], conf["nerf"], conf["neus_renderer"]
    dc = conf.get("depth_extract_network") or rc
    var = variance_net.variance.detach().cpu().clone() if detach else variance_net.variance
    if dtype is not None:
        var = var.to(dtype)
    return Nets(
        sdf=grab(sdf), color=grab(color), variance=var, nerf=grab(nerf), depth=grab(depth),
        sdf_spec=SDFSpec(n_lin=sc["n_layers"] + 1, skip_in=tuple(sc["skip_in"]), multires=sc["multires"],
                         scale=sc["scale"]),
        color_spec=RenderNetSpec(n_lin=rc["n_layers"] + 1, mode=rc["mode"], multires_view=rc["multires_view"],
                                 squeeze_out=rc["squeeze_out"]),
        depth_spec=RenderNetSpec(n_lin=dc["n_layers"] + 1, mode=dc["mode"], multires_view=dc["multires_view"],
                                 squeeze_out=dc["squeeze_out"]),
        nerf_spec=NeRFSpec(D=nc["D"], skips=tuple(nc["skips"]), multires=nc["multires"],
                           multires_view=nc["multires_view"], gen_depth_feats=nc.get("gen_depth_feats", False)),
        n_samples=rr["n_samples"], n_importance=rr["n_importance"], n_outside=rr["n_outside"],
        up_sample_steps=rr["up_sample_steps"], perturb=rr["perturb"])


def excl_cumprod_weights(alpha: torch.Tensor) -> torch.Tensor:
    """alpha * exclusive-cumprod(1 - alpha + 1e-7) (renderer.py:126, 187-188, 301)."""
    ones = torch.ones([alpha.shape[0], 1], dtype=alpha.dtype)
    return alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-7], -1), -1)[:, :-1]


def sample_pdf_det(bins, weights, n_samples, return_inds=False):
    """sample_pdf with det=True (renderer.py:44-74)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.0 + 0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, dtype=bins.dtype)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    samples = bin_lo + t * (bin_hi - bin_lo)
    if return_inds:
        return samples, inds, cdf
    return samples


def up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s, return_inds=False):
    """NeuSRenderer.up_sample (renderer.py:147-191)."""
    B, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    sdf = sdf.reshape(B, n)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos = torch.cat([torch.zeros([B, 1], dtype=z_vals.dtype), cos_val[:, :-1]], dim=-1)
    cos_val = torch.minimum(prev_cos, cos_val)
    cos_val = cos_val.clip(-1e3, 0.0) * inside
    dist = next_z - prev_z
    prev_esti = mid_sdf - cos_val * dist * 0.5
    next_esti = mid_sdf + cos_val * dist * 0.5
    prev_cdf = torch.sigmoid(prev_esti * inv_s)
    next_cdf = torch.sigmoid(next_esti * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    weights = excl_cumprod_weights(alpha)
    return sample_pdf_det(z_vals, weights, n_importance, return_inds=return_inds)


def cat_z_vals(nets: Nets, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False, return_index=False):
    """NeuSRenderer.cat_z_vals (renderer.py:193-207)."""
    B, n = z_vals.shape
    _, m = new_z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
    z_vals = torch.cat([z_vals, new_z_vals], dim=-1)
    z_vals, index = torch.sort(z_vals, dim=-1)
    if not last:
        new_sdf = sdf_value(nets.sdf, pts.reshape(-1, 3), nets.sdf_spec).reshape(B, m)
        sdf = torch.cat([sdf, new_sdf], dim=-1)
        sdf = torch.gather(sdf, 1, index)
    if return_index:
        return z_vals, sdf, index
    return z_vals, sdf


def render_core_outside(nets: Nets, rays_o, rays_d, z_vals, sample_dist):
    """NeuSRenderer.render_core_outside (renderer.py:100-145); only the entries the caller uses."""
    B, n = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]
    r = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
    pts4 = torch.cat([pts / r, 1.0 / r], dim=-1)
    dirs = rays_d[:, None, :].expand(B, n, 3)
    sigma, rgb, dpt = nerf_forward(nets.nerf, pts4.reshape(-1, 4), dirs.reshape(-1, 3), nets.nerf_spec)
    alpha = 1.0 - torch.exp(-F.softplus(sigma.reshape(B, n)) * dists)
    out = {"alpha": alpha, "sampled_color": rgb.reshape(B, n, -1), "z_vals": mid_z,
           "sampled_feat": dpt.reshape(B, n, -1) if dpt is not None else None}
    return out


def render_core(nets: Nets, rays_o, rays_d, z_vals, sample_dist, background_alpha=None,
                background_sampled_feat=None, background_sampled_color=None, background_rgb=None,
                cos_anneal_ratio=0.0, depth_before_color=False):
    """NeuSRenderer.render_core (renderer.py:209-330)."""
    B, n = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)

    out = sdf_forward(nets.sdf, pts, nets.sdf_spec)
    sdf, feat = out[:, :1], out[:, 1:]
    # The reference turns requires_grad on in place for `pts` (fields.py:98, quirk C.10); when pts
    # comes from leaf rays that do not require grad this only matters for the autograd graph.
    if not pts.requires_grad:
        pts.requires_grad_(True)
    gradients = sdf_gradient(nets.sdf, pts, nets.sdf_spec).squeeze(1)

    sampled_feat = None
    if nets.depth is not None:
        sampled_feat = rendering_forward(nets.depth, pts, gradients, dirs, feat, nets.depth_spec)
        if depth_before_color:
            feat = torch.cat([feat, sampled_feat], dim=-1)
        sampled_feat = sampled_feat.reshape(B, n, -1)
    sampled_color = rendering_forward(nets.color, pts, gradients, dirs, feat, nets.color_spec).reshape(B, n, -1)

    inv_s = inv_s_from_variance(nets.variance).expand(B * n, 1)
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    p = prev_cdf - next_cdf
    c = prev_cdf
    alpha = ((p + 1e-5) / (c + 1e-5)).reshape(B, n).clip(0.0, 1.0)

    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(B, n)
    inside = (pts_norm < 1.0).to(z_vals.dtype).detach()
    relax = (pts_norm < 1.2).to(z_vals.dtype).detach()

    if background_alpha is not None:
        alpha = alpha * inside + background_alpha[:, :n] * (1.0 - inside)
        alpha = torch.cat([alpha, background_alpha[:, n:]], dim=-1)
        sampled_color = sampled_color * inside[:, :, None] + \
            background_sampled_color[:, :n] * (1.0 - inside)[:, :, None]
        sampled_color = torch.cat([sampled_color, background_sampled_color[:, n:]], dim=1)
        if nets.depth is not None:
            sampled_feat = sampled_feat * inside[:, :, None] + \
                background_sampled_feat[:, :n] * (1.0 - inside)[:, :, None]
            sampled_feat = torch.cat([sampled_feat, background_sampled_feat[:, n:]], dim=1)

    weights = excl_cumprod_weights(alpha)
    weights_sum = weights.sum(dim=-1, keepdim=True)
    color = (sampled_color * weights[:, :, None]).sum(dim=1)
    d_feats = None
    if nets.depth is not None:
        d_feats = (sampled_feat * weights[:, :, None]).sum(dim=1)
    if background_rgb is not None:
        color = color + background_rgb * (1.0 - weights_sum)

    g = gradients.reshape(B, n, 3)
    gerr = (torch.linalg.norm(g, ord=2, dim=-1) - 1.0) ** 2
    gerr = (relax * gerr).sum() / (relax.sum() + 1e-5)
    return {"d_feats": d_feats, "color": color, "sdf": sdf, "dists": dists, "gradients": g,
            "s_val": 1.0 / inv_s, "mid_z_vals": mid_z, "weights": weights, "cdf": c.reshape(B, n),
            "gradient_error": gerr, "inside_sphere": inside}


def coarse_z_vals(nets: Nets, near, far, perturb, dtype=torch.float32):
    """Coarse / outside sample placement (renderer.py:333-359).  Draws torch.rand in the
    reference's order ([B,1] then [B,n_outside]) when perturb > 0."""
    B = near.shape[0]
    z = torch.linspace(0.0, 1.0, nets.n_samples, dtype=dtype)
    z_vals = near + (far - near) * z[None, :]
    z_out = None
    if nets.n_outside > 0:
        z_out = torch.linspace(1e-3, 1.0 - 1.0 / (nets.n_outside + 1.0), nets.n_outside, dtype=dtype)
    if perturb > 0:
        t_rand = torch.rand([B, 1], dtype=dtype) - 0.5
        z_vals = z_vals + t_rand * 2.0 / nets.n_samples
        if nets.n_outside > 0:
            mids = 0.5 * (z_out[..., 1:] + z_out[..., :-1])
            upper = torch.cat([mids, z_out[..., -1:]], -1)
            lower = torch.cat([z_out[..., :1], mids], -1)
            t_rand = torch.rand([B, z_out.shape[-1]], dtype=dtype)
            z_out = lower[None, :] + (upper - lower)[None, :] * t_rand
    if nets.n_outside > 0:
        z_out = far / torch.flip(z_out, dims=[-1]) + 1.0 / nets.n_samples
    return z_vals, z_out


def hierarchical_z_vals(nets: Nets, rays_o, rays_d, z_vals, trace=None):
    """The up-sampling loop (renderer.py:367-385), under no_grad."""
    B = rays_o.shape[0]
    with torch.no_grad():
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
        sdf = sdf_value(nets.sdf, pts.reshape(-1, 3), nets.sdf_spec).reshape(B, nets.n_samples)
        for i in range(nets.up_sample_steps):
            new_z, inds, cdf = up_sample(rays_o, rays_d, z_vals, sdf,
                                         nets.n_importance // nets.up_sample_steps, 64 * 2 ** i,
                                         return_inds=True)
            new_z = new_z.detach()
            if trace is not None:
                trace.append({"z_in": z_vals.clone(), "sdf_in": sdf.clone(), "new_z": new_z.clone(),
                              "inds": inds.clone(), "cdf": cdf.clone()})
            z_vals, sdf, index = cat_z_vals(nets, rays_o, rays_d, z_vals, new_z, sdf,
                                            last=(i + 1 == nets.up_sample_steps), return_index=True)
            if trace is not None:
                trace[-1]["z_out"] = z_vals.clone()
                trace[-1]["sort_index"] = index.clone()
    return z_vals


def render(nets: Nets, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None,
           cos_anneal_ratio=0.0, depth_before_color=False, trace=None):
    """NeuSRenderer.render (renderer.py:332-439)."""
    B = len(rays_o)
    sample_dist = 2.0 / nets.n_samples
    perturb = nets.perturb if perturb_overwrite < 0 else perturb_overwrite
    z_vals, z_out = coarse_z_vals(nets, near, far, perturb, dtype=rays_o.dtype)
    n = nets.n_samples
    if nets.n_importance > 0:
        z_vals = hierarchical_z_vals(nets, rays_o, rays_d, z_vals, trace=trace)
        n = nets.n_samples + nets.n_importance

    bg = {"alpha": None, "sampled_color": None, "sampled_feat": None, "z_vals": None}
    if nets.n_outside > 0:
        z_feed, _ = torch.sort(torch.cat([z_vals, z_out], dim=-1), dim=-1)
        bg = render_core_outside(nets, rays_o, rays_d, z_feed, sample_dist)

    fine = render_core(nets, rays_o, rays_d, z_vals, sample_dist,
                       background_alpha=bg["alpha"], background_sampled_feat=bg["sampled_feat"],
                       background_sampled_color=bg["sampled_color"], background_rgb=background_rgb,
                       cos_anneal_ratio=cos_anneal_ratio, depth_before_color=depth_before_color)
    weights = fine["weights"]
    return {
        "render_feats": fine["d_feats"],
        "color_fine": fine["color"],
        "s_val": fine["s_val"].reshape(B, n).mean(dim=-1, keepdim=True),
        "cdf_fine": fine["cdf"],
        "weight_sum": weights.sum(dim=-1, keepdim=True),
        "weight_max": torch.max(weights, dim=-1, keepdim=True)[0],
        "gradients": fine["gradients"],
        "weights": weights,
        "z_vals": bg["z_vals"] if bg["z_vals"] is not None else fine["mid_z_vals"],
        "gradient_error": fine["gradient_error"],
        "inside_sphere": fine["inside_sphere"],
        "_fine_z_vals": z_vals,        # oracle-only extra: the up-sampled z fed to render_core
    }


def extract_fields(nets: Nets, bound_min, bound_max, resolution, block=64) -> np.ndarray:
    """extract_fields with query_func = -sdf (renderer.py:10-30, 441-446)."""
    X = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution).split(block)
    Y = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution).split(block)
    Z = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution).split(block)
    u = np.zeros([resolution] * 3, dtype=np.float32)
    with torch.no_grad():
        for xi, xs in enumerate(X):
            for yi, ys in enumerate(Y):
                for zi, zs in enumerate(Z):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                    val = (-sdf_value(nets.sdf, pts, nets.sdf_spec)).reshape(len(xs), len(ys), len(zs))
                    u[xi * block: xi * block + len(xs), yi * block: yi * block + len(ys),
                      zi * block: zi * block + len(zs)] = val.numpy()
    return u


# --------------------------------------------------------------------------------------------
# Measurement protocol of SURVEY.md section 8(d): synthetic rays and the driver's loss
# --------------------------------------------------------------------------------------------
def synthetic_rays(n_rays: int, seed: int = 1234, dtype=torch.float32):
    """Cameras on a radius-2.5 sphere looking roughly at the origin; near/far as
    Dataset.near_far_from_sphere (dpt_models/dataset.py:111-118)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n_rays, 3, generator=g)
    o = 2.5 * o / o.norm(dim=-1, keepdim=True)
    d = -o + 0.3 * torch.randn(n_rays, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    a = (d * d).sum(-1, keepdim=True)
    b = 2.0 * (o * d).sum(-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return o.to(dtype), d.to(dtype), (mid - 1.0).to(dtype), (mid + 1.0).to(dtype)


def driver_loss(out, true_rgb, igr_weight=0.1, mask_weight=0.0, gt_feats=None, depth_weight=1.0):
    """The training loss of dpt_runner.py:228-243 with mask == 1 everywhere."""
    mask = torch.ones_like(out["weight_sum"])
    mask_sum = mask.sum() + 1e-5
    color_error = (out["color_fine"] - true_rgb) * mask
    loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    loss = loss + out["gradient_error"] * igr_weight
    loss = loss + F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask) * mask_weight
    if gt_feats is not None and out.get("render_feats") is not None:
        err = (out["render_feats"] - gt_feats) * mask
        loss = loss + F.l1_loss(err, torch.zeros_like(err), reduction="sum") / mask_sum * depth_weight
    return loss
