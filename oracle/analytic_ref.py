"""CPU mirror of the hand-derived passes implemented by the CUDA kernels (TEST INFRASTRUCTURE ONLY).

`vdn_oracle.py` restates the reference and differentiates it with autograd (which is what the reference does).
This file restates, in plain torch on the CPU, the *algorithms the kernels use instead of autograd* - the
analytic normals pass, the two-phase SDF backward of SURVEY.md Appendix A exactly as csrc/sdf_net.cu sequences
it, and the closed-form compositing backward of csrc/rays.cu - so `tests/test_analytic_cpu.py` can prove the
formulas against autograd in fp64 without a GPU.  Nothing under vdn_nerf_b200/ imports it.
"""
from __future__ import annotations

import math

import torch

from . import vdn_oracle as vo

INV_SQRT2 = 1.0 / math.sqrt(2.0)


def sp(z):
    return torch.where(z * 100 > 20, z, torch.log1p(torch.exp(torch.clamp(z * 100, max=20.0))) / 100)


def sp1(z):
    return torch.where(z * 100 > 20, torch.ones_like(z), torch.sigmoid(z * 100))


def sp2(z):
    s = torch.sigmoid(z * 100)
    return torch.where(z * 100 > 20, torch.zeros_like(z), 100 * s * (1 - s))


def embed_jac(x, L):
    """J_e = d e / d x as [N, d_e, d] (block diagonal per coordinate)."""
    N, d = x.shape
    d_e = d * (1 + 2 * L)
    J = torch.zeros(N, d_e, d, dtype=x.dtype)
    for j in range(d):
        J[:, j, j] = 1.0
        for k in range(L):
            f = 2.0 ** k
            J[:, d + 2 * k * d + j, j] = f * torch.cos(f * x[:, j])
            J[:, d + 2 * k * d + d + j, j] = -f * torch.sin(f * x[:, j])
    return J


def sdf_passes(p, x, spec: vo.SDFSpec, d_sdf, d_feat, d_n, want_dx=True):
    """Forward, normals pass and the two-phase backward, layer by layer as the kernels do.
    Returns (out, normals, dW list, db list, d_x)."""
    L = spec.n_lin
    skip = spec.skip_in[0] if len(spec.skip_in) else -1
    W = [vo.effective_weight(p, f"lin{l}") for l in range(L)]
    b = [p[f"lin{l}.bias"] for l in range(L)]
    y = x * spec.scale
    e = vo.embed(y, spec.multires)
    d_e = e.shape[1]
    # forward: store Z_l and the layer inputs u_l
    Z, U = [], []
    h = e
    for l in range(L):
        u = torch.cat([h, e], 1) * INV_SQRT2 if l == skip else h
        U.append(u)
        z = u @ W[l].T + b[l]
        Z.append(z)
        h = sp(z) if l < L - 1 else z
    out = torch.cat([Z[-1][:, :1] / spec.scale, Z[-1][:, 1:]], -1)
    # normals pass: G[l] = d sdf_raw / d u_l
    G = [None] * (L + 1)
    row = W[L - 1][0:1, :]

    def gin(l):     # d sdf_raw / d h_l
        if l == L - 2:
            return row.expand(x.shape[0], -1)
        g = G[l + 1]
        if l + 1 == skip:
            return g[:, : W[l].shape[0]] * INV_SQRT2
        return g
    for l in range(L - 2, -1, -1):
        delta = sp1(Z[l]) * gin(l)
        G[l] = delta @ W[l]
    de = G[0]
    if skip >= 0:
        de = de + G[skip][:, W[skip].shape[1] - d_e:] * INV_SQRT2
    J = embed_jac(y, spec.multires)
    normals = torch.einsum("nc,ncj->nj", de, J)

    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros_like(v) for v in b]
    N = x.shape[0]
    ZG = [torch.zeros_like(Z[l]) for l in range(L - 1)]
    # phase 1
    if d_n is not None:
        deb = torch.einsum("ncj,nj->nc", J, d_n)
        qbar = deb
        for l in range(L - 1):
            dbar = qbar @ W[l].T
            gi = gin(l)
            dW[l] += (sp1(Z[l]) * gi).T @ qbar
            ZG[l] = sp2(Z[l]) * gi * dbar
            nxt = sp1(Z[l]) * dbar
            if l + 1 == skip:
                qbar = torch.cat([nxt * INV_SQRT2, deb * INV_SQRT2], 1)
            else:
                qbar = nxt
        dW[L - 1][0] += qbar.sum(0)
    # phase 2
    zl = torch.zeros_like(Z[-1])
    if d_sdf is not None:
        zl[:, 0] = d_sdf.reshape(-1) / spec.scale
    if d_feat is not None:
        zl[:, 1:] = d_feat
    zbar = zl
    ebar = torch.zeros_like(e)
    for l in range(L - 1, -1, -1):
        dW[l] += zbar.T @ U[l]
        db[l] += zbar.sum(0)
        ubar = zbar @ W[l]
        if l == 0:
            ebar = ebar + ubar
            break
        if l == skip:
            hbar = ubar[:, : Z[l - 1].shape[1]] * INV_SQRT2
            ebar = ebar + ubar[:, Z[l - 1].shape[1]:] * INV_SQRT2
        else:
            hbar = ubar
        zbar = sp1(Z[l - 1]) * hbar + ZG[l - 1]
    d_x = None
    if want_dx:
        d_x = spec.scale * torch.einsum("nc,ncj->nj", ebar, J)
        if d_n is not None:
            Lm = spec.multires
            d = x.shape[1]
            acc = torch.zeros_like(x)
            for j in range(d):
                for k in range(Lm):
                    f = 2.0 ** k
                    acc[:, j] += f * f * (-torch.sin(f * y[:, j]) * de[:, d + 2 * k * d + j]
                                          - torch.cos(f * y[:, j]) * de[:, d + 2 * k * d + d + j])
            d_x = d_x + spec.scale * d_n * acc
    return out, normals, dW, db, d_x


def weight_norm_backward(v, g, dW):
    norm = v.norm(dim=1, keepdim=True)
    dot = (dW * v).sum(1, keepdim=True)
    dg = dot / norm
    dv = (g / norm) * (dW - v * dot / (norm * norm))
    return dv, dg


def composite_forward(o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance,
                      bg_rgb, r):
    """Same quantities as vdn_composite_fwd, vectorised over rays."""
    B, S = mid_z.shape
    inv_s = torch.exp(variance * 10.0).clip(1e-6, 1e6)
    g = nrm.reshape(B, S, 3)
    tc = (d[:, None, :] * g).sum(-1)
    a1 = -tc * 0.5 + 0.5
    ic = -(torch.relu(a1) * (1 - r) + torch.relu(-tc) * r)
    s = sdf.reshape(B, S)
    en = s + ic * dists * 0.5
    ep = s - ic * dists * 0.5
    P, Nn = torch.sigmoid(ep * inv_s), torch.sigmoid(en * inv_s)
    raw = (P - Nn + 1e-5) / (P + 1e-5)
    alpha = raw.clip(0, 1)
    pn = (o[:, None, :] + d[:, None, :] * mid_z[..., None]).norm(dim=-1)
    inside = (pn < 1.0).to(mid_z.dtype)
    relax = (pn < 1.2).to(mid_z.dtype)
    c = col.reshape(B, S, 3)
    f = feat.reshape(B, S, -1) if feat is not None else None
    if sigma_bg is not None:
        NB = sigma_bg.numel() // B
        sg = sigma_bg.reshape(B, NB)
        abg = 1 - torch.exp(-torch.nn.functional.softplus(sg) * dists_bg) if dists_bg is not None else sg
        alpha = torch.cat([alpha * inside + abg[:, :S] * (1 - inside), abg[:, S:]], 1)
        cb = rgb_bg.reshape(B, NB, 3)
        c = torch.cat([c * inside[..., None] + cb[:, :S] * (1 - inside)[..., None], cb[:, S:]], 1)
        if f is not None:
            fb = feat_bg.reshape(B, NB, -1)
            f = torch.cat([f * inside[..., None] + fb[:, :S] * (1 - inside)[..., None], fb[:, S:]], 1)
    T = torch.cumprod(torch.cat([torch.ones(B, 1, dtype=alpha.dtype), 1 - alpha + 1e-7], 1), 1)[:, :-1]
    w = alpha * T
    color = (w[..., None] * c).sum(1)
    if bg_rgb is not None:
        color = color + bg_rgb.reshape(1, 3) * (1 - w.sum(1, keepdim=True))
    dfeat = (w[..., None] * f).sum(1) if f is not None else None
    gn = g.norm(dim=-1)
    eik_num = (relax * (gn - 1) ** 2).sum(1)
    eik_den = relax.sum(1)
    return w, P, inside, color, dfeat, eik_num, eik_den


def composite_backward(o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance,
                       bg_rgb, r, d_color, d_weights, d_cdf, d_dfeat, d_eik_num):
    """Closed-form backward exactly as csrc/rays.cu composite_bwd_kernel computes it (vectorised)."""
    B, S = mid_z.shape
    dt = mid_z.dtype
    inv_s = torch.exp(variance * 10.0).clip(1e-6, 1e6)
    g = nrm.reshape(B, S, 3)
    tc = (d[:, None, :] * g).sum(-1)
    a1 = -tc * 0.5 + 0.5
    ic = -(torch.relu(a1) * (1 - r) + torch.relu(-tc) * r)
    s = sdf.reshape(B, S)
    en = s + ic * dists * 0.5
    ep = s - ic * dists * 0.5
    P, Nn = torch.sigmoid(ep * inv_s), torch.sigmoid(en * inv_s)
    raw = (P - Nn + 1e-5) / (P + 1e-5)
    alpha_f = raw.clip(0, 1)
    pn = (o[:, None, :] + d[:, None, :] * mid_z[..., None]).norm(dim=-1)
    inside = (pn < 1.0).to(dt)
    relax = (pn < 1.2).to(dt)
    c = col.reshape(B, S, 3)
    f = feat.reshape(B, S, -1) if feat is not None else None
    has_bg = sigma_bg is not None
    if has_bg:
        NB = sigma_bg.numel() // B
        sg = sigma_bg.reshape(B, NB)
        if dists_bg is not None:
            spv = torch.nn.functional.softplus(sg)
            ex = torch.exp(-spv * dists_bg)
            abg = 1 - ex
        else:
            abg = sg
        alpha = torch.cat([alpha_f * inside + abg[:, :S] * (1 - inside), abg[:, S:]], 1)
        cb = rgb_bg.reshape(B, NB, 3)
        cc = torch.cat([c * inside[..., None] + cb[:, :S] * (1 - inside)[..., None], cb[:, S:]], 1)
        ins = torch.cat([inside, torch.zeros(B, NB - S, dtype=dt)], 1)
        if f is not None:
            fb = feat_bg.reshape(B, NB, -1)
            ff = torch.cat([f * inside[..., None] + fb[:, :S] * (1 - inside)[..., None], fb[:, S:]], 1)
    else:
        NB = 0
        alpha, cc, ff = alpha_f, c, f
        ins = torch.ones(B, S, dtype=dt)
    NW = alpha.shape[1]
    T = torch.cumprod(torch.cat([torch.ones(B, 1, dtype=dt), 1 - alpha + 1e-7], 1), 1)[:, :-1]
    w = alpha * T
    wbar = (cc * d_color[:, None, :]).sum(-1)
    if bg_rgb is not None:
        wbar = wbar - (d_color * bg_rgb.reshape(1, 3)).sum(-1, keepdim=True)
    if d_weights is not None:
        wbar = wbar + d_weights
    if f is not None and d_dfeat is not None:
        wbar = wbar + (ff * d_dfeat[:, None, :]).sum(-1)
    ww = wbar * w
    R = torch.flip(torch.cumsum(torch.flip(ww, [1]), 1), [1]) - ww          # sum over m > k
    abar = wbar * T - R / (1 - alpha + 1e-7)
    out = {}
    cbar = w[..., None] * d_color[:, None, :]
    fin = ins[:, :S] if has_bg else torch.ones(B, S, dtype=dt)
    out["d_col"] = (cbar[:, :S] * fin[..., None]).reshape(-1, 3)
    if has_bg:
        om = 1 - ins
        out["d_rgb_bg"] = (cbar * om[..., None]).reshape(-1, 3)
    if f is not None:
        fbar = w[..., None] * (d_dfeat[:, None, :] if d_dfeat is not None else torch.zeros(B, 1, f.shape[-1], dtype=dt))
        out["d_feat"] = (fbar[:, :S] * fin[..., None]).reshape(B * S, -1)
        if has_bg:
            out["d_feat_bg"] = (fbar * (1 - ins)[..., None]).reshape(B * NB, -1)
    abar_f = abar[:, :S] * fin
    rawbar = torch.where((raw >= 0) & (raw <= 1), abar_f, torch.zeros_like(abar_f))
    den = P + 1e-5
    Pbar = rawbar * Nn / (den * den)
    Nbar = -rawbar / den
    if d_cdf is not None:
        Pbar = Pbar + d_cdf
    tp, tn = Pbar * P * (1 - P), Nbar * Nn * (1 - Nn)
    epb, enb = tp * inv_s, tn * inv_s
    dinv = (tp * ep + tn * en).sum()
    out["d_sdf"] = (epb + enb).reshape(-1, 1)
    icb = (enb - epb) * dists * 0.5
    tcb = icb * (0.5 * (1 - r) * (a1 > 0).to(dt) + r * (-tc > 0).to(dt))
    gn = g.norm(dim=-1)
    En = d_eik_num if d_eik_num is not None else torch.zeros(B, dtype=dt)
    ek = torch.where(gn > 0, En[:, None] * relax * 2 * (gn - 1) / gn, torch.zeros_like(gn))
    out["d_nrm"] = (tcb[..., None] * d[:, None, :] + ek[..., None] * g).reshape(-1, 3)
    out["d_dirs"] = (tcb[..., None] * g).sum(1)
    rawv = torch.exp(variance * 10.0)
    passv = ((rawv >= 1e-6) & (rawv <= 1e6)).to(dt)
    out["d_variance"] = dinv * 10.0 * inv_s * passv
    if has_bg:
        abar_bg = torch.cat([abar[:, :S] * (1 - inside), abar[:, S:]], 1)
        if dists_bg is not None:
            dsp = torch.where(sg > 20, torch.ones_like(sg), torch.sigmoid(sg))
            out["d_sigma_bg"] = (abar_bg * ex * dists_bg * dsp).reshape(sigma_bg.shape)
            out["d_dists_bg"] = abar_bg * ex * spv
        else:
            out["d_sigma_bg"] = abar_bg.reshape(sigma_bg.shape)
    return out


# ---------------------------------------------------------------------------------------------------------
# Arithmetic of the fused tcgen05 chain kernel (csrc/sdf_chain_tc.cuh), emulated on the CPU
# ---------------------------------------------------------------------------------------------------------
B2 = 144.26950408889634          # beta / ln 2 for beta = 100
# degree-4 fit of log2(1+w)/w on [0,1] - the constants of softplus_base2() in csrc/sdf_chain_tc.cuh
_Q = (0.04008112847805023, -0.1803952157497406, 0.4036492109298706, -0.7047332525253296, 1.4414016008377075)


def to_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest, ties away from zero, on the 13 dropped mantissa bits: the two-instruction form the kernels
    use (csrc/gemm_tc.cuh), equal to cvt.rna.tf32.f32 for finite inputs."""
    i = x.detach().to(torch.float32).contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def softplus_base2(t: torch.Tensor) -> torch.Tensor:
    """a' = log2(1 + 2^t) = max(t, 0) + w q(w), w = 2^-|t|, in fp32 as the kernel evaluates it."""
    t = t.to(torch.float32)
    w = torch.exp2(-t.abs())
    q = torch.full_like(w, _Q[0])
    for c in _Q[1:]:
        q = q * w + c
    return w * q + torch.clamp(t, min=0.0)


def sdf_chain_emulated(p, x, spec: vo.SDFSpec, return_z: bool = False):
    """SDF value through the chain kernel's arithmetic: fp16 operands (weights and activations rounded to nearest),
    exact products, fp32 accumulation, base-2 softplus units (t = z * beta/ln2; the beta scaling cancels between
    layers), skip 1/sqrt2 applied to the skip layer's accumulator.  Returns sdf [N,1] in fp32 (and, with `return_z`,
    the pre-activations z_l = t_l * ln2/beta of layers 0..L-2 that the training-forward variant stores)."""
    f16 = lambda a: a.to(torch.float16).to(torch.float32)
    L = spec.n_lin
    skip = spec.skip_in[0] if len(spec.skip_in) else -1
    y = (x * spec.scale).to(torch.float32)
    e = vo.embed(y, spec.multires).to(torch.float32) * B2
    h = f16(e)
    zs = []
    for l in range(L):
        W = f16(vo.effective_weight(p, f"lin{l}").to(torch.float32))
        b = p[f"lin{l}.bias"].to(torch.float32) * B2
        dsc = INV_SQRT2 if l == skip else 1.0
        if l == skip:
            h = torch.cat([h, f16(e)], dim=1)
        acc = h @ W.t()                                    # fp16-representable operands: fp32 products are exact
        t = acc * dsc + b
        if l == L - 1:
            sdf = (t[:, :1] / B2) / spec.scale
            return (sdf, zs) if return_z else sdf
        zs.append(t / B2)
        h = f16(softplus_base2(t))
    raise AssertionError("unreachable")


def sdf_normals_emulated(p, x, spec: vo.SDFSpec) -> torch.Tensor:
    """Normals through the tensor-core mode's arithmetic: pre-activations from the fused chain (above), then the
    layer-wise reverse pass of csrc/sdf_net.cu (vdn_sdf_normals) with both GEMM operands rounded to tf32 and fp32
    accumulation - delta_l = sp'(z_l) * Gin_l is formed in fp32 in the operand prologue and rounded once."""
    L = spec.n_lin
    skip = spec.skip_in[0] if len(spec.skip_in) else -1
    _, Z = sdf_chain_emulated(p, x, spec, return_z=True)
    W = [vo.effective_weight(p, f"lin{l}").to(torch.float32) for l in range(L)]
    y = (x * spec.scale).to(torch.float32)
    d_e = x.shape[1] * (1 + 2 * spec.multires)
    G = [None] * (L + 1)

    def gin(l):
        if l == L - 2:
            return W[L - 1][0:1, :].expand(x.shape[0], -1)
        g = G[l + 1]
        return g[:, : W[l].shape[0]] * INV_SQRT2 if l + 1 == skip else g
    for l in range(L - 2, -1, -1):
        delta = to_tf32(sp1(Z[l]) * gin(l))
        G[l] = delta @ to_tf32(W[l])
    de = G[0]
    if skip >= 0:
        de = de + G[skip][:, W[skip].shape[1] - d_e:] * INV_SQRT2
    J = embed_jac(y, spec.multires)
    return torch.einsum("nc,ncj->nj", de, J)


def sdf_normals_chain_emulated(p, x, spec: vo.SDFSpec) -> torch.Tensor:
    """Normals through the FUSED chains' arithmetic (csrc/sdf_chains.cuh): forward as `sdf_chain_emulated`; what is saved
    per layer is a'_l = log2(1 + 2^t_l) rounded to bf16 (the only copy that reaches HBM), softplus'(z_l) is recomputed
    from it as 1 - 2^-a'; delta_l and the weights enter the reverse-pass MMAs as fp16, fp32 accumulation."""
    f16 = lambda a: a.to(torch.float16).to(torch.float32)
    b16 = lambda a: a.to(torch.bfloat16).to(torch.float32)
    L = spec.n_lin
    skip = spec.skip_in[0] if len(spec.skip_in) else -1
    y = (x * spec.scale).to(torch.float32)
    e = vo.embed(y, spec.multires).to(torch.float32) * B2
    d_e = e.shape[1]
    h = f16(e)
    saved = []
    for l in range(L - 1):
        W = f16(vo.effective_weight(p, f"lin{l}").to(torch.float32))
        b = p[f"lin{l}.bias"].to(torch.float32) * B2
        if l == skip:
            h = torch.cat([h, f16(e)], dim=1)
        t = (h @ W.t()) * (INV_SQRT2 if l == skip else 1.0) + b
        a = softplus_base2(t)
        saved.append(b16(a))
        h = f16(a)
    Wf = [vo.effective_weight(p, f"lin{l}").to(torch.float32) for l in range(L)]
    a = Wf[L - 1][0:1, :].expand(x.shape[0], -1)
    de_skip = None
    for l in range(L - 2, -1, -1):
        od = Wf[l].shape[0]
        S = 1.0 - torch.exp2(-saved[l][:, :od])
        delta = f16(S * a[:, :od])
        a = delta @ f16(Wf[l])
        if l == skip:
            a = a * INV_SQRT2
            de_skip = a[:, Wf[l].shape[1] - d_e:]
            a = a[:, : Wf[l].shape[1] - d_e]
    de = a + (de_skip if de_skip is not None else 0.0)
    J = embed_jac(y, spec.multires)
    return torch.einsum("nc,ncj->nj", de, J)


def loss_scale(*cots) -> float:
    """sigma = 2^-e with the largest |cotangent| in [2^(e-1), 2^e) (csrc/sdf_chains.cuh: amax_sigma_kernel); 1 for all-zero."""
    am = max(float(c.abs().max()) for c in cots if c is not None)
    if not (am > 0.0) or math.isinf(am):
        return 1.0
    _, e = math.frexp(am)
    return 2.0 ** -max(-100, min(100, e))


def sdf_backward_emulated(p, x, spec: vo.SDFSpec, d_sdf, d_feat, d_n, sigma=None):
    """The two-phase SDF backward of `sdf_passes` with the STORAGE ROUNDING of the fused chains (csrc/sdf_chains.cuh): fp16
    weights; every tensor the chains keep in 16 bits rounded to fp16 where the kernel stores it - the saved activations
    a' (from which softplus' and softplus'' are recomputed), the normals-pass deltas, and the cotangent tensors q-bar,
    z-bar^g, z-bar, which hold `sigma` times the true value (saturating at the fp16 range).  Products and sums are exact
    (fp64): this isolates what the 16-bit formats and the loss scale cost.  sigma None: the kernel's choice; 1.0: what
    an unscaled fp16 backward would do.  Returns (dW list, db list)."""
    dt = torch.float64
    h16 = lambda a: a.to(torch.float16).to(dt)                                   # round to nearest, fp16 range

    def c16(a, s):                                                               # cotangent storage: fp16(sigma * a) / sigma
        return (a * s).clamp(-65504.0, 65504.0).to(torch.float16).to(dt) / s
    L = spec.n_lin
    skip = spec.skip_in[0] if len(spec.skip_in) else -1
    W = [h16(vo.effective_weight(p, f"lin{l}").to(dt)) for l in range(L)]
    b = [p[f"lin{l}.bias"].to(dt) for l in range(L)]
    y = x.to(dt) * spec.scale
    e = vo.embed(y, spec.multires)
    er = h16(e * B2) / B2                                                        # E16 holds kB2 * e
    d_e = e.shape[1]
    if sigma is None:
        # |J_e n-bar| <= 2^multires |n-bar| (sdf_chain_backward's bound)
        sigma = loss_scale(d_feat, None if d_sdf is None else d_sdf / spec.scale,
                           None if d_n is None else d_n * float(2 ** spec.multires))
    # forward with rounded saved activations; S1 = softplus' recomputed from the saved a' = kB2 * softplus(z)
    H, U, S1 = [], [], []
    h = er
    for l in range(L):
        u = torch.cat([h, er], 1) * INV_SQRT2 if l == skip else h
        U.append(u)
        z = u @ W[l].T + b[l]
        if l < L - 1:
            a = h16(sp(z) * B2)                                                  # A16_l
            h = a / B2
            H.append(h)
            S1.append(1.0 - torch.exp2(-a))
    # normals pass with rounded deltas (D16_l)
    G = [None] * (L + 1)
    row = W[L - 1][0:1, :]
    D = [None] * (L - 1)

    def gin(l):
        if l == L - 2:
            return row.expand(x.shape[0], -1)
        g = G[l + 1]
        return g[:, : W[l].shape[0]] * INV_SQRT2 if l + 1 == skip else g
    for l in range(L - 2, -1, -1):
        D[l] = h16(S1[l] * gin(l))
        G[l] = D[l] @ W[l]
    J = embed_jac(y, spec.multires)
    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros_like(v) for v in b]
    ZG = [None] * (L - 1)
    # phase 1: q-bar_l and z-bar^g_l are stored (Q16, ZG16)
    if d_n is not None:
        deb = c16(torch.einsum("ncj,nj->nc", J, d_n.to(dt)), sigma)
        qbar = deb
        for l in range(L - 1):
            dbar = qbar @ W[l].T
            dW[l] += D[l].T @ qbar
            ZG[l] = c16(100.0 * (1.0 - S1[l]) * D[l] * dbar, sigma)             # softplus''(z) * gin = 100 (1 - S) delta
            nxt = S1[l] * dbar
            qbar = c16(torch.cat([nxt * INV_SQRT2, deb * INV_SQRT2], 1) if l + 1 == skip else nxt, sigma)
        dW[L - 1][0] += qbar.sum(0)
    # phase 2: z-bar_l is stored (FB16 / SB, ZB16)
    zl = torch.zeros(x.shape[0], W[-1].shape[0], dtype=dt)
    if d_sdf is not None:
        zl[:, 0] = d_sdf.reshape(-1).to(dt) / spec.scale
    if d_feat is not None:
        zl[:, 1:] = d_feat.to(dt)
    zbar = c16(zl, sigma)
    for l in range(L - 1, -1, -1):
        dW[l] += zbar.T @ U[l]
        db[l] += zbar.sum(0)
        if l == 0:
            break
        ubar = zbar @ W[l]
        hbar = ubar[:, : H[l - 1].shape[1]] * INV_SQRT2 if l == skip else ubar
        zb = S1[l - 1] * hbar
        if ZG[l - 1] is not None:
            zb = zb + ZG[l - 1]
        zbar = c16(zb, sigma)
    return dW, db
