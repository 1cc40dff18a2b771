"""fp64 proof, on the CPU, that the hand-derived passes the kernels implement equal autograd of the oracle."""
import math

import pytest
import torch

from oracle import analytic_ref as ar
from oracle import vdn_oracle as vo


def _sdf_params(dtype, d_hidden=64, n_lin=9, skip=4, multires=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    d_e = 3 * (1 + 2 * multires)
    p = {}
    for l in range(n_lin):
        fin = d_e if l == 0 else d_hidden
        fout = 1 + d_hidden if l == n_lin - 1 else d_hidden
        if l + 1 == skip:
            fout -= d_e
        p[f"lin{l}.weight_v"] = (torch.randn(fout, fin, generator=g) / math.sqrt(fin)).to(dtype)
        p[f"lin{l}.weight_g"] = (0.5 + torch.rand(fout, 1, generator=g)).to(dtype)
        p[f"lin{l}.bias"] = (0.05 * torch.randn(fout, generator=g)).to(dtype)
    return p


@pytest.mark.parametrize("skip", [4, -1])
def test_sdf_two_phase_backward_matches_autograd(skip):
    dt = torch.float64
    spec = vo.SDFSpec(n_lin=9, skip_in=(skip,) if skip >= 0 else (), multires=6, scale=1.0)
    p = _sdf_params(dt, skip=skip)
    for v in p.values():
        v.requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    N = 37
    x = (0.6 * torch.randn(N, 3, generator=g)).to(dt).requires_grad_(True)
    d_sdf = torch.randn(N, 1, generator=g).to(dt)
    d_feat = torch.randn(N, 64, generator=g).to(dt)
    d_n = torch.randn(N, 3, generator=g).to(dt)
    # autograd through the restated reference (double backward)
    out = vo.sdf_forward(p, x, spec)
    nrm = vo.sdf_gradient(p, x, spec).squeeze(1)
    loss = (out[:, :1] * d_sdf).sum() + (out[:, 1:] * d_feat).sum() + (nrm * d_n).sum()
    names = list(p.keys())
    grads = torch.autograd.grad(loss, [p[k] for k in names] + [x])
    ref = dict(zip(names, grads[:-1]))
    # analytic passes
    with torch.no_grad():
        out2, nrm2, dW, db, d_x = ar.sdf_passes(p, x.detach(), spec, d_sdf, d_feat, d_n)
    assert torch.allclose(out2, out.detach(), rtol=1e-10, atol=1e-12)
    assert torch.allclose(nrm2, nrm.detach(), rtol=1e-9, atol=1e-11)
    for l in range(spec.n_lin):
        dv, dg = ar.weight_norm_backward(p[f"lin{l}.weight_v"].detach(), p[f"lin{l}.weight_g"].detach(), dW[l])
        for name, got in ((f"lin{l}.weight_v", dv), (f"lin{l}.weight_g", dg), (f"lin{l}.bias", db[l])):
            want = ref[name]
            assert torch.allclose(got, want, rtol=1e-7, atol=1e-9 * (1 + want.abs().max())), name
    assert torch.allclose(d_x, grads[-1], rtol=1e-7, atol=1e-9 * (1 + grads[-1].abs().max()))


@pytest.mark.parametrize("with_bg,with_feat,alpha_mode", [(True, True, False), (True, False, True), (False, False, False),
                                                           (False, True, False)])
def test_composite_closed_form_backward(with_bg, with_feat, alpha_mode):
    dt = torch.float64
    g = torch.Generator().manual_seed(3)
    B, S, NB, Fd = 5, 16, 22, 7
    o = torch.randn(B, 3, generator=g).to(dt)
    o = 1.4 * o / o.norm(dim=-1, keepdim=True)
    d = (-o + 0.3 * torch.randn(B, 3, generator=g).to(dt))
    d = (d / d.norm(dim=-1, keepdim=True)).requires_grad_(True)
    mid = (1.4 + torch.sort(torch.rand(B, S, generator=g), dim=1)[0].to(dt) * 1.2 - 0.6)
    dists = 0.02 + 0.05 * torch.rand(B, S, generator=g).to(dt)
    leaf = lambda *s: torch.randn(*s, generator=g).to(dt).requires_grad_(True)
    sdf = (0.1 * torch.randn(B * S, 1, generator=g)).to(dt).requires_grad_(True)
    nrm, col = leaf(B * S, 3), leaf(B * S, 3)
    feat = leaf(B * S, Fd) if with_feat else None
    variance = torch.tensor(0.3, dtype=dt, requires_grad=True)
    sigma_bg = rgb_bg = feat_bg = dists_bg = None
    if with_bg:
        if alpha_mode:
            sigma_bg = torch.rand(B, NB, generator=g).to(dt).mul(0.5).requires_grad_(True)
        else:
            sigma_bg = leaf(B * NB, 1)
            dists_bg = (0.02 + 0.05 * torch.rand(B, NB, generator=g).to(dt)).requires_grad_(True)
        rgb_bg = leaf(B * NB, 3)
        feat_bg = leaf(B * NB, Fd) if with_feat else None
    bg_rgb = torch.ones(1, 3, dtype=dt)
    r = 0.5
    w, cdf, inside, color, dfeat, en, ed = ar.composite_forward(o, d, mid, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg,
                                                                feat_bg, dists_bg, variance, bg_rgb, r)
    cot = lambda t: torch.randn(t.shape, generator=g).to(dt)
    d_color, d_w, d_cdf, d_en = cot(color), cot(w), cot(cdf), cot(en)
    d_dfeat = cot(dfeat) if dfeat is not None else None
    loss = (color * d_color).sum() + (w * d_w).sum() + (cdf * d_cdf).sum() + (en * d_en).sum()
    if dfeat is not None:
        loss = loss + (dfeat * d_dfeat).sum()
    wrt = {"d_sdf": sdf, "d_nrm": nrm, "d_col": col, "d_variance": variance, "d_dirs": d}
    if with_feat:
        wrt["d_feat"] = feat
    if with_bg:
        wrt["d_sigma_bg"] = sigma_bg
        wrt["d_rgb_bg"] = rgb_bg
        if not alpha_mode:
            wrt["d_dists_bg"] = dists_bg
        if with_feat:
            wrt["d_feat_bg"] = feat_bg
    grads = dict(zip(wrt.keys(), torch.autograd.grad(loss, list(wrt.values()), allow_unused=True)))
    with torch.no_grad():
        got = ar.composite_backward(o, d.detach(), mid, dists, sdf.detach(), nrm.detach(), col.detach(),
                                    feat.detach() if feat is not None else None,
                                    sigma_bg.detach() if sigma_bg is not None else None,
                                    rgb_bg.detach() if rgb_bg is not None else None,
                                    feat_bg.detach() if feat_bg is not None else None,
                                    dists_bg.detach() if dists_bg is not None else None, variance.detach(), bg_rgb, r,
                                    d_color, d_w, d_cdf, d_dfeat, d_en)
    for k, want in grads.items():
        if k == "d_dirs":
            # the kernel returns only the true_cos path of d (pts_norm masks are detached in the reference)
            pass
        assert torch.allclose(got[k], want, rtol=1e-8, atol=1e-10 * (1 + want.abs().max())), k


def test_tf32_rounding_trick_matches_round_to_nearest_ties_away():
    """The two-instruction rounding of the tensor-core operands equals round-to-nearest (ties away) on 10 mantissa bits."""
    g = torch.Generator().manual_seed(3)
    x = torch.cat([torch.randn(20000, generator=g) * 10 ** torch.randint(-6, 6, (20000,), generator=g).float(),
                   torch.tensor([0.0, -0.0, 1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -1.0 - 2 ** -11, 3.4e38 * 0.5, 1e-38])])
    got = ar.to_tf32(x).double()
    xd = x.double()
    ex = torch.floor(torch.log2(xd.abs().clamp(min=1e-300)))
    ulp = torch.pow(torch.tensor(2.0, dtype=torch.float64), ex - 10)
    want = torch.sign(xd) * torch.floor(xd.abs() / ulp + 0.5) * ulp
    normal = xd.abs() >= 2.0 ** -126
    assert torch.equal(got[normal], want[normal])
    assert bool((got[~normal].abs() <= xd[~normal].abs() + 2.0 ** -136).all())
    assert bool(((ar.to_tf32(x).view(torch.int32) & 0x1FFF) == 0).all())


def test_base2_softplus_polynomial_error_budget():
    """a' = log2(1+2^t) through the kernel's degree-4 polynomial: |err| < 3.3e-5 in a' (2.3e-7 in softplus units), an
    order of magnitude below the fp16 rounding of a' - and the base-2 form is softplus(beta=100) exactly."""
    t = torch.linspace(-40.0, 40.0, 400001, dtype=torch.float64)
    want = torch.log2(1.0 + torch.exp2(-t.abs())) + t.clamp(min=0.0)
    got = ar.softplus_base2(t.float()).double()
    assert float((got - want).abs().max()) < 3.4e-5
    z = t / ar.B2
    assert float((want / ar.B2 - vo.softplus100(z)).abs().max()) < 1e-9      # threshold branch differs by < 3e-9


def test_chain_kernel_arithmetic_within_tensor_core_tolerance():
    """fp16 operands + fp32 accumulation + base-2 softplus, emulated on the CPU for the reference's geometric
    initialisation (womsk_white widths): the SDF value stays within the 2e-3 tensor-core tolerance of the fp64 oracle."""
    from tests import util
    mods, conf = util.build("womsk_white")
    nets64 = util.oracle_nets(mods, conf, dtype=torch.float64)
    nets32 = util.oracle_nets(mods, conf, dtype=torch.float32)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4096, 3, generator=g) * 2.4 - 1.2
    want = vo.sdf_value(nets64.sdf, x.double(), nets64.sdf_spec)
    got = ar.sdf_chain_emulated(nets32.sdf, x, nets32.sdf_spec).double()
    err = float((got - want).abs().max() / want.abs().max())
    print(f"chain arithmetic (CPU emulation) rel err {err:.2e}")
    assert err < 2e-3
    # normals: chain pre-activations + tf32 layer-wise reverse pass (vdn_sdf_normals' arithmetic)
    wantn = vo.sdf_gradient(nets64.sdf, x.double(), nets64.sdf_spec).detach().squeeze(1)
    gotn = ar.sdf_normals_emulated(nets32.sdf, x, nets32.sdf_spec).double()
    errn = float((gotn - wantn).abs().max() / wantn.abs().max())
    print(f"tensor-core normals (CPU emulation) rel err {errn:.2e}")
    assert errn < 2e-3


def test_fused_chain_normals_arithmetic_with_bf16_saved_activations():
    """The fused training chains keep ONE copy of every layer's activation in HBM, in bf16, and recompute softplus' from
    it (csrc/sdf_chains.cuh).  Emulated on the CPU at the reference's initialisation: the normals stay within the 2e-3
    tensor-core tolerance, no worse than with the tf32 layer-wise pass."""
    from tests import util
    mods, conf = util.build("womsk_white")
    nets64 = util.oracle_nets(mods, conf, dtype=torch.float64)
    nets32 = util.oracle_nets(mods, conf, dtype=torch.float32)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4096, 3, generator=g) * 2.4 - 1.2
    wantn = vo.sdf_gradient(nets64.sdf, x.double(), nets64.sdf_spec).detach().squeeze(1)
    gotn = ar.sdf_normals_chain_emulated(nets32.sdf, x, nets32.sdf_spec).double()
    errn = float((gotn - wantn).abs().max() / wantn.abs().max())
    print(f"fused-chain normals (CPU emulation, bf16 saved activations) rel err {errn:.2e}")
    assert errn < 2e-3


@pytest.mark.parametrize("cot_scale", [1.0, 1e-6])
def test_loss_scaled_fp16_backward_error_budget(cot_scale):
    """The fused backward chains store every cotangent tensor in fp16 times a per-call power-of-two loss scale
    (csrc/sdf_chains.cuh).  Emulated on the CPU at the reference's initialisation (storage rounding only, exact sums):
    parameter gradients stay within the chain tolerance of the fp64 analytic backward whatever the magnitude of the
    incoming cotangents - while an UNSCALED fp16 backward of cotangents of the size a 512-ray step really produces
    (1e-6 and below) loses them to underflow."""
    from tests import util
    mods, conf = util.build("womsk_white")
    nets = util.oracle_nets(mods, conf, dtype=torch.float64)
    p, spec = nets.sdf, nets.sdf_spec
    g = torch.Generator().manual_seed(5)
    n = 600
    x = (torch.rand(n, 3, generator=g) * 2.0 - 1.0).double()
    d_sdf = torch.randn(n, 1, generator=g).double() * cot_scale
    d_feat = torch.randn(n, 256, generator=g).double() * 0.1 * cot_scale
    d_n = torch.randn(n, 3, generator=g).double() * cot_scale
    with torch.no_grad():
        _, _, dW, db, _ = ar.sdf_passes(p, x, spec, d_sdf, d_feat, d_n, want_dx=False)
        dWs, dbs = ar.sdf_backward_emulated(p, x, spec, d_sdf, d_feat, d_n)
        dW1, db1 = ar.sdf_backward_emulated(p, x, spec, d_sdf, d_feat, d_n, sigma=1.0)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    worst = max(max(rel(a, b) for a, b in zip(dWs, dW)), max(rel(a, b) for a, b in zip(dbs, db)))
    worst1 = max(max(rel(a, b) for a, b in zip(dW1, dW)), max(rel(a, b) for a, b in zip(db1, db)))
    print(f"cotangents x {cot_scale:g}: loss-scaled fp16 backward rel err {worst:.2e}, unscaled {worst1:.2e}")
    assert worst < 1e-2
    if cot_scale < 1e-3:
        assert worst1 > 10 * worst           # without the scale the small cotangents are flushed / badly quantised
