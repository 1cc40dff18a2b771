"""Field networks with the reference's constructor kwargs, parameter names and call signatures.

Drop-in for ``dpt_models/fields.py`` of the reference: ``SDFNetwork`` (fields.py:9-108),
``RenderingNetwork`` (112-176), ``NeRF`` (264-355), ``SingleVarianceNetwork`` (358-364).  The modules own
ordinary ``nn.Parameter``s under the reference's ``state_dict`` keys (``lin{l}.weight_g / weight_v / bias``
from old-style ``nn.utils.weight_norm``; ``pts_linears.{i}.*`` ... for the NeRF field) and consume the RNG
exactly as the reference constructors do, so ``torch.manual_seed(s)`` followed by construction yields
bit-identical initial parameters.  All arithmetic of ``forward`` runs in the sm_100a kernels of
``libvdn_b200.so`` through ``ops``; there is no PyTorch fallback.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .embedder import embed_out_dim


def _weight_norm(lin: nn.Linear) -> nn.Linear:
    # Old-style weight_norm gives the reference's parameter names (lin{l}.weight_g / lin{l}.weight_v).
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return nn.utils.weight_norm(lin)


def _triple(lin: nn.Module):
    """(weight-or-v, g-or-None, bias) of a linear layer, whether or not it is weight-normed."""
    if hasattr(lin, "weight_v"):
        return (lin.weight_v, lin.weight_g, lin.bias)
    return (lin.weight, None, lin.bias)


class SDFNetwork(nn.Module):
    """Softplus(beta=100) MLP with one skip connection; outputs [sdf | feature] (reference fields.py:9-108)."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        self.d_in, self.d_out, self.d_hidden, self.n_layers = d_in, d_out, d_hidden, n_layers
        self.multires = multires
        self.skip_in = tuple(skip_in)
        self.scale = scale
        d_e = embed_out_dim(multires, d_in)
        widths = [d_e] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(widths)
        n_lin = self.num_layers - 1
        for l in range(n_lin):
            fan_out = widths[l + 1] - widths[0] if (l + 1) in self.skip_in else widths[l + 1]
            lin = nn.Linear(widths[l], fan_out)
            if geometric_init:
                self._geometric_init(lin, l, n_lin, widths, fan_out, bias, inside_outside)
            if weight_norm:
                lin = _weight_norm(lin)
            setattr(self, f"lin{l}", lin)
        self._handle = None

    def _geometric_init(self, lin, l, n_lin, widths, fan_out, bias, inside_outside):
        # Same draws, in the same order, as the reference (fields.py:45-63).
        init = torch.nn.init
        if l == n_lin - 1:
            sign = -1.0 if inside_outside else 1.0
            init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(widths[l]), std=0.0001)
            init.constant_(lin.bias, bias if inside_outside else -bias)
        elif self.multires > 0 and l == 0:
            init.constant_(lin.bias, 0.0)
            init.constant_(lin.weight[:, 3:], 0.0)
            init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(fan_out))
        elif self.multires > 0 and l in self.skip_in:
            init.constant_(lin.bias, 0.0)
            init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(fan_out))
            init.constant_(lin.weight[:, -(widths[0] - 3):], 0.0)
        else:
            init.constant_(lin.bias, 0.0)
            init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(fan_out))

    # -- kernel plumbing -------------------------------------------------------------------------
    def handle(self) -> ops.SdfHandle:
        if self._handle is None:
            skips = [s for s in self.skip_in if 0 < s < self.num_layers - 1]
            if len(skips) > 1:
                raise NotImplementedError("SDFNetwork kernels support at most one skip connection")
            skip = skips[0] if skips else -1
            n_lin = self.num_layers - 1
            self._handle = ops.SdfHandle(self.d_in, self.multires, self.d_hidden, self.n_layers, self.d_out, skip,
                                         self.scale,
                                         lambda: [[_triple(getattr(self, f"lin{l}"))] for l in range(n_lin)])
        return self._handle

    def _apply(self, fn, *a, **k):
        self._handle = None  # parameters are re-created by .to()/.cuda(): rebuild the packed-weight cache
        return super()._apply(fn, *a, **k)

    # -- reference API ---------------------------------------------------------------------------
    def forward(self, inputs):
        out, _ = ops.sdf_eval(self.handle(), inputs, want_normals=False, want_feature=True)
        return out

    def sdf(self, x):
        h = self.handle()
        if not torch.is_grad_enabled() or not (x.requires_grad or any(p.requires_grad for p in h.mlp.params)):
            return ops.sdf_value(h, x)
        out, _ = ops.sdf_eval(h, x, want_normals=False, want_feature=False)
        return out

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        """d sdf / d x as [N, 1, 3] (reference fields.py:97-108), computed analytically in-kernel."""
        _, normals = ops.sdf_eval(self.handle(), x, want_normals=True, want_feature=False)
        return normals.unsqueeze(1)

    def forward_with_gradient(self, x):
        """Fused (forward(x), gradient(x).squeeze(1)): one forward pass instead of the reference's two."""
        return ops.sdf_eval(self.handle(), x, want_normals=True, want_feature=True)

    def forward_split(self, x):
        """(sdf [N,1], feature [N,d_out-1], normals [N,3]) as separate tensors - what render_core consumes; saves the
        slicing of a concatenated output and the zero-filled concatenated cotangent autograd would build for it."""
        return ops.sdf_eval_split(self.handle(), x, want_normals=True, want_feature=True)


class RenderingNetwork(nn.Module):
    """View-dependent colour / depth-feature head (reference fields.py:112-176)."""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True):
        super().__init__()
        self.mode = mode
        self.squeeze_out = squeeze_out
        self.d_feature, self.d_out, self.d_hidden, self.n_layers = d_feature, d_out, d_hidden, n_layers
        self.multires_view = multires_view
        widths = [d_in + d_feature] + [d_hidden] * n_layers + [d_out]
        if multires_view > 0:
            widths[0] += embed_out_dim(multires_view, 3) - 3
        self.num_layers = len(widths)
        for l in range(self.num_layers - 1):
            lin = nn.Linear(widths[l], widths[l + 1])
            if weight_norm:
                lin = _weight_norm(lin)
            setattr(self, f"lin{l}", lin)
        self._in0 = widths[0]
        self._handle = None

    def handle(self) -> ops.RenderNetHandle:
        if self._handle is None:
            n_lin = self.num_layers - 1
            h = ops.RenderNetHandle(self.d_feature, self.mode, self.d_out, self.d_hidden, self.n_layers,
                                    self.multires_view, self.squeeze_out,
                                    lambda: [[_triple(getattr(self, f"lin{l}"))] for l in range(n_lin)])
            if h.in0 != self._in0:
                raise ValueError(f"RenderingNetwork: d_in/mode give input width {self._in0}, kernels expect {h.in0}")
            self._handle = h
        return self._handle

    def _apply(self, fn, *a, **k):
        self._handle = None
        return super()._apply(fn, *a, **k)

    def forward(self, points, normals, view_dirs, feature_vectors):
        return ops.rendernet_eval(self.handle(), points, normals, view_dirs, feature_vectors)


class NeRF(nn.Module):
    """NeRF++ background field (reference fields.py:264-355); only the use_viewdirs=True path exists."""

    def __init__(self, D=8, W=256, d_in=3, d_in_view=3, gen_depth_feats=False, dpt_dim=1, multires=0,
                 multires_view=0, output_ch=4, skips=[4], rgb_dims=3, use_viewdirs=False):
        super().__init__()
        self.D, self.W, self.d_in, self.d_in_view = D, W, d_in, d_in_view
        self.gen_depth_feats, self.dpt_dim = gen_depth_feats, dpt_dim
        self.multires, self.multires_view = multires, multires_view
        self.input_ch = embed_out_dim(multires, d_in) if multires > 0 else 3
        self.input_ch_view = embed_out_dim(multires_view, d_in_view) if multires_view > 0 else 3
        self.skips = list(skips)
        self.use_viewdirs = use_viewdirs
        self.rgb_dims = rgb_dims
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.input_ch, W)] +
            [nn.Linear(W + self.input_ch, W) if i in self.skips else nn.Linear(W, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_view + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, rgb_dims)
            if gen_depth_feats:
                self.dpt_linear = nn.Linear(W // 2, dpt_dim)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self._handle = None

    def handle(self) -> ops.NerfHandle:
        if self._handle is None:
            assert self.use_viewdirs, "NeRF without view directions is not on the reference path (fields.py:355)"
            if len(self.skips) > 1:
                raise NotImplementedError("NeRF kernels support at most one skip connection")
            if self.input_ch != embed_out_dim(self.multires, self.d_in):
                raise NotImplementedError("NeRF with multires=0 requires d_in == 3")
            skip = self.skips[0] if self.skips else -1

            def sources():
                src = [[_triple(l)] for l in self.pts_linears]
                src.append([_triple(self.alpha_linear), _triple(self.feature_linear)])
                src.append([_triple(self.views_linears[0])])
                head = [_triple(self.rgb_linear)]
                if self.gen_depth_feats:
                    head.append(_triple(self.dpt_linear))
                src.append(head)
                return src

            self._handle = ops.NerfHandle(self.D, self.W, self.d_in, self.d_in_view, self.multires,
                                          self.multires_view, skip, self.rgb_dims,
                                          self.dpt_dim if self.gen_depth_feats else 0, sources)
        return self._handle

    def _apply(self, fn, *a, **k):
        self._handle = None
        return super()._apply(fn, *a, **k)

    def forward(self, input_pts, input_views):
        assert self.use_viewdirs
        return ops.nerf_eval(self.handle(), input_pts, input_views)


class SingleVarianceNetwork(nn.Module):
    """One learned scalar; forward returns exp(10 * variance) per row (reference fields.py:358-364).

    Inside ``NeuSRenderer`` the scalar is read by the compositing kernel directly (and its gradient is produced
    by the compositing backward kernel); this forward exists for API compatibility and is a single scalar op.
    """

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)
