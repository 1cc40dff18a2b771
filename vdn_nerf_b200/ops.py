"""torch.autograd bindings of the sm_100a kernels (ctypes -> libvdn_b200.so).

Every function here launches CUDA kernels on torch's current stream with raw device pointers of
torch-allocated tensors; nothing in this module computes on the CPU and nothing falls back to PyTorch
ops.  The backward passes are the hand-derived ones of SURVEY.md Appendix A, not autograd graphs.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import check, int_array, ptr_array


_FLAG = {}          # device index -> (int32 device tensor the tcgen05 kernels raise on a barrier time-out)
_FLAG_DEV = None    # device whose flag is currently installed in the library
_POLL = {}          # device index -> (pinned host int32 tensor, CUDA event of the last asynchronous flag copy)


def _install_flag(dev: int) -> None:
    """The fault flag is a torch tensor on the device the kernels run on (the library owns no device memory)."""
    global _FLAG_DEV
    if dev not in _FLAG:
        _FLAG[dev] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", dev))
    check(_lib.load().vdn_set_fault_flag(ctypes.c_void_p(_FLAG[dev].data_ptr())), "vdn_set_fault_flag")
    _FLAG_DEV = dev


def _stream():
    """Current stream of the current device; keeps the library's fault flag on that device."""
    dev = torch.cuda.current_device()
    if _FLAG_DEV != dev:
        _install_flag(dev)
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _prep(t: Optional[torch.Tensor], device=None) -> Optional[torch.Tensor]:
    """Contiguous fp32 CUDA view of `t` (no copy when it already is one).  Kernels are launched on the CURRENT device's
    current stream, so a tensor living on another device is an error (torch.cuda.set_device / torch.cuda.device first)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.VdnLibraryError("vdn_nerf_b200 kernels need CUDA tensors (there is no CPU path); got " + str(t.device))
    if t.device.index != torch.cuda.current_device():
        raise _lib.VdnLibraryError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                                   "the kernels launch on the current device's stream; wrap the call in "
                                   "`with torch.cuda.device(tensor.device):`")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def launch_count() -> int:
    return int(_lib.load().vdn_launch_count())


_PRECISIONS = {"fp32": 0, "tf32": 1}


def set_precision(mode: str) -> None:
    """'fp32': exact FFMA kernels (parity <= 1e-5); 'tf32': tcgen05 tensor-core kernels (parity <= 2e-3)."""
    if mode not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    check(_lib.load().vdn_set_mode(_PRECISIONS[mode]), "vdn_set_mode")


def get_precision() -> str:
    m = _lib.load().vdn_get_mode()
    return [k for k, v in _PRECISIONS.items() if v == m][0]


def set_chain(on: bool) -> None:
    """Tensor-core mode: run the training passes as fused layer chains (default) or layer by layer (diagnostics)."""
    check(_lib.load().vdn_set_chain(1 if on else 0), "vdn_set_chain")


def get_chain() -> bool:
    return bool(_lib.load().vdn_get_chain())


def tc_fault() -> int:
    """Non-zero if a tensor-core kernel hit a barrier time-out on the current device since the flag was last reset
    (synchronises)."""
    dev = torch.cuda.current_device()
    if dev not in _FLAG:
        return 0
    return int(_FLAG[dev].item())


def reset_fault() -> None:
    dev = torch.cuda.current_device()
    if dev in _FLAG:
        _FLAG[dev].zero_()


def poll_fault() -> None:
    """Asynchronous fault check for the training loop: raises if the flag copy enqueued by an EARLIER call shows a
    barrier time-out, then enqueues a fresh device -> pinned-host copy of the flag on the current stream.  Never
    synchronises; a fault is therefore reported one call late (call it once per step).  No-op during graph capture."""
    if not torch.cuda.is_available() or torch.cuda.is_current_stream_capturing():
        return
    dev = torch.cuda.current_device()
    if dev not in _FLAG:
        return
    host, ev = _POLL.get(dev, (None, None))
    if host is not None and ev.query():
        if int(host[0]) != 0:
            host[0] = 0
            _FLAG[dev].zero_()
            raise _lib.VdnLibraryError("a tcgen05 kernel timed out on a barrier (device fault flag raised): the results of "
                                       "the previous step(s) are not valid")
    elif host is not None:
        return                      # the previous copy is still in flight: do not pile up
    if host is None:
        host = torch.zeros(1, dtype=torch.int32).pin_memory()
        ev = torch.cuda.Event()
    host.copy_(_FLAG[dev], non_blocking=True)
    ev.record()
    _POLL[dev] = (host, ev)


# ----------------------------------------------------------------------------------------------------
# Packed parameters
# ----------------------------------------------------------------------------------------------------
_FORCE_REPACK = False
_PARAM_EPOCH = 0        # bumped by in-place parameter updates that bypass autograd's version counters (driver.FusedAdam)


def bump_param_epoch() -> None:
    """Invalidate every packed-weight cache: call after writing parameters through raw pointers."""
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1


_REPACK_SERIAL = 0      # one value per force_repack(True) window (= per graph capture)


def force_repack(on: bool) -> None:
    """While on, PackedMLP.packed() does not trust the cache it built before: the first call of every network inside the
    window re-materialises the weights, later calls with unchanged parameters reuse that buffer.  Used while a training
    step is captured into a CUDA graph, so that ONE pack launch per network is part of the graph and every replay sees
    the parameters the optimiser has written since."""
    global _FORCE_REPACK, _REPACK_SERIAL
    _FORCE_REPACK = bool(on)
    if on:
        _REPACK_SERIAL += 1


# Gradient arena (data-parallel training, dist.FlatGradAllReduce): when a parameter has a slot in a flat all-reduce
# buffer, the weight-norm backward writes its gradient THERE, autograd adopts that tensor as `.grad`, and the pack /
# unpack copies around the collective disappear.  A slot is handed out once per step and only while `.grad` is None, so
# a network evaluated twice in one graph, or gradients accumulated over several backward calls, stay correct.
_GRAD_ARENA = {}
_ARENA_USED = set()


def set_grad_arena(params, views) -> None:
    _GRAD_ARENA.clear()
    _ARENA_USED.clear()
    for p, v in zip(params, views):
        _GRAD_ARENA[id(p)] = v


def reset_grad_arena_use() -> None:
    _ARENA_USED.clear()


def _grad_buffer(p: torch.Tensor) -> torch.Tensor:
    v = _GRAD_ARENA.get(id(p))
    if v is None or id(p) in _ARENA_USED or p.grad is not None or v.device != p.device or v.shape != p.shape:
        return torch.empty_like(p)
    _ARENA_USED.add(id(p))
    return v.view_as(v)         # a fresh alias: autograd may adopt it as .grad without cloning


class PackedMLP:
    """Effective (weight-normed, padded, transposed) weights of one network in one device buffer.

    `sources[l]` is a list of one or two `(weight, g_or_None, bias_or_None)` parameter triples that are
    stacked by rows to form packed layer l.  The buffer is rebuilt (one kernel launch) only when a
    parameter's storage or version counter changes, i.e. once per optimiser step.
    """

    def __init__(self, in_dims: Sequence[int], out_dims: Sequence[int], sources, rot: Optional[Sequence[int]] = None,
                 orot: Optional[Sequence[int]] = None):
        self.in_dims = [int(v) for v in in_dims]
        self.out_dims = [int(v) for v in out_dims]
        self.L = len(self.in_dims)
        self.sources = sources
        assert len(sources) == self.L
        self._in = int_array(self.in_dims)
        self._out = int_array(self.out_dims)
        rows = []
        self.params: List[torch.Tensor] = []      # flat list in (layer, source, [weight, g, bias]) order
        for l, srcs in enumerate(sources):
            assert 1 <= len(srcs) <= 2
            tot = 0
            for (w, g, b) in srcs:
                assert w.shape[1] == self.in_dims[l], (l, tuple(w.shape), self.in_dims[l])
                tot += w.shape[0]
                self.params.append(w)
                if g is not None:
                    self.params.append(g)
                if b is not None:
                    self.params.append(b)
            assert tot == self.out_dims[l], (l, tot, self.out_dims[l])
            rows += [srcs[0][0].shape[0], srcs[1][0].shape[0] if len(srcs) > 1 else 0]
        self._rows = int_array(rows)
        self._rot = int_array(list(rot) if rot is not None else [0] * self.L)
        # output rotation of the fp16 tile images (stacked [scalar ; features] heads present their features first)
        self._orot = int_array(list(orot) if orot is not None else [0] * self.L)
        lib = _lib.load()
        L = self.L
        self.off_w = (ctypes.c_longlong * L)()
        self.off_wt = (ctypes.c_longlong * L)()
        self.off_b = (ctypes.c_longlong * L)()
        self.total = int(lib.vdn_mlp_layout(L, self._in, self._out, self.off_w, self.off_wt, self.off_b))
        if self.total <= 0:
            raise _lib.VdnLibraryError("invalid MLP layout")
        self._key = None
        self._window_key = None
        self._packed = None

    def _src_ptrs(self, pick):
        ptrs = []
        for srcs in self.sources:
            for s in range(2):
                t = pick(srcs[s]) if s < len(srcs) else None
                ptrs.append(t.data_ptr() if t is not None else 0)
        return ptr_array(ptrs)

    def packed(self) -> torch.Tensor:
        key = (_PARAM_EPOCH,) + tuple((p.data_ptr(), p._version) for p in self.params)
        if _FORCE_REPACK and self._packed is not None and self._window_key == (_REPACK_SERIAL, key):
            return self._packed           # packed earlier in this capture window, parameters untouched since
        if key != self._key or self._packed is None or _FORCE_REPACK:
            dev = self.params[0].device
            for p in self.params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.VdnLibraryError("network parameters must be contiguous fp32 CUDA tensors; "
                                               "call .to('cuda') on the module")
            buf = torch.empty(self.total, device=dev, dtype=torch.float32)
            lib = _lib.load()
            check(lib.vdn_mlp_pack(self.L, self._in, self._out, self._src_ptrs(lambda s: s[0]),
                                   self._src_ptrs(lambda s: s[1]), self._src_ptrs(lambda s: s[2]), self._rows,
                                   self._rot, self._orot, _p(buf), _stream()), "vdn_mlp_pack")
            self._packed = buf
            self._window_key = (_REPACK_SERIAL, key) if _FORCE_REPACK else None
            # a pack recorded during graph capture has not run: never let an eager call trust that buffer
            self._key = None if (_FORCE_REPACK or torch.cuda.is_current_stream_capturing()) else key
        return self._packed

    def invalidate(self) -> None:
        """Forget the cached packed weights (parameters changed in a way the version counters do not see, e.g. through
        `.data` or by a captured optimiser step)."""
        self._key = None
        self._window_key = None

    def unpack_grads(self, dpacked: torch.Tensor, needs: Sequence[bool]) -> List[Optional[torch.Tensor]]:
        """Packed gradient -> gradients of self.params (same order); weight-norm backward included."""
        grads: List[Optional[torch.Tensor]] = []
        dv, dg, db = [], [], []
        i = 0
        for srcs in self.sources:
            for s in range(2):
                if s >= len(srcs):
                    dv.append(0); dg.append(0); db.append(0)
                    continue
                w, g, b = srcs[s]
                gw = _grad_buffer(w) if needs[i] else None
                grads.append(gw); i += 1
                gg = None
                if g is not None:
                    gg = _grad_buffer(g) if needs[i] else None
                    grads.append(gg); i += 1
                gb = None
                if b is not None:
                    gb = _grad_buffer(b) if needs[i] else None
                    grads.append(gb); i += 1
                dv.append(gw.data_ptr() if gw is not None else 0)
                dg.append(gg.data_ptr() if gg is not None else 0)
                db.append(gb.data_ptr() if gb is not None else 0)
        lib = _lib.load()
        check(lib.vdn_mlp_unpack_grads(self.L, self._in, self._out, self._src_ptrs(lambda s: s[0]),
                                       self._src_ptrs(lambda s: s[1]), self._rows, self._rot, _p(dpacked), ptr_array(dv),
                                       ptr_array(dg), ptr_array(db), _stream()), "vdn_mlp_unpack_grads")
        return grads

    def weight_view(self, packed: torch.Tensor, l: int) -> torch.Tensor:
        """[out_dim, in_dim] view of the effective weight of layer l (tests / diagnostics)."""
        in_ld = (self.in_dims[l] + 15) // 16 * 16
        out_ld = (self.out_dims[l] + 15) // 16 * 16
        w = packed[self.off_w[l]: self.off_w[l] + out_ld * in_ld].view(out_ld, in_ld)
        return w[: self.out_dims[l], : self.in_dims[l]]


# ----------------------------------------------------------------------------------------------------
# Embedder
# ----------------------------------------------------------------------------------------------------
class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, multires):
        shape = x.shape
        d = shape[-1]
        x2 = _prep(x.reshape(-1, d))
        N = x2.shape[0]
        out = torch.empty(N, d * (1 + 2 * multires), device=x2.device, dtype=torch.float32)
        check(_lib.load().vdn_embed_fwd(_p(x2), N, d, multires, _p(out), _stream()), "vdn_embed_fwd")
        ctx.save_for_backward(x2)
        ctx.multires = multires
        ctx.shape = shape
        return out.reshape(*shape[:-1], out.shape[-1])

    @staticmethod
    def backward(ctx, d_out):
        (x2,) = ctx.saved_tensors
        N, d = x2.shape
        d_out = _prep(d_out.reshape(N, -1))
        d_x = torch.empty_like(x2)
        check(_lib.load().vdn_embed_bwd(_p(x2), N, d, ctx.multires, _p(d_out), _p(d_x), _stream()), "vdn_embed_bwd")
        return d_x.reshape(ctx.shape), None


def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    if multires <= 0:
        return x
    return _EmbedFn.apply(x, int(multires))


# ----------------------------------------------------------------------------------------------------
# SDF network
# ----------------------------------------------------------------------------------------------------
class SdfHandle:
    """Static description of one SDFNetwork instance (config ints + packed-weight cache)."""

    def __init__(self, d_in, multires, d_hidden, n_layers, d_out, skip, scale, sources_fn):
        self.cfg_list = [d_in, multires, d_hidden, n_layers, d_out, skip]
        self.cfg = int_array(self.cfg_list)
        self.scale = float(scale)
        self.d_in, self.d_out = d_in, d_out
        lib = _lib.load()
        ind = (ctypes.c_int * 16)()
        outd = (ctypes.c_int * 16)()
        L = lib.vdn_sdf_layer_dims(self.cfg, ind, outd)
        if L <= 0:
            raise _lib.VdnLibraryError(f"unsupported SDFNetwork configuration {self.cfg_list}")
        self.L = L
        orot = (ctypes.c_int * 16)()
        lib.vdn_sdf_layer_orot(self.cfg, orot)
        self.mlp = PackedMLP(list(ind[:L]), list(outd[:L]), sources_fn(), orot=list(orot[:L]))


def sdf_value(h: SdfHandle, x: torch.Tensor) -> torch.Tensor:
    """sdf only, no autograd (up-sampling and grid queries): [N, 1]."""
    lib = _lib.load()
    x = _prep(x)
    N = x.shape[0]
    packed = h.mlp.packed()
    out = torch.empty(N, 1, device=x.device, dtype=torch.float32)
    blob = torch.empty(int(lib.vdn_sdf_blob_floats(h.cfg, N, 0)), device=x.device, dtype=torch.float32)
    check(lib.vdn_sdf_forward(h.cfg, h.scale, _p(packed), _p(x), N, _p(out), 1, None, 0, _p(blob), 0, _stream()),
          "vdn_sdf_forward")
    return out


class _SdfFn(torch.autograd.Function):
    """(x, params...) -> (sdf [N,1], feature [N, d_out-1] or None, normals [N, d_in] or None).

    The value and the feature are separate outputs (the C ABI takes separate pointers and leading dimensions), so
    autograd hands their cotangents over as they are instead of scattering both into a zero-filled [N, d_out]."""

    @staticmethod
    def forward(ctx, h: SdfHandle, x, want_normals: bool, want_feature: bool, *params):
        lib = _lib.load()
        x = _prep(x)
        N = x.shape[0]
        packed = h.mlp.packed()
        need_grad = any(ctx.needs_input_grad[4:]) or ctx.needs_input_grad[1]
        save = 1 if (need_grad or want_normals) else 0
        dev = x.device
        sdf = torch.empty(N, 1, device=dev, dtype=torch.float32)
        feat = torch.empty(N, h.d_out - 1, device=dev, dtype=torch.float32) if want_feature else None
        blob = torch.empty(int(lib.vdn_sdf_blob_floats(h.cfg, N, save)), device=dev, dtype=torch.float32)
        check(lib.vdn_sdf_forward(h.cfg, h.scale, _p(packed), _p(x), N, _p(sdf), 1, _p(feat), h.d_out - 1, _p(blob),
                                  save, _stream()), "vdn_sdf_forward")
        normals = None
        blobg = None
        if want_normals:
            normals = torch.empty(N, h.d_in, device=dev, dtype=torch.float32)
            blobg = torch.empty(int(lib.vdn_sdf_blobg_floats(h.cfg, N)), device=dev, dtype=torch.float32)
            check(lib.vdn_sdf_normals(h.cfg, h.scale, _p(packed), _p(x), N, _p(blob), _p(blobg), _p(normals),
                                      _stream()), "vdn_sdf_normals")
        ctx.h = h
        ctx.want_normals = want_normals
        ctx.want_feature = want_feature
        if need_grad:
            ctx.save_for_backward(x, packed, blob, blobg)
        return sdf, feat, normals

    @staticmethod
    def backward(ctx, d_sdf, d_feat, d_normals):
        lib = _lib.load()
        h = ctx.h
        x, packed, blob, blobg = ctx.saved_tensors
        N = x.shape[0]
        dev = x.device
        d_sdf = _prep(d_sdf) if d_sdf is not None else None
        ldf = h.d_out - 1
        if d_feat is None or not ctx.want_feature:
            d_feat = None
        elif d_feat.is_cuda and d_feat.dtype == torch.float32 and d_feat.dim() == 2 and d_feat.stride(1) == 1:
            ldf = d_feat.stride(0)        # e.g. a column slice of the colour network's input cotangent: read in place
        else:
            d_feat = _prep(d_feat)
        d_normals = _prep(d_normals) if (d_normals is not None and ctx.want_normals) else None
        dpacked = torch.zeros(h.mlp.total, device=dev, dtype=torch.float32)
        ws = torch.empty(int(lib.vdn_sdf_bwd_ws_floats(h.cfg, N)), device=dev, dtype=torch.float32)
        d_x = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        check(lib.vdn_sdf_backward(h.cfg, h.scale, _p(packed), _p(x), N, _p(blob), _p(blobg), _p(d_sdf), 1, _p(d_feat),
                                   ldf, _p(d_normals), _p(dpacked), _p(d_x), _p(ws), _stream()),
              "vdn_sdf_backward")
        grads = h.mlp.unpack_grads(dpacked, ctx.needs_input_grad[4:])
        return (None, d_x, None, None, *grads)


def sdf_eval_split(h: SdfHandle, x: torch.Tensor, want_normals: bool, want_feature: bool = True):
    """(sdf [N,1], feature [N, d_out-1] | None, normals [N, d_in] | None)."""
    return _SdfFn.apply(h, x, want_normals, want_feature, *h.mlp.params)


def sdf_eval(h: SdfHandle, x: torch.Tensor, want_normals: bool, want_feature: bool = True):
    """Reference-shaped result: (out [N, d_out] = [sdf | feature] (or [N,1]), normals | None)."""
    sdf, feat, normals = sdf_eval_split(h, x, want_normals, want_feature)
    out = torch.cat([sdf, feat], dim=1) if want_feature else sdf
    return out, normals


# ----------------------------------------------------------------------------------------------------
# Rendering network (colour / depth-feature heads)
# ----------------------------------------------------------------------------------------------------
_MODES = {"idr": 0, "no_view_dir": 1, "no_normal": 2}


class RenderNetHandle:
    def __init__(self, d_feature, mode, d_out, d_hidden, n_layers, multires_view, squeeze_out, sources_fn):
        if mode not in _MODES:
            raise ValueError(f"unknown RenderingNetwork mode {mode!r}")
        self.mode = _MODES[mode]
        self.cfg_list = [d_feature, self.mode, d_out, d_hidden, n_layers, multires_view, int(bool(squeeze_out))]
        self.cfg = int_array(self.cfg_list)
        self.d_feature, self.d_out, self.multires_view = d_feature, d_out, multires_view
        lib = _lib.load()
        ind = (ctypes.c_int * 16)()
        outd = (ctypes.c_int * 16)()
        L = lib.vdn_rendernet_layer_dims(self.cfg, ind, outd)
        if L <= 0:
            raise _lib.VdnLibraryError(f"unsupported RenderingNetwork configuration {self.cfg_list}")
        self.L = L
        self.in0 = int(ind[0])
        self.ld_in = (self.in0 + 15) // 16 * 16
        rot = (ctypes.c_int * 16)()
        lib.vdn_rendernet_layer_rot(self.cfg, rot)
        self.mlp = PackedMLP(list(ind[:L]), list(outd[:L]), sources_fn(), rot=list(rot[:L]))
        # column ranges inside the (rotated) input row [feature | points | view embedding | normals]
        nview = 3 * (1 + 2 * multires_view) if self.mode != 1 else 0
        F = d_feature
        self.col_feat = (0, F)
        self.col_pts = (F, F + 3)
        self.col_view = (F + 3, F + 3 + nview) if self.mode != 1 else None
        self.col_nrm = (F + 3 + nview, F + 3 + nview + 3) if self.mode != 2 else None


class _RenderNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h: RenderNetHandle, points, normals, view_dirs, feats, *params):
        lib = _lib.load()
        points = _prep(points)
        N = points.shape[0]
        dev = points.device
        normals = _prep(normals) if normals is not None else None
        view_dirs = _prep(view_dirs) if view_dirs is not None else None
        if feats.dtype != torch.float32:
            feats = feats.float()
        if feats.stride(-1) != 1 or feats.stride(0) < feats.shape[1]:
            feats = feats.contiguous()
        ldf = feats.stride(0)
        packed = h.mlp.packed()
        out = torch.empty(N, h.d_out, device=dev, dtype=torch.float32)
        blob = torch.empty(int(lib.vdn_rendernet_blob_floats(h.cfg, N)), device=dev, dtype=torch.float32)
        check(lib.vdn_rendernet_forward(h.cfg, _p(packed), _p(points), _p(normals), _p(view_dirs), _p(feats), ldf, N,
                                        _p(out), _p(blob), _stream()), "vdn_rendernet_forward")
        ctx.h = h
        ctx.save_for_backward(packed, blob, out, view_dirs)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        h = ctx.h
        packed, blob, out, view_dirs = ctx.saved_tensors
        N = out.shape[0]
        dev = out.device
        d_out = _prep(d_out)
        need_in = any(ctx.needs_input_grad[1:5])
        dpacked = torch.zeros(h.mlp.total, device=dev, dtype=torch.float32)
        ws = torch.empty(int(lib.vdn_rendernet_bwd_ws_floats(h.cfg, N)), device=dev, dtype=torch.float32)
        d_cin = torch.empty(N, h.ld_in, device=dev, dtype=torch.float32) if need_in else None
        check(lib.vdn_rendernet_backward(h.cfg, _p(packed), N, _p(blob), _p(out), _p(d_out), _p(dpacked), _p(d_cin),
                                         _p(ws), _stream()), "vdn_rendernet_backward")
        grads = h.mlp.unpack_grads(dpacked, ctx.needs_input_grad[5:])
        d_points = d_normals = d_view = d_feat = None
        if need_in:
            if ctx.needs_input_grad[1]:
                d_points = d_cin[:, h.col_pts[0]: h.col_pts[1]]
            if ctx.needs_input_grad[2] and h.col_nrm is not None:
                d_normals = d_cin[:, h.col_nrm[0]: h.col_nrm[1]]
            if ctx.needs_input_grad[3] and h.col_view is not None:
                de = d_cin[:, h.col_view[0]: h.col_view[1]].contiguous()
                d_view = torch.empty(N, 3, device=dev, dtype=torch.float32)
                check(lib.vdn_embed_bwd(_p(view_dirs), N, 3, h.multires_view, _p(de), _p(d_view), _stream()),
                      "vdn_embed_bwd")
            if ctx.needs_input_grad[4]:
                d_feat = d_cin[:, h.col_feat[0]: h.col_feat[1]]
        return (None, d_points, d_normals, d_view, d_feat, *grads)


def rendernet_eval(h: RenderNetHandle, points, normals, view_dirs, feats):
    return _RenderNetFn.apply(h, points, normals, view_dirs, feats, *h.mlp.params)


# ----------------------------------------------------------------------------------------------------
# NeRF background field
# ----------------------------------------------------------------------------------------------------
class NerfHandle:
    def __init__(self, D, W, d_in, d_in_view, multires, multires_view, skip, rgb_dims, dpt_dim, sources_fn):
        self.cfg_list = [D, W, d_in, d_in_view, multires, multires_view, skip, rgb_dims, dpt_dim]
        self.cfg = int_array(self.cfg_list)
        self.d_in, self.rgb_dims, self.dpt_dim = d_in, rgb_dims, dpt_dim
        lib = _lib.load()
        ind = (ctypes.c_int * 16)()
        outd = (ctypes.c_int * 16)()
        L = lib.vdn_nerf_layer_dims(self.cfg, ind, outd)
        if L <= 0:
            raise _lib.VdnLibraryError(f"unsupported NeRF configuration {self.cfg_list}")
        self.L = L
        rot = [0] * L
        if skip >= 0:
            rot[skip + 1] = int(ind[0])      # skip layer input is stored as [hidden | embedding]
        orot = (ctypes.c_int * 16)()
        lib.vdn_nerf_layer_orot(self.cfg, orot)
        self.mlp = PackedMLP(list(ind[:L]), list(outd[:L]), sources_fn(), rot=rot, orot=list(orot[:L]))


class _NerfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h: NerfHandle, pts, views, *params):
        lib = _lib.load()
        pts = _prep(pts)
        views = _prep(views)
        N = pts.shape[0]
        dev = pts.device
        packed = h.mlp.packed()
        sigma = torch.empty(N, 1, device=dev, dtype=torch.float32)
        rgb = torch.empty(N, h.rgb_dims, device=dev, dtype=torch.float32)
        dpt = torch.empty(N, h.dpt_dim, device=dev, dtype=torch.float32) if h.dpt_dim > 0 else None
        blob = torch.empty(int(lib.vdn_nerf_blob_floats(h.cfg, N)), device=dev, dtype=torch.float32)
        check(lib.vdn_nerf_forward(h.cfg, _p(packed), _p(pts), _p(views), N, _p(sigma), _p(rgb), _p(dpt), _p(blob),
                                   _stream()), "vdn_nerf_forward")
        ctx.h = h
        ctx.save_for_backward(packed, blob, pts, views)
        if dpt is None:
            return sigma, rgb, None
        return sigma, rgb, dpt

    @staticmethod
    def backward(ctx, d_sigma, d_rgb, d_dpt):
        lib = _lib.load()
        h = ctx.h
        packed, blob, pts, views = ctx.saved_tensors
        N = pts.shape[0]
        dev = pts.device
        d_sigma = _prep(d_sigma) if d_sigma is not None else None
        d_rgb = _prep(d_rgb) if d_rgb is not None else None
        d_dpt = _prep(d_dpt) if (d_dpt is not None and h.dpt_dim > 0) else None
        dpacked = torch.zeros(h.mlp.total, device=dev, dtype=torch.float32)
        ws = torch.empty(int(lib.vdn_nerf_bwd_ws_floats(h.cfg, N)), device=dev, dtype=torch.float32)
        d_pts = torch.empty_like(pts) if ctx.needs_input_grad[1] else None
        d_views = torch.empty_like(views) if ctx.needs_input_grad[2] else None
        check(lib.vdn_nerf_backward(h.cfg, _p(packed), _p(pts), _p(views), N, _p(blob), _p(d_sigma), _p(d_rgb),
                                    _p(d_dpt), _p(dpacked), _p(d_pts), _p(d_views), _p(ws), _stream()),
              "vdn_nerf_backward")
        grads = h.mlp.unpack_grads(dpacked, ctx.needs_input_grad[3:])
        return (None, d_pts, d_views, *grads)


def nerf_eval(h: NerfHandle, pts, views):
    return _NerfFn.apply(h, pts, views, *h.mlp.params)


# ----------------------------------------------------------------------------------------------------
# Per-ray kernels
# ----------------------------------------------------------------------------------------------------
def ray_points(o, d, z):
    o, d, z = _prep(o), _prep(d), _prep(z)
    B, n = z.shape
    pts = torch.empty(B * n, 3, device=z.device, dtype=torch.float32)
    check(_lib.load().vdn_ray_points(_p(o), _p(d), _p(z), B, n, _p(pts), _stream()), "vdn_ray_points")
    return pts


def upsample_step(o, d, z_in, sdf_prev, sdf_new, perm_prev, inv_s, n_imp, want_inds=False, want_sdf=True):
    """One iteration of the hierarchical resampling loop; see include/vdn_b200.h."""
    o, d, z_in = _prep(o), _prep(d), _prep(z_in)
    B, n = z_in.shape
    dev = z_in.device
    sdf_prev = _prep(sdf_prev)
    n_prev = sdf_prev.shape[1]
    n_new_prev = 0
    if perm_prev is not None:
        sdf_new = _prep(sdf_new)
        n_new_prev = sdf_new.shape[1]
    z_out = torch.empty(B, n + n_imp, device=dev, dtype=torch.float32)
    sdf_out = torch.empty(B, n, device=dev, dtype=torch.float32) if want_sdf else None
    perm_out = torch.empty(B, n + n_imp, device=dev, dtype=torch.uint8)
    new_z = torch.empty(B, n_imp, device=dev, dtype=torch.float32)
    new_pts = torch.empty(B * n_imp, 3, device=dev, dtype=torch.float32)
    inds = torch.empty(B, n_imp, device=dev, dtype=torch.int64) if want_inds else None
    check(_lib.load().vdn_upsample_step(_p(o), _p(d), _p(z_in), n, _p(sdf_prev), n_prev,
                                        _p(sdf_new) if perm_prev is not None else None, n_new_prev, _p(perm_prev),
                                        float(inv_s), n_imp, B, _p(z_out), _p(sdf_out), _p(perm_out), _p(new_z),
                                        _p(new_pts), _p(inds), _stream()), "vdn_upsample_step")
    return z_out, sdf_out, perm_out, new_z, new_pts, inds


def merge_sorted(za, zb):
    """Stable merge of sorted za [B,n] with zb [B,m]: (z [B,n+m], source index uint8 [B,n+m])."""
    za, zb = _prep(za), _prep(zb)
    B, n = za.shape
    m = zb.shape[1]
    z = torch.empty(B, n + m, device=za.device, dtype=torch.float32)
    perm = torch.empty(B, n + m, device=za.device, dtype=torch.uint8)
    check(_lib.load().vdn_merge_sorted(_p(za), n, _p(zb), m, B, _p(z), _p(perm), _stream()), "vdn_merge_sorted")
    return z, perm


def fine_prep(o, d, z, sample_dist):
    o, d, z = _prep(o), _prep(d), _prep(z)
    B, S = z.shape
    dev = z.device
    dists = torch.empty(B, S, device=dev, dtype=torch.float32)
    mid = torch.empty(B, S, device=dev, dtype=torch.float32)
    pts = torch.empty(B * S, 3, device=dev, dtype=torch.float32)
    check(_lib.load().vdn_fine_prep(_p(o), _p(d), _p(z), float(sample_dist), B, S, _p(dists), _p(mid), _p(pts),
                                    _stream()), "vdn_fine_prep")
    return dists, mid, pts


def bg_prep(o, d, z_fine, z_outside, sample_dist):
    o, d, z_fine, z_outside = _prep(o), _prep(d), _prep(z_fine), _prep(z_outside)
    B, S = z_fine.shape
    NO = z_outside.shape[1]
    dev = z_fine.device
    nt = S + NO
    dists = torch.empty(B, nt, device=dev, dtype=torch.float32)
    mid = torch.empty(B, nt, device=dev, dtype=torch.float32)
    pts4 = torch.empty(B * nt, 4, device=dev, dtype=torch.float32)
    check(_lib.load().vdn_bg_prep(_p(o), _p(d), _p(z_fine), S, _p(z_outside), NO, float(sample_dist), B, _p(dists),
                                  _p(mid), _p(pts4), _stream()), "vdn_bg_prep")
    return dists, mid, pts4


class _CompositeFn(torch.autograd.Function):
    """Alpha from the sigmoid CDF, background blend, transmittance scan, compositing, Eikonal sums."""

    @staticmethod
    def forward(ctx, o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance, bg_rgb,
                cos_anneal):
        lib = _lib.load()
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            # only reachable with n_importance == 0 and differentiable near / far (the up-sampled z are detached,
            # renderer.py:190, 368): the closed-form backward does not return these two cotangents
            raise NotImplementedError("gradients with respect to the sample depths (mid_z / dists) are not implemented; "
                                      "detach near / far or use n_importance > 0")
        o, d, mid_z, dists = _prep(o), _prep(d), _prep(mid_z), _prep(dists)
        B, S = mid_z.shape
        dev = mid_z.device
        sdf, nrm, col = _prep(sdf), _prep(nrm), _prep(col)
        feat = _prep(feat) if feat is not None else None
        F = feat.shape[-1] if feat is not None else 0
        NB = 0
        if sigma_bg is not None:
            sigma_bg, rgb_bg, dists_bg = _prep(sigma_bg), _prep(rgb_bg), _prep(dists_bg)
            NB = sigma_bg.numel() // B
            feat_bg = _prep(feat_bg) if (feat_bg is not None and F > 0) else None
            if F > 0 and feat_bg is None:
                raise ValueError("depth features given for the fine samples but not for the background")
        variance = _prep(variance.reshape(1))
        bg = _prep(bg_rgb.reshape(-1)) if bg_rgb is not None else None
        NW = NB if NB > 0 else S
        weights = torch.empty(B, NW, device=dev, dtype=torch.float32)
        cdf = torch.empty(B, S, device=dev, dtype=torch.float32)
        inside = torch.empty(B, S, device=dev, dtype=torch.float32)
        color = torch.empty(B, 3, device=dev, dtype=torch.float32)
        dfeat = torch.empty(B, F, device=dev, dtype=torch.float32) if F > 0 else None
        en = torch.empty(B, device=dev, dtype=torch.float32)
        ed = torch.empty(B, device=dev, dtype=torch.float32)
        check(lib.vdn_composite_fwd(B, S, NB, F, _p(o), _p(d), _p(mid_z), _p(dists), _p(sdf), _p(nrm), _p(col),
                                    _p(feat), _p(sigma_bg), _p(rgb_bg), _p(feat_bg), _p(dists_bg), _p(variance),
                                    _p(bg), float(cos_anneal), _p(weights), _p(cdf), _p(inside), _p(color), _p(dfeat),
                                    _p(en), _p(ed), _stream()), "vdn_composite_fwd")
        ctx.dims = (B, S, NB, F)
        ctx.cos_anneal = float(cos_anneal)
        ctx.has_bg_rgb = bg is not None
        ctx.save_for_backward(o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance, bg)
        ctx.mark_non_differentiable(inside, ed)
        if dfeat is None:
            return weights, cdf, inside, color, None, en, ed
        return weights, cdf, inside, color, dfeat, en, ed

    @staticmethod
    def backward(ctx, d_weights, d_cdf, _d_inside, d_color, d_dfeat, d_en, _d_ed):
        lib = _lib.load()
        (o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance, bg) = ctx.saved_tensors
        B, S, NB, F = ctx.dims
        dev = mid_z.device
        z = lambda t: _prep(t) if t is not None else None
        d_weights, d_cdf, d_dfeat, d_en = z(d_weights), z(d_cdf), z(d_dfeat), z(d_en)
        d_color = _prep(d_color) if d_color is not None else torch.zeros(B, 3, device=dev, dtype=torch.float32)
        g_sdf = torch.empty_like(sdf)
        g_nrm = torch.empty_like(nrm)
        g_col = torch.empty_like(col)
        g_feat = torch.empty_like(feat) if F > 0 else None
        g_sig = torch.empty_like(sigma_bg) if NB > 0 else None
        g_rgb = torch.empty_like(rgb_bg) if NB > 0 else None
        g_fbg = torch.empty_like(feat_bg) if (NB > 0 and F > 0) else None
        need_dists_bg = NB > 0 and ctx.needs_input_grad[11]
        g_dbg = torch.empty_like(dists_bg) if need_dists_bg else None
        g_var = torch.empty(B, device=dev, dtype=torch.float32)
        g_dirs = torch.empty(B, 3, device=dev, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        check(lib.vdn_composite_bwd(B, S, NB, F, _p(o), _p(d), _p(mid_z), _p(dists), _p(sdf), _p(nrm), _p(col),
                                    _p(feat), _p(sigma_bg), _p(rgb_bg), _p(feat_bg), _p(dists_bg), _p(variance),
                                    _p(bg), ctx.cos_anneal, _p(d_color), _p(d_weights), _p(d_cdf), _p(d_dfeat),
                                    _p(d_en), _p(g_sdf), _p(g_nrm), _p(g_col), _p(g_feat), _p(g_sig), _p(g_rgb),
                                    _p(g_fbg), _p(g_dbg), _p(g_var), _p(g_dirs), _stream()), "vdn_composite_bwd")
        g_variance = g_var.sum().reshape(()) if ctx.needs_input_grad[12] else None
        return (None, g_dirs, None, None, g_sdf, g_nrm, g_col, g_feat, g_sig, g_rgb, g_fbg, g_dbg, g_variance, None,
                None)


def composite(o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance, bg_rgb,
              cos_anneal):
    return _CompositeFn.apply(o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg, variance,
                              bg_rgb, cos_anneal)


def grid_sdf(h: SdfHandle, xs, ys, zs, i0, i1, out_mul, u_slab):
    """u_slab[(i1-i0), ny, nz] = out_mul * sdf(lattice points); all tensors on the GPU."""
    lib = _lib.load()
    ny, nz = ys.numel(), zs.numel()
    count = (i1 - i0) * ny * nz
    dev = u_slab.device
    packed = h.mlp.packed()
    pts = torch.empty(count * 3, device=dev, dtype=torch.float32)
    blob = torch.empty(int(lib.vdn_sdf_blob_floats(h.cfg, count, 0)), device=dev, dtype=torch.float32)
    check(lib.vdn_grid_sdf(h.cfg, h.scale, _p(packed), _p(xs), _p(ys), _p(zs), ny, nz, i0, i1, float(out_mul),
                           _p(u_slab), _p(pts), _p(blob), _stream()), "vdn_grid_sdf")


# ----------------------------------------------------------------------------------------------------
# Marching cubes on the device
# ----------------------------------------------------------------------------------------------------
_MC_TABLES = {}


def _mc_tables(dev):
    key = str(dev)
    if key not in _MC_TABLES:
        from .mcubes_table import EDGE_AXIS, TRI_COUNT, TRI_TABLE
        _MC_TABLES[key] = (torch.from_numpy(TRI_COUNT.copy()).to(dev), torch.from_numpy(TRI_TABLE.reshape(-1).copy()).to(dev),
                           torch.tensor([a for a, _ in EDGE_AXIS], dtype=torch.int32, device=dev),
                           torch.tensor([ax for _, ax in EDGE_AXIS], dtype=torch.int32, device=dev))
    return _MC_TABLES[key]


def marching_cubes(u: torch.Tensor, threshold: float = 0.0):
    """Isosurface of the device-resident field u[nx, ny, nz] at `threshold` (inside = u > threshold), like
    `mcubes.marching_cubes` (reference renderer.py:36): (vertices [V,3] float32 in grid-index coordinates, triangles [F,3]
    int64, outward oriented), both on the device.  Vertices shared by neighbouring triangles are welded."""
    lib = _lib.load()
    u = _prep(u)
    nx, ny, nz = u.shape
    dev = u.device
    cnt, tab, ec, ea = _mc_tables(dev)
    cells = (nx - 1) * (ny - 1) * (nz - 1)
    counts = torch.empty(cells, device=dev, dtype=torch.int32)
    check(lib.vdn_mc_count(_p(u), nx, ny, nz, float(threshold), _p(cnt), _p(counts), _stream()), "vdn_mc_count")
    incl = torch.cumsum(counts, 0, dtype=torch.int64)
    total = int(incl[-1])                                 # the output size is data dependent: one host read
    if total == 0:
        return torch.zeros(0, 3, device=dev), torch.zeros(0, 3, device=dev, dtype=torch.int64)
    offsets = incl - counts
    keys = torch.empty(total * 3, device=dev, dtype=torch.int64)
    pos = torch.empty(total * 3, 3, device=dev, dtype=torch.float32)
    check(lib.vdn_mc_emit(_p(u), nx, ny, nz, float(threshold), _p(cnt), _p(tab), _p(ec), _p(ea), _p(offsets), _p(keys),
                          _p(pos), _stream()), "vdn_mc_emit")
    uniq, inv = torch.unique(keys, return_inverse=True)
    verts = torch.empty(uniq.numel(), 3, device=dev, dtype=torch.float32)
    verts[inv] = pos                                      # duplicates of a key carry the identical position
    return verts, inv.reshape(-1, 3)
