"""Host-side contract tests that need no GPU: the drop-in boundary (constructor kwargs, state_dict keys,
parameter counts), the C-ABI surface (header == binding == exported symbols), packed-weight layout, sharding
helpers, and the "fail loudly" rule."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from tests import util
from vdn_nerf_b200 import _lib, configs, dist as vdist, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vdn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vdn_[a-z0-9_]+)\s*\(", src)))


def test_abi_header_binding_and_library_agree():
    hdr = header_symbols()
    assert hdr == sorted(_lib.SIGNATURES.keys())
    lib = _lib.load()                      # loads without a GPU; raises if a symbol is missing
    assert lib.vdn_abi_version() == _lib.ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vdn_[a-z0-9_]+)", out)))
    assert exported == hdr


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_state_dict_keys_and_param_counts():
    (nerf, sdf, var, col, dep), conf = util.build("womsk_white_wdepth")
    count = lambda m: sum(p.numel() for p in m.parameters())
    assert count(sdf) == 529076 and count(col) == 273414 and count(dep) == 297408 and count(nerf) == 618980
    assert count(var) == 1
    keys = list(sdf.state_dict().keys())
    assert keys[:3] == ["lin0.bias", "lin0.weight_g", "lin0.weight_v"] and len(keys) == 27
    assert tuple(sdf.lin3.weight_v.shape) == (217, 256) and tuple(sdf.lin8.weight_v.shape) == (257, 256)
    assert tuple(col.lin0.weight_v.shape) == (256, 289)
    nk = list(nerf.state_dict().keys())
    for k in ("pts_linears.5.weight", "views_linears.0.weight", "feature_linear.bias", "alpha_linear.weight",
              "rgb_linear.weight", "dpt_linear.weight"):
        assert k in nk
    assert tuple(nerf.pts_linears[5].weight.shape) == (256, 340)
    assert list(var.state_dict().keys()) == ["variance"]
    (nerf2, *_), _ = util.build("womsk_white")
    assert count(nerf2) == 606596
    # state_dict round trip (checkpoint compatibility, dpt_runner.py:350-381)
    sd = {k: v.clone() for k, v in sdf.state_dict().items()}
    (_, sdf_b, *_), _ = util.build("womsk_white_wdepth", seed=1)
    sdf_b.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(sdf.state_dict().values(), sdf_b.state_dict().values()))


def test_layer_dims_and_layout_from_the_library():
    (nerf, sdf, var, col, dep), conf = util.build("womsk_white_wdepth")
    h = sdf.handle()
    assert h.mlp.in_dims == [39] + [256] * 8
    assert h.mlp.out_dims == [256, 256, 256, 217, 256, 256, 256, 256, 257]
    assert col.handle().mlp.in_dims == [289, 256, 256, 256, 256] and dep.handle().mlp.out_dims[-1] == 96
    nh = nerf.handle()
    assert nh.mlp.in_dims == [84, 256, 256, 256, 256, 340, 256, 256, 256, 283, 128]
    assert nh.mlp.out_dims == [256] * 8 + [257, 128, 99]
    # offsets are contiguous, 16-float aligned, W | W^T | b per layer
    m = h.mlp
    off = 0
    for l in range(m.L):
        ild, old = (m.in_dims[l] + 15) // 16 * 16, (m.out_dims[l] + 15) // 16 * 16
        assert (m.off_w[l], m.off_wt[l], m.off_b[l]) == (off, off + old * ild, off + 2 * old * ild)
        off += 2 * old * ild + old
    assert m.total >= off                       # the tcgen05 tile images follow
    assert len(m.params) == 27 and len(nh.mlp.params) == 26


def test_no_cpu_fallback():
    (nerf, sdf, var, col, dep), conf = util.build("womsk_white")
    with pytest.raises(_lib.VdnLibraryError):
        sdf(torch.zeros(4, 3))
    with pytest.raises(_lib.VdnLibraryError):
        ops.ray_points(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vdn_nerf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), fn


def test_shard_range_partitions():
    for n in (0, 1, 7, 512, 513):
        for ws in (1, 2, 3, 8):
            spans = [vdist.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_host_helpers_of_the_graph_capturable_step():
    """No host->device traffic per step: the sampler's host-evaluated linspace is cached on the device, the packed
    weight cache can be forced to re-materialise (what a CUDA-graph capture needs), the split SDF entry point fails as
    loudly as the reference-shaped one on CPU tensors."""
    from vdn_nerf_b200.renderer import NeuSRenderer
    mods, conf = util.build("womsk_white")
    rend = NeuSRenderer(*mods, **conf["neus_renderer"])
    a = rend._lin(0.0, 1.0, 64, torch.device("cpu"))
    assert a is rend._lin(0.0, 1.0, 64, torch.device("cpu"))
    assert torch.equal(a, torch.linspace(0.0, 1.0, 64))
    assert rend._lin(1e-3, 1.0 - 1.0 / 33.0, 32, torch.device("cpu")) is not a
    assert ops._FORCE_REPACK is False
    ops.force_repack(True)
    assert ops._FORCE_REPACK is True
    ops.force_repack(False)
    assert ops._FORCE_REPACK is False
    with pytest.raises(_lib.VdnLibraryError):
        mods[1].forward_split(torch.zeros(4, 3))


def test_committed_ncu_summaries_feed_the_bench_line():
    """bench.py's `roofline.traffic` comes from the committed `ncu --set full` summary of the dominant kernel: the file must
    parse, cover the eight chain launches of a step, and its DRAM bytes must be close to the algorithmic bytes the chains
    are built to move (no re-reads) - the check that caught the software L2 prefetch doubling the reads."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per_launch = bench.ncu_traffic("chain_train")
    assert per_launch is not None
    text = open(os.path.join(root, "profiles", "r02_ncu_full_chain_train.txt")).read()
    assert text.count("== launch") == 8
    assert "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum" in text      # tcgen05 MMAs were counted
    algorithmic = 561610752.0            # bench.py: algorithmic_bytes_per_launch of the 512-ray womsk_white step
    assert 0.9 < per_launch / algorithmic < 1.1, per_launch
    assert bench.ncu_traffic("wgrad16") > 5e8 and bench.ncu_traffic("no_such_kernel") is None


def test_gradient_arena_hands_out_each_slot_once_per_step():
    """ops.set_grad_arena / _grad_buffer (used by the weight-norm backward when dist.FlatGradAllReduce is active): a slot of
    the flat all-reduce buffer is handed out as a fresh alias once per step and only while `.grad` is None; everything
    else gets a private tensor, so repeated use of a network in one graph and gradient accumulation stay correct."""
    a = torch.nn.Parameter(torch.zeros(4, 3))
    b = torch.nn.Parameter(torch.zeros(5))
    flat = torch.zeros(17)
    views = [flat[:12].view(4, 3), flat[12:].view(5)]
    try:
        ops.set_grad_arena([a, b], views)
        g1 = ops._grad_buffer(a)
        assert g1.data_ptr() == views[0].data_ptr() and g1 is not views[0]        # an alias autograd may adopt
        assert ops._grad_buffer(a).data_ptr() != views[0].data_ptr()               # second use in the same step: private
        b.grad = torch.ones(5)
        assert ops._grad_buffer(b).data_ptr() != views[1].data_ptr()               # accumulation: never the slot
        b.grad = None
        ops.reset_grad_arena_use()
        assert ops._grad_buffer(a).data_ptr() == views[0].data_ptr()               # next step: handed out again
        assert ops._grad_buffer(b).data_ptr() == views[1].data_ptr()
        c = torch.nn.Parameter(torch.zeros(2, 2))
        assert ops._grad_buffer(c).shape == (2, 2)                                 # not registered: private
        ops.set_grad_arena([c], [flat[:6].view(2, 3)])
        assert ops._grad_buffer(c).shape == (2, 2)                                 # stale entry of another shape: ignored
    finally:
        ops.set_grad_arena([], [])
