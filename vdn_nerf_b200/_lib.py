"""ctypes binding of libvdn_b200.so (the C ABI declared in include/vdn_b200.h).

There is no CPU fallback: if the shared library is missing, or a call returns a CUDA error, this module
raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C
vdn_nerf_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdn_b200.so")
ABI_VERSION = 5

P, I, L, F = c_void_p, c_int, c_longlong, c_float

# name -> (restype, argtypes); must list every symbol declared in include/vdn_b200.h
SIGNATURES = {
    "vdn_abi_version": (I, []),
    "vdn_launch_count": (L, []),
    "vdn_error_string": (c_char_p, [I]),
    "vdn_prof_enable": (I, [I]),
    "vdn_prof_read": (I, [I, P, P, P]),
    "vdn_prof_read_bytes": (I, [I, P]),
    "vdn_mlp_layout": (L, [I, P, P, P, P, P]),
    "vdn_mlp_pack": (I, [I, P, P, P, P, P, P, P, P, P, P]),
    "vdn_mlp_unpack_grads": (I, [I, P, P, P, P, P, P, P, P, P, P, P]),
    "vdn_set_mode": (I, [I]),
    "vdn_get_mode": (I, []),
    "vdn_set_chain": (I, [I]),
    "vdn_get_chain": (I, []),
    "vdn_tc_fault": (I, []),
    "vdn_set_fault_flag": (I, [P]),
    "vdn_tc_fault_async": (I, [P, P]),
    "vdn_debug_timeline": (I, [P]),
    "vdn_sdf_layer_dims": (I, [P, P, P]),
    "vdn_sdf_layer_orot": (I, [P, P]),
    "vdn_sdf_blob_floats": (L, [P, L, I]),
    "vdn_sdf_blobg_floats": (L, [P, L]),
    "vdn_sdf_bwd_ws_floats": (L, [P, L]),
    "vdn_sdf_forward": (I, [P, F, P, P, L, P, I, P, I, P, I, P]),
    "vdn_sdf_normals": (I, [P, F, P, P, L, P, P, P, P]),
    "vdn_sdf_backward": (I, [P, F, P, P, L, P, P, P, I, P, I, P, P, P, P, P]),
    "vdn_grid_sdf": (I, [P, F, P, P, P, P, I, I, I, I, F, P, P, P, P]),
    "vdn_rendernet_layer_dims": (I, [P, P, P]),
    "vdn_rendernet_layer_rot": (I, [P, P]),
    "vdn_rendernet_blob_floats": (L, [P, L]),
    "vdn_rendernet_bwd_ws_floats": (L, [P, L]),
    "vdn_rendernet_forward": (I, [P, P, P, P, P, P, I, L, P, P, P]),
    "vdn_rendernet_backward": (I, [P, P, L, P, P, P, P, P, P, P]),
    "vdn_nerf_layer_dims": (I, [P, P, P]),
    "vdn_nerf_layer_orot": (I, [P, P]),
    "vdn_nerf_blob_floats": (L, [P, L]),
    "vdn_nerf_bwd_ws_floats": (L, [P, L]),
    "vdn_nerf_forward": (I, [P, P, P, P, L, P, P, P, P, P]),
    "vdn_nerf_backward": (I, [P, P, P, P, L, P, P, P, P, P, P, P, P, P]),
    "vdn_embed_fwd": (I, [P, L, I, I, P, P]),
    "vdn_embed_bwd": (I, [P, L, I, I, P, P, P]),
    "vdn_ray_points": (I, [P, P, P, L, I, P, P]),
    "vdn_upsample_step": (I, [P, P, P, I, P, I, P, I, P, F, I, L, P, P, P, P, P, P, P]),
    "vdn_merge_sorted": (I, [P, I, P, I, L, P, P, P]),
    "vdn_fine_prep": (I, [P, P, P, F, L, I, P, P, P, P]),
    "vdn_bg_prep": (I, [P, P, P, I, P, I, F, L, P, P, P, P]),
    "vdn_composite_fwd": (I, [L, I, I, I] + [P] * 14 + [F] + [P] * 7 + [P]),
    "vdn_composite_bwd": (I, [L, I, I, I] + [P] * 14 + [F] + [P] * 15 + [P]),
    "vdn_adam_step": (I, [I, P, P, P, P, P, P, P]),
    "vdn_color_loss": (I, [P, P, P, L, P, P, P]),
    "vdn_raygen_fwd": (I, [P, P, L, P, P, P, P, P]),
    "vdn_raygen_bwd": (I, [P, P, L, P, P, P, P, P]),
    "vdn_mc_count": (I, [P, I, I, I, F, P, P, P]),
    "vdn_mc_emit": (I, [P, I, I, I, F, P, P, P, P, P, P, P, P]),
}

_lib = None


class VdnLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VdnLibraryError(
            f"{LIB_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -c \"import __graft_entry__ as g; g.build()\"` from the repository root.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.vdn_abi_version() != ABI_VERSION:
        raise VdnLibraryError(f"libvdn_b200.so ABI {lib.vdn_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().vdn_error_string(code)
        raise VdnLibraryError(f"{what} failed: CUDA error {code} ({msg.decode() if msg else '?'})")


def int_array(values):
    arr = (c_int * len(values))(*[int(v) for v in values])
    return arr


def ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))(*[c_void_p(p) if p else None for p in ptrs])
    return arr
