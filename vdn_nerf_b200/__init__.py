"""vdn_nerf_b200 - B200-native (sm_100a) implementation of VDN-NeRF's neural-SDF volume-rendering hot path.

Drop-in mirrors of the reference's ``dpt_models`` classes:

    from vdn_nerf_b200.fields import SDFNetwork, RenderingNetwork, NeRF, SingleVarianceNetwork
    from vdn_nerf_b200.renderer import NeuSRenderer
    from vdn_nerf_b200.embedder import get_embedder

Importing the package does not need a GPU; every compute call goes to libvdn_b200.so and raises if the
library is not built or the tensors are not on a CUDA device (there is no CPU fallback).
"""
from .embedder import get_embedder  # noqa: F401
from .fields import NeRF, RenderingNetwork, SDFNetwork, SingleVarianceNetwork  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401
from .renderer import NeuSRenderer, extract_fields, extract_fields_sdf, extract_geometry  # noqa: F401

__all__ = ["get_embedder", "SDFNetwork", "RenderingNetwork", "NeRF", "SingleVarianceNetwork", "NeuSRenderer",
           "extract_fields", "extract_fields_sdf", "extract_geometry", "set_precision", "get_precision"]
