"""Positional embedder with the reference's call signature.

Mirrors ``get_embedder(multires, input_dims) -> (fn, out_dim)`` of the reference
(dpt_models/embedder.py:39-51).  The returned callable runs the sm_100a kernel ``vdn_embed_fwd``
(csrc/pointwise.cu); it is differentiable with respect to its input through ``vdn_embed_bwd``.
Layout of the output row is the reference's: ``[x | sin(2^0 x) | cos(2^0 x) | ... ]``, each block
``input_dims`` wide (embedder.py:15-36).

Inside the field networks the embedding is fused into the first-layer kernels and this function is
not on the path; it exists because ``get_embedder`` is part of the reference's public surface.
"""
from __future__ import annotations


def embed_out_dim(multires: int, input_dims: int) -> int:
    return input_dims * (1 + 2 * multires) if multires > 0 else input_dims


class Embedder:
    def __init__(self, multires: int, input_dims: int = 3):
        self.multires = int(multires)
        self.input_dims = int(input_dims)
        self.out_dim = embed_out_dim(self.multires, self.input_dims)

    def embed(self, inputs):
        from . import ops  # deferred: importing the package must not require the GPU library
        return ops.embed(inputs, self.multires)

    __call__ = embed


def get_embedder(multires, input_dims=3):
    eo = Embedder(multires, input_dims)
    return eo.embed, eo.out_dim
