"""Drop-in for the reference's hashgrid/PyHashGrid.py: hash grid addressed in
world space inside an axis-aligned box (points are clamped into the box)."""
from ._embedding import _EncodeFn, _HashGridBase
from .lib.HASHGRID import embedding_forward_cuda, embedding_backward_cuda  # noqa: F401  (surface)


class HashEmbeddingAutoGrad(_EncodeFn):
    """autograd.Function(points, features, block_corner, block_size, resolution) -- PyHashGrid.py:9-32."""

    @staticmethod
    def forward(ctx, points, features, block_corner, block_size, resolution):
        return _EncodeFn.forward(ctx, points, features, block_corner, block_size, resolution)

    @staticmethod
    def backward(ctx, grad_in):
        return _EncodeFn.backward(ctx, grad_in)[:5]


def HashEmbedding(points, features, block_corner, block_size, resolution):
    return HashEmbeddingAutoGrad.apply(points, features, block_corner, block_size, resolution)


class PyHashGrid(_HashGridBase):
    _bbox_variant = True
