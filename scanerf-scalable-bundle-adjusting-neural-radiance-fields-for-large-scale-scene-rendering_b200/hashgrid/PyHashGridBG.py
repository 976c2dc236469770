"""Drop-in for the reference's hashgrid/PyHashGridBG.py: hash grid over the
contracted space [-2,2]^3 (foreground tile in [-1,1]^3, background outside)."""
from ._embedding import _EncodeFn, _HashGridBase
from .lib.HASHGRID import embedding_bg_forward_cuda, embedding_bg_backward_cuda  # noqa: F401  (surface)


class HashEmbeddingBGAutoGrad(_EncodeFn):
    """autograd.Function(points, features, resolution) -- PyHashGridBG.py:9-30."""

    @staticmethod
    def forward(ctx, points, features, resolution):
        return _EncodeFn.forward(ctx, points, features, None, None, resolution)

    @staticmethod
    def backward(ctx, grad_in):
        gp, gf = _EncodeFn.backward(ctx, grad_in)[:2]
        return gp, gf, None


def HashEmbeddingBG(points, features, resolution):
    return HashEmbeddingBGAutoGrad.apply(points, features, resolution)


class PyHashGridBG(_HashGridBase):
    _bbox_variant = False
