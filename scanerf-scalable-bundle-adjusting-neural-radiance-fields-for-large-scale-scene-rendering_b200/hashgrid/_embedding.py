"""Shared machinery of PyHashGrid / PyHashGridBG: the autograd bridge to the
sm_100a encode kernels and the level-resolution ladder.

Reference behaviour restated from hashgrid/PyHashGrid.py:9-86 and
hashgrid/PyHashGridBG.py:9-86 (operator contract: outputs [B,L,2] pre-filled with
zeros, backward returns dense grad_points / grad_features, None for the rest).
"""
import torch
import torch.nn as nn

from . import _gradmode
from .lib import HASHGRID as _ops


class _EncodeFn(torch.autograd.Function):
    """points [B,3], features [L,T,2] -> [B,L,2]; corner/size None selects the
    contracted-space (BG) variant."""

    @staticmethod
    def forward(ctx, points, features, block_corner, block_size, resolution, direct_grad=False):
        out = torch.zeros((points.shape[0], features.shape[0], 2), dtype=torch.float32, device=points.device)
        _ops._encode_fwd(points, out, features, block_corner, block_size, resolution)
        ctx.bbox = block_corner is not None
        ctx.block_size = block_size
        ctx.direct_grad = bool(direct_grad)
        if ctx.bbox:
            ctx.save_for_backward(points, features, resolution, block_corner)
        else:
            ctx.save_for_backward(points, features, resolution)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.bbox:
            points, features, resolution, corner = ctx.saved_tensors
            size = ctx.block_size
        else:
            points, features, resolution = ctx.saved_tensors
            corner = size = None
        need_p = ctx.needs_input_grad[0]
        grad_points = torch.zeros_like(points) if need_p else None
        if ctx.direct_grad and _gradmode.mode() is not None and features.is_leaf and features.requires_grad:
            # Inside `table_backward(...)` (the training step's explicit opt-in): scatter straight into the
            # parameter's .grad (the kernel accumulates): no 2 GiB zeros_like + no dense `grad += new` pass per
            # call as in PyHashGridBG.py:27-28.  Outside of it the dense gradient is returned, as the reference does.
            if features.grad is None:
                features.grad = torch.zeros_like(features)
            _ops._encode_bwd(points, grad_out.contiguous(), grad_points, features.grad, features, corner, size, resolution)
            return grad_points, None, None, None, None, None
        # dense grad_features as the reference returns it; the encodes of one backward pass share it (_gradmode)
        grad_features, first = _gradmode.shared_table_grad(features)
        _ops._encode_bwd(points, grad_out.contiguous(), grad_points, grad_features, features, corner, size, resolution)
        return grad_points, (grad_features if first else None), None, None, None, None


def resolution_ladder(base_resolution, finest_resolution, n_levels):
    """Per-level (per-axis) vertex counts res_l = int(base * b**l) with
    b = exp((ln fin - ln base) / (L-1)), evaluated by torch exactly like the
    reference does (PyHashGridBG.py:55-62) so the int ladder is identical."""
    base = torch.as_tensor(base_resolution)
    fin = torch.as_tensor(finest_resolution)
    b = torch.exp((torch.log(fin) - torch.log(base)) / (n_levels - 1))
    return torch.stack([(base * b ** i).int() for i in range(n_levels)], 0)


class _HashGridBase(nn.Module):
    _bbox_variant = False
    # module calls may accumulate the table gradient directly into features.grad, but only inside an explicit
    # `_gradmode.table_backward(...)` context (see _EncodeFn.backward); otherwise -- and always for the
    # reference-named autograd Functions -- a dense grad_features tensor is returned
    direct_grad = True

    def __init__(self, device, bbox_corner, bbox_size, n_levels=16, n_features_per_level=2,
                 log2_hashmap_size=19, base_resolution=16, finest_resolution=512, init_mode="xavier"):
        super().__init__()
        assert n_features_per_level == 2, "we only support dim=2"
        self.bbox_corner = bbox_corner
        self.bbox_size = bbox_size
        self.device = device
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.finest_resolution = finest_resolution
        self.out_dim = n_levels * n_features_per_level
        ladder = resolution_ladder(base_resolution, finest_resolution, n_levels)
        self.b = torch.exp((torch.log(torch.as_tensor(finest_resolution)) -
                            torch.log(torch.as_tensor(base_resolution))) / (n_levels - 1))
        if ladder.dim() == 1:          # scalar resolutions -> same count on every axis
            ladder = ladder[:, None].repeat(1, 3)
        self.resolution = ladder.to(self.device)
        table = torch.zeros(n_levels, 2 ** log2_hashmap_size, n_features_per_level,
                            dtype=torch.float32, device=self.device)
        self.features = nn.Parameter(table)
        if init_mode == "kaiming":
            nn.init.kaiming_normal_(self.features)
        elif init_mode == "xavier":
            nn.init.xavier_normal_(self.features)
        elif init_mode == "uniform":
            nn.init.uniform_(self.features, -1e-4, 1e-4)

    def forward(self, x):
        """x [..., 3] -> [..., n_levels * 2]"""
        lead = list(x.shape[:-1])
        pts = x.reshape(-1, 3)
        if self._bbox_variant:
            size = self.bbox_size
            if not torch.is_tensor(size):
                size = torch.full((3,), float(size), dtype=torch.float32, device=pts.device)
            elif size.numel() == 1:
                size = size.reshape(1).repeat(3).to(torch.float32)
            out = _EncodeFn.apply(pts, self.features, self.bbox_corner, size, self.resolution, self.direct_grad)
        else:
            out = _EncodeFn.apply(pts, self.features, None, None, self.resolution, self.direct_grad)
        return out.reshape(*lead, self.out_dim)
