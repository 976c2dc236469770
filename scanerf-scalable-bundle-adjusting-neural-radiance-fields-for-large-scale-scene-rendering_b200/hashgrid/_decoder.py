"""Host-side mirror of the reference's decoder container (network.py:127-190, ShallowMLP):
same sub-module names, hence the same state_dict keys and `decoder.pth` files, same call
signature `decoder(x[..., 35], weight_feature=mask32) -> {sigma, tint, diffuse, specular}`.

It exists so that the hot path can be driven without the reference checkout (bench, tests,
smoke); when the reference's own `network.ShallowMLP` is passed to HashGrid.render_*_rays
instead, `decoder_params()` reads the very same attributes from it.
"""
import math

import torch
import torch.nn as nn

# state_dict order of network.ShallowMLP(32) = the order of the flat inference layout
# (rendering.py:101-113, hashgrid/include/decoder.h:48-67)
LAYERS = [("Spatial_MLP", 0, 32, 64), ("Spatial_MLP", 2, 64, 64), ("sigma_layer", 0, 32, 1),
          ("diffuse_layer", 0, 32, 3), ("tint_layer", 0, 32, 3), ("Directional_MLP", 0, 48, 64),
          ("Directional_MLP", 2, 64, 64), ("Directional_MLP", 4, 64, 3)]
N_PARAMS = sum(i * o + o for _, _, i, o in LAYERS)      # 13 994 (decoder.h:20)

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh16(d):
    """Degree-3 real spherical harmonics of unit vectors, network.py:38-77."""
    x, y, z = d[..., 0:1], d[..., 1:2], d[..., 2:3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    return torch.cat([
        torch.ones_like(x) * _C0, _C1 * y, _C1 * z, _C1 * x,
        _C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy),
        _C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
        _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy),
        _C3[5] * z * (xx - yy), _C3[6] * x * (xx - 3 * yy)], -1)


class _Gauss(nn.Module):
    """exp(-x^2 / (2 sigma^2)), network.py:79-84"""

    def __init__(self, sigma=0.1):
        super().__init__()
        self.item = 1.0 / (-2 * (sigma ** 2))

    def forward(self, x):
        return torch.exp((x ** 2) * self.item)


class _Stack(nn.Module):
    """Carrier of an `mlp` Sequential laid out like network.GeneralMLP (network.py:127-148)."""

    def __init__(self, widths, act, output_act=False):
        super().__init__()
        layers = []
        for k in range(len(widths) - 1):
            layers.append(nn.Linear(widths[k], widths[k + 1]))
            if k < len(widths) - 2 or output_act:
                layers.append(act)
        self.mlp = nn.Sequential(*layers)

    def forward(self, x):
        return self.mlp(x)


class ShallowMLP(nn.Module):
    def __init__(self, in_channel=32):
        super().__init__()
        assert in_channel == 32
        self.Spatial_MLP = _Stack([32, 64, 64], _Gauss(0.1))
        self.sigma_layer = _Stack([32, 1], nn.Softplus(), True)
        self.diffuse_layer = _Stack([32, 3], nn.Sigmoid(), True)
        self.tint_layer = _Stack([32, 3], nn.Sigmoid(), True)
        self.Directional_MLP = _Stack([48, 64, 64, 3], _Gauss(0.1))
        self.color_act = nn.Sigmoid()
        for m in self.modules():                 # network.init_model(decoder, "xavier"), tile.py:139
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                m.bias.data.fill_(0.0)

    def inference_sigma(self, x):
        return self.sigma_layer(self.Spatial_MLP(x)[..., :32])

    def forward(self, x, **kwargs):
        feats, v = x[..., :-3], x[..., -3:]
        v = v / (v.norm(2, dim=-1, keepdim=True) + 1e-8)
        H = self.Spatial_MLP(feats * kwargs["weight_feature"])
        sigma = self.sigma_layer(H[..., :32])
        tint = self.tint_layer(H[..., :32])
        c_d = self.diffuse_layer(H[..., :32])
        c_s = self.color_act(self.Directional_MLP(torch.cat([H[..., 32:], sh16(v)], -1)))
        return {"diffuse": c_d, "specular": c_s, "sigma": sigma, "tint": tint}


def decoder_params(decoder):
    """The 16 parameter tensors (weight, bias per Linear, LAYERS order) of a ShallowMLP-shaped
    module -- this mirror or the reference's network.ShallowMLP -- or None if it is not one."""
    out = []
    try:
        for name, k, i, o in LAYERS:
            lin = getattr(decoder, name).mlp[k]
            if tuple(lin.weight.shape) != (o, i):
                return None
            out += [lin.weight, lin.bias]
    except (AttributeError, IndexError, TypeError):
        return None
    return out


def flatten_for_inference(decoder):
    """13 994 floats: per Linear, bias then W^T flattened input-major (rendering.py:101-113)."""
    ps = decoder_params(decoder)
    chunks = []
    for k in range(0, len(ps), 2):
        chunks += [ps[k + 1].detach().flatten(), ps[k].detach().t().contiguous().flatten()]
    return torch.cat(chunks)
