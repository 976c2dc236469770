"""Pure-PyTorch (CPU) restatement of the Python half of the reference's per-tile
hot path -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).

The native ops (hash encode, sampler, ray generation) come from the C oracle
(oracle/native.py); everything the reference does in torch is restated here in
torch, each function citing the reference file:line it follows.  Pinned by
tests/golden/py_golden_*.npz, which were produced by importing the reference's own
network.py / hashgrid/__init__.py / camera.py (tests/golden/make_py_golden.py).
"""
import math

import numpy as np
import torch

from . import native as on

# --------------------------------------------------------------------------- #
# hash encode as an autograd op backed by the C oracle
# --------------------------------------------------------------------------- #


class HashEncodeCPU(torch.autograd.Function):
    """hashgrid/PyHashGridBG.py:9-30 with the CUDA op replaced by the C oracle."""

    @staticmethod
    def forward(ctx, points, table, res):
        out = on.hash_encode_fwd(points.detach().numpy(), table.detach().numpy(), res.numpy())
        ctx.save_for_backward(points, table, res)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, g):
        points, table, res = ctx.saved_tensors
        gp, gt = on.hash_encode_bwd(points.detach().numpy(), g.contiguous().numpy(), table.detach().numpy(), res.numpy())
        return torch.from_numpy(gp), torch.from_numpy(gt), None


def resolution_ladder(base, fin, n_levels=16):
    """hashgrid/PyHashGridBG.py:55-62 (torch fp32 arithmetic on purpose)."""
    base, fin = torch.as_tensor(base), torch.as_tensor(fin)
    b = torch.exp((torch.log(fin) - torch.log(base)) / (n_levels - 1))
    r = torch.stack([(base * b ** i).int() for i in range(n_levels)], 0)
    return r if r.dim() == 2 else r[:, None].repeat(1, 3)


# --------------------------------------------------------------------------- #
# field: contraction, level mask, decoder MLP
# --------------------------------------------------------------------------- #

def contract_fore(x, min_bbox, bbox_size):
    """hashgrid/__init__.py:394-395"""
    return (x - min_bbox) / bbox_size * 4.0 - 2.0


def contract_bg(x, min_bbox, bbox_size):
    """hashgrid/__init__.py:397-411"""
    x = (x - min_bbox) / bbox_size * 4.0 - 2.0
    n, _ = torch.max(torch.abs(x), dim=-1, keepdim=True)
    return x * ((2 - 1.0 / n) / n)


def level_mask(global_step, device="cpu"):
    """hashgrid/__init__.py:228-235 -- BARF-style coarse-to-fine weights, [16]."""
    alpha = max(min(global_step / 10000 * 8 + 8, 16), 0)
    k = torch.arange(16, dtype=torch.float32, device=device)
    return (1 - torch.cos((alpha - k).clamp(min=0, max=1) * np.pi)) / 2


_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh16(d):
    """network.py:38-77 with deg=3: 16 real SH basis values of unit directions."""
    x, y, z = d[..., 0:1], d[..., 1:2], d[..., 2:3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    return torch.cat([
        torch.ones_like(x) * _C0,
        _C1 * y, _C1 * z, _C1 * x,
        _C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy),
        _C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
        _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy),
        _C3[5] * z * (xx - yy), _C3[6] * x * (xx - 3 * yy)], -1)


def gauss_act(x, sigma=0.1):
    """network.py:79-84"""
    return torch.exp((x ** 2) * (1.0 / (-2 * sigma ** 2)))


MLP_KEYS = ["Spatial_MLP.mlp.0", "Spatial_MLP.mlp.2", "sigma_layer.mlp.0", "diffuse_layer.mlp.0",
            "tint_layer.mlp.0", "Directional_MLP.mlp.0", "Directional_MLP.mlp.2", "Directional_MLP.mlp.4"]
MLP_SHAPES = [(64, 32), (64, 64), (1, 32), (3, 32), (3, 32), (64, 48), (64, 64), (3, 64)]


def init_mlp(gen):
    """Random decoder parameters with the reference's layout (state_dict keys of
    network.ShallowMLP(32)) and init (xavier_normal_ weights, zero bias: network.py:202-205)."""
    p = {}
    for k, (o, i) in zip(MLP_KEYS, MLP_SHAPES):
        std = math.sqrt(2.0 / (i + o))
        p[k + ".weight"] = torch.randn(o, i, generator=gen) * std
        p[k + ".bias"] = torch.zeros(o)
    return p


def shallow_mlp(p, feat, viewdirs, mask32):
    """network.py:172-190 (ShallowMLP.forward), functional.  feat [...,32], viewdirs [...,3]."""
    lin = lambda k, x: torch.nn.functional.linear(x, p[k + ".weight"], p[k + ".bias"])
    v = viewdirs / (viewdirs.norm(2, dim=-1, keepdim=True) + 1e-8)
    H = lin(MLP_KEYS[1], gauss_act(lin(MLP_KEYS[0], feat * mask32)))
    sigma = torch.nn.functional.softplus(lin(MLP_KEYS[2], H[..., :32]))
    c_d = torch.sigmoid(lin(MLP_KEYS[3], H[..., :32]))
    tint = torch.sigmoid(lin(MLP_KEYS[4], H[..., :32]))
    x2 = torch.cat([H[..., 32:], sh16(v)], -1)
    h = gauss_act(lin(MLP_KEYS[5], x2))
    h = gauss_act(lin(MLP_KEYS[6], h))
    c_s = torch.sigmoid(lin(MLP_KEYS[7], h))
    return {"sigma": sigma, "diffuse": c_d, "tint": tint, "specular": c_s}


# --------------------------------------------------------------------------- #
# compositing
# --------------------------------------------------------------------------- #

def integrate_weights(sigma, dists, rays_d, infinity):
    """hashgrid/__init__.py:344-360.  sigma [R,S,1], dists [R,S] -> weights [R,S,1], T_left [R]."""
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    if infinity:
        dists = dists.clone()
        dists[:, -1] = 1e10
    alpha = 1.0 - torch.exp(-sigma * dists[..., None])
    T = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1, 1)), 1.0 - alpha + 1e-6], 1), 1)[:, :-1]
    return alpha * T, T[:, -1, 0]


def composite(heads, z_vals, dists, rays_d, infinity, train=True):
    """hashgrid/__init__.py:563-596 (tail of render_batch_rays)."""
    w, T_left = integrate_weights(heads["sigma"], dists, rays_d, infinity)
    acc = lambda a: torch.sum(w * a, 1)
    out = {"depth": acc(z_vals[..., None]), "tint": acc(heads["tint"]), "diffuse": acc(heads["diffuse"]),
           "specular": acc(heads["tint"] * heads["specular"]), "T_left": T_left, "weights": w}
    out["rgb"] = torch.clamp(out["diffuse"] + out["specular"], 0, 1)
    if train:
        out["l2_reg_specular"] = torch.mean(torch.sum(w.detach() * heads["specular"] ** 2, 1))
    return out


def render_batch_rays(table, res, mlp, rays_o, rays_d, z_vals, dists, min_bbox, bbox_size, global_step,
                      background, infinity, train=True):
    """hashgrid/__init__.py:512-596 (render_batch_rays) on CPU."""
    R, S = z_vals.shape
    samples = rays_o[:, None, :] + z_vals[..., None] * rays_d[:, None, :]
    cx = (contract_bg if background else contract_fore)(samples.reshape(-1, 3), min_bbox, bbox_size)
    feat = HashEncodeCPU.apply(cx, table, res).reshape(R, S, 32)
    mask32 = level_mask(global_step)[None, None, :].repeat_interleave(2, dim=-1)
    heads = shallow_mlp(mlp, feat, rays_d[:, None, :].repeat(1, S, 1), mask32)
    return composite(heads, z_vals, dists, rays_d, infinity, train), heads


def inverse_z_sampling(rays_o, rays_d, bbox_center, bbox_size_doubled, S, invalid_underground=True):
    """hashgrid/__init__.py:287-337 (IZ background sampling; box = the un-doubled tile)."""
    bounds = torch.from_numpy(on.ray_aabb(rays_o.numpy(), rays_d.numpy(), bbox_center.numpy(),
                                          (bbox_size_doubled / 2.0).numpy()))[:, 0]
    if invalid_underground:
        out_pt = rays_o + bounds[:, 1:] * rays_d
        floor = (bbox_center - bbox_size_doubled / 4.0)[1]
        valid = ~(torch.abs(out_pt[:, 1] - floor) < 0.0001)
    else:
        valid = torch.ones(rays_o.shape[0], dtype=torch.bool)
    bounds = bounds.clone()
    bounds[torch.any(bounds == -1, dim=-1), 1:] = 0.1
    t = torch.linspace(0.0, 1.0, steps=S)[None, :]
    z = 1.0 / (1.0 / (bounds[:, 1:] + 1e-6) * (1.0 - t) + 1.0 / 1e6 * t)
    d = torch.cat([z[:, 1:] - z[:, :-1], 1e-6 * torch.ones(z.shape[0], 1)], -1)
    return z, d, valid


# --------------------------------------------------------------------------- #
# poses (camera.py)
# --------------------------------------------------------------------------- #

def _taylor(x, kind, nth=10):
    """camera.py:118-141 (taylor_A / taylor_B / taylor_C)."""
    ans = torch.zeros_like(x)
    denom = 1.0
    for i in range(nth + 1):
        if kind == "A":
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        elif kind == "B":
            denom *= (2 * i + 1) * (2 * i + 2)
        else:
            denom *= (2 * i + 2) * (2 * i + 3)
        ans = ans + (-1) ** i * x ** (2 * i) / denom
    return ans


def se3_to_SE3(wu):
    """camera.py:84-95"""
    w, u = wu.split([3, 3], dim=-1)
    w0, w1, w2 = w.unbind(-1)
    O = torch.zeros_like(w0)
    wx = torch.stack([torch.stack([O, -w2, w1], -1), torch.stack([w2, O, -w0], -1), torch.stack([-w1, w0, O], -1)], -2)
    theta = w.norm(dim=-1)[..., None, None]
    I = torch.eye(3)
    A, B, C = _taylor(theta, "A"), _taylor(theta, "B"), _taylor(theta, "C")
    R = I + A * wx + B * wx @ wx
    V = I + B * wx + C * wx @ wx
    return torch.cat([R, V @ u[..., None]], -1)


def pose_invert(p):
    """camera.py:37-43"""
    R, t = p[..., :3], p[..., 3:]
    Ri = R.transpose(-1, -2)
    return torch.cat([Ri, -Ri @ t], -1)


def pose_compose_pair(a, b):
    """camera.py:53-60: pose_new(x) = b o a (x)"""
    Ra, ta, Rb, tb = a[..., :3], a[..., 3:], b[..., :3], b[..., 3:]
    return torch.cat([Rb @ Ra, Rb @ ta + tb], -1)


def rays_from_poses(se3_refine, base_w2c, Ks, px, py):
    """camera_utils.py:65-89 + camera.py:259-281 (get_center_and_ray_v2): the same pixel
    set (px, py) for every camera; returns rays_o, rays_d [N, P, 3]."""
    w2c = pose_compose_pair(se3_to_SE3(se3_refine), base_w2c)
    c2w = pose_invert(w2c)
    X = (px.float() + 0.5)[None, :, None]
    Y = (py.float() + 0.5)[None, :, None]
    hom = torch.cat([X.expand(len(Ks), -1, -1), Y.expand(len(Ks), -1, -1), torch.ones_like(X).expand(len(Ks), -1, -1)], -1)
    cam = hom @ Ks.inverse().transpose(-1, -2)
    R, t = c2w[..., :3], c2w[..., 3]
    world = cam @ R.transpose(-1, -2) + t[:, None, :]
    center = t[:, None, :].expand_as(world)
    return center, world - center
