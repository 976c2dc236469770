"""CPU: pin the oracle's torch restatement (oracle/torch_ref.py) against golden vectors
produced by the reference's own Python code (tests/golden/make_py_golden.py)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import torch_ref as tr


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(a, b, tol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b).max()
    assert err <= tol * max(np.abs(b).max(), 1e-12), f"{what}: abs err {err} vs scale {np.abs(b).max()}"


def golden_loss(out, g):
    """The scalar the golden gradients were taken of (tests/golden/make_py_golden.py)."""
    return ((out["rgb"] - _t(g["target"])) ** 2).mean() + 0.01 * out["l2_reg_specular"] + 0.05 * out["depth"].mean() \
        + 0.1 * (out["T_left"] * _t(g["w_tleft"])).mean()


def _params(g):
    return {k[2:]: _t(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith("p.")}


def test_mlp_forward_backward_matches_reference_network_py():
    g = _load("py_golden_mlp.npz")
    p = _params(g)
    feat, dirs = _t(g["feat"]).requires_grad_(True), _t(g["dirs"]).requires_grad_(True)
    mask16 = tr.level_mask(int(g["global_step"]))
    _close(mask16, g["mask16"], 1e-6, "level mask")
    out = tr.shallow_mlp(p, feat, dirs, mask16[None, :].repeat_interleave(2, dim=-1))
    for k in ("sigma", "diffuse", "tint", "specular"):
        _close(out[k].detach(), g["out." + k], 1e-5, k)
    sum((out[k] * _t(g["w." + k])).sum() for k in out).backward()
    _close(feat.grad, g["g_feat"], 1e-4, "g_feat")
    _close(dirs.grad, g["g_dirs"], 1e-4, "g_dirs")
    for k, v in p.items():
        _close(v.grad, g["g." + k], 1e-4, "g." + k)


def test_render_batch_rays_matches_reference_hashgrid_py():
    for tag, bg in (("fg", False), ("bg", True)):
        g = _load(f"py_golden_render_{tag}.npz")
        p = _params(g)
        table = _t(g["table"]).clone().requires_grad_(True)
        o, d = _t(g["rays_o"]).requires_grad_(True), _t(g["rays_d"]).requires_grad_(True)
        z, dist = _t(g["z_vals"]), _t(g["dists"])
        if bg:   # the IZ sampler restatement reproduces the reference's z_vals
            z2, d2, valid = tr.inverse_z_sampling(o.detach(), d.detach(), _t(g["bbox_center"]), _t(g["bbox_size"]), z.shape[1])
            _close(z2, z, 1e-6, "iz z_vals"); _close(d2, dist, 1e-6, "iz dists")
            assert np.array_equal(valid.numpy(), g["valid"])
        out, _ = tr.render_batch_rays(table, _t(g["res"]), p, o, d, z, dist, _t(g["min_bbox"]), _t(g["bbox_size"]),
                                      int(g["global_step"]), background=bg, infinity=bg)
        for k in ("rgb", "depth", "T_left", "weights", "diffuse", "specular", "tint", "l2_reg_specular"):
            _close(out[k].detach(), g["out." + k], 2e-5, f"{tag} {k}")
        loss = golden_loss(out, g)
        assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5 * max(abs(float(g["loss"])), 1.0)
        loss.backward()
        _close(o.grad, g["g_rays_o"], 1e-4, f"{tag} g_rays_o")
        _close(d.grad, g["g_rays_d"], 1e-4, f"{tag} g_rays_d")
        idx = g["g_table_idx"]
        _close(table.grad[idx[:, 0], idx[:, 1]], g["g_table_val"], 1e-4, f"{tag} g_table")
        assert int((table.grad.abs().sum(-1) > 0).sum()) == len(idx), "same set of touched table entries"
        for k, v in p.items():
            _close(v.grad, g["g." + k], 2e-4, f"{tag} g.{k}")


def test_pose_chain_matches_reference_camera_py():
    g = _load("py_golden_poses.npz")
    se3 = _t(g["se3"]).requires_grad_(True)
    _close(tr.se3_to_SE3(se3).detach(), g["SE3"], 1e-6, "se3_to_SE3")
    base = tr.pose_invert(_t(g["c2w"]))
    W = int(g["W"])
    idx = _t(g["ray_idx"])
    ro, rd = tr.rays_from_poses(se3, base, _t(g["Ks"]), idx % W, idx // W)
    _close(ro.detach(), g["rays_o"], 1e-5, "rays_o")
    _close(rd.detach(), g["rays_d"], 1e-5, "rays_d")
    ((ro * _t(g["w_o"])).sum() + (rd * _t(g["w_d"])).sum()).backward()
    _close(se3.grad, g["g_se3"], 1e-4, "g_se3")
