#!/usr/bin/env python
"""Generate golden vectors from the REFERENCE's own Python code (run in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

What is imported from the reference, unmodified:
  network.py          ShallowMLP, sh_encoding, init_model
  camera.py           Lie.se3_to_SE3, Pose.compose/invert, get_center_and_ray_v2
  hashgrid/__init__.py HashGrid.render_batch_rays / cal_integrate_weight / contract_* /
                       weight_feature / inverse_z_sampling (methods called on a stand-in
                       object: the class constructor needs the CUDA extensions)
The CUDA extension modules the reference imports at module scope (hashgrid.lib.HASHGRID,
cuda) and the absent third-party modules (easydict, imageio, plyfile, matplotlib) are
replaced by stubs; the hash encode inside render_batch_rays is served by the C oracle
(oracle/native.py), ray_aabb_intersection likewise.

Output: tests/golden/py_golden_{mlp,render,poses}.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SCANERF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import native as on  # noqa: E402
from oracle import torch_ref as tr  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _EasyDict(dict):
    def __getattr__(self, k):
        return self[k]

    def __setattr__(self, k, v):
        self[k] = v


def import_reference():
    _stub("easydict", EasyDict=_EasyDict)
    for n in ("imageio", "plyfile", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "cv2"):
        if n not in sys.modules:
            _stub(n, PlyData=None, PlyElement=None)
    names = ("Sampler ray_block_intersection sample_points prepare_points sort_by_key pts_inference accumulate_color "
             "ray_firsthit_block inverse_z_sampling bg_pts_inference get_last_block update_outgoing_bidx "
             "update_outgoing_bidx_v2 bg_pts_inference_v2 process_occupied_grid embedding_forward_cuda "
             "embedding_backward_cuda embedding_bg_forward_cuda embedding_bg_backward_cuda").split()

    def ray_aabb_intersection(o, d, c, s, bounds):
        bounds.copy_(torch.from_numpy(on.ray_aabb(o.numpy(), d.numpy(), c.numpy(), s.numpy())[:, 0]))

    _stub("cuda", ray_aabb_intersection=ray_aabb_intersection, sample_points_contract=None, voxelize_mesh=None,
          sample_points_grid=None)
    _stub("tools", tools=None)
    _stub("tools.tools")
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_hashgrid", os.path.join(REF, "hashgrid", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(REF, "hashgrid")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_hashgrid"] = mod
    _stub("ref_hashgrid.lib")
    _stub("ref_hashgrid.lib.HASHGRID", **{n: None for n in names})
    spec.loader.exec_module(mod)
    import network, camera  # noqa: E401
    return mod, network, camera


class _OracleHE(torch.nn.Module):
    """Stands in for PyHashGridBG inside the reference's render_batch_rays."""

    def __init__(self, table, res):
        super().__init__()
        self.features = torch.nn.Parameter(table)
        self.resolution = res

    def forward(self, x):
        return tr.HashEncodeCPU.apply(x.reshape(-1, 3), self.features, self.resolution).reshape(*x.shape[:-1], 32)


def main():
    torch.manual_seed(0)
    ref_hg, network, camera = import_reference()
    g = torch.Generator().manual_seed(1234)

    # ---------------- decoder MLP (network.py) ----------------
    dec = network.ShallowMLP(32)
    network.init_model(dec, "xavier")
    for p in dec.parameters():               # non-zero biases so they are exercised
        if p.dim() == 1:
            p.data = 0.1 * torch.randn(p.shape, generator=g)
    state = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    N = 257
    feat = (torch.randn(N, 32, generator=g) * 0.3).requires_grad_(True)
    dirs = torch.randn(N, 3, generator=g).requires_grad_(True)
    mask = ref_hg.HashGrid.weight_feature(types.SimpleNamespace(device="cpu"), 4321)
    mask32 = mask[None, :].repeat_interleave(2, dim=-1)
    out = dec(torch.cat([feat, dirs], -1), weight_feature=mask32)
    wts = {k: torch.randn(v.shape, generator=g) for k, v in out.items()}
    sum((out[k] * wts[k]).sum() for k in out).backward()
    np.savez_compressed(
        os.path.join(HERE, "py_golden_mlp.npz"),
        feat=feat.detach().numpy(), dirs=dirs.detach().numpy(), mask16=mask.numpy(), global_step=4321,
        **{"p." + k: v.numpy() for k, v in state.items()},
        **{"out." + k: v.detach().numpy() for k, v in out.items()},
        **{"w." + k: v.numpy() for k, v in wts.items()},
        g_feat=feat.grad.numpy(), g_dirs=dirs.grad.numpy(),
        **{"g." + k: p.grad.numpy() for k, p in dec.named_parameters()})

    # ---------------- render_batch_rays (hashgrid/__init__.py) ----------------
    L, T = 16, 2 ** 12
    res = tr.resolution_ladder(torch.tensor([24., 16., 36.]), torch.tensor([1536., 1024., 2304.]))
    table0 = torch.randn(L, T, 2, generator=g) * 0.5
    corner, size = torch.tensor([0., 0., 0.]), torch.tensor([20., 13., 30.])
    center = corner + size / 2
    size2 = size * 2
    min_bbox = center - size2 / 2
    R, S = 48, 24
    for tag, bg in (("fg", False), ("bg", True)):
        he = _OracleHE(table0.clone(), res)
        me = types.SimpleNamespace(device="cpu", HE=he, min_bbox=min_bbox, bbox_size=size2, bbox_center=center)
        H = ref_hg.HashGrid
        for fn in ("weight_feature", "cal_integrate_weight", "accumulate", "contract_fore", "contract_bg",
                   "invalid_sampling_underground", "inverse_z_sampling"):
            setattr(me, fn, types.MethodType(getattr(H, fn), me))
        o = (corner + size * (0.2 + 0.6 * torch.rand(R, 3, generator=g))).requires_grad_(True)
        d = (torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * 1.3).requires_grad_(True)
        if bg:
            z, dist, valid = me.inverse_z_sampling(o.detach(), d.detach(), S, True)
        else:
            near = torch.rand(R, 1, generator=g) * 0.5
            step = 0.05 + 0.2 * torch.rand(R, 1, generator=g)
            z = near + step * torch.arange(S)[None, :]
            dist = step.expand(R, S).contiguous()
            valid = torch.ones(R, dtype=torch.bool)
        for p in dec.parameters():
            p.grad = None
        out, ok = H.render_batch_rays(me, o, d, z, dist, dec, ref_hg.TRAIN, me.contract_bg if bg else me.contract_fore,
                                      infinity=bg, global_step=6000)
        assert ok
        tgt = torch.rand(R, 3, generator=g)
        w_tl = torch.rand(R, generator=g)
        loss = ((out["rgb"] - tgt) ** 2).mean() + 0.01 * out["l2_reg_specular"] + 0.05 * out["depth"].mean() \
            + 0.1 * (out["T_left"] * w_tl).mean()
        loss.backward()
        gt = he.features.grad
        nz = torch.nonzero(gt.abs().sum(-1) > 0)
        np.savez_compressed(
            os.path.join(HERE, f"py_golden_render_{tag}.npz"),
            table=table0.numpy(), res=res.numpy(), min_bbox=min_bbox.numpy(), bbox_size=size2.numpy(),
            bbox_center=center.numpy(), rays_o=o.detach().numpy(), rays_d=d.detach().numpy(), z_vals=z.numpy(),
            dists=dist.numpy(), valid=valid.numpy(), global_step=6000, target=tgt.numpy(), w_tleft=w_tl.numpy(), loss=float(loss.detach()),
            **{"p." + k: v.numpy() for k, v in state.items()},
            **{"out." + k: v.detach().numpy() for k, v in out.items()},
            g_rays_o=o.grad.numpy(), g_rays_d=d.grad.numpy(), g_table_idx=nz.numpy(),
            g_table_val=gt[nz[:, 0], nz[:, 1]].numpy(),
            **{"g." + k: p.grad.numpy() for k, p in dec.named_parameters()})

    # ---------------- poses -> rays (camera.py / camera_utils.py:65-89) ----------------
    Ncam, H_, W_ = 5, 54, 96
    c2w = torch.cat([torch.linalg.qr(torch.randn(Ncam, 3, 3, generator=g))[0], torch.randn(Ncam, 3, 1, generator=g)], -1)
    Ks = torch.tensor([[60.0, 0, 48.0], [0, 62.0, 27.0], [0, 0, 1.0]]).expand(Ncam, 3, 3).contiguous()
    base = camera.pose.invert(c2w)
    se3 = (0.05 * torch.randn(Ncam, 6, generator=g)).requires_grad_(True)
    rts = camera.pose.compose([camera.lie.se3_to_SE3(se3), base])
    ray_idx = torch.randint(0, H_ * W_, (40,), generator=g)
    ro, rd = camera.get_center_and_ray_v2(H_, W_, rts, Ks, ray_idx)
    wo, wd = torch.randn(ro.shape, generator=g), torch.randn(rd.shape, generator=g)
    ((ro * wo).sum() + (rd * wd).sum()).backward()
    np.savez_compressed(os.path.join(HERE, "py_golden_poses.npz"), c2w=c2w.numpy(), Ks=Ks.numpy(), se3=se3.detach().numpy(),
                        H=H_, W=W_, ray_idx=ray_idx.numpy(), SE3=camera.lie.se3_to_SE3(se3).detach().numpy(),
                        w2c=rts.detach().numpy(), rays_o=ro.detach().numpy(), rays_d=rd.detach().numpy(),
                        w_o=wo.numpy(), w_d=wd.numpy(), g_se3=se3.grad.numpy())
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
