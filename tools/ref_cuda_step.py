#!/usr/bin/env python
"""Measure the REFERENCE's training step on this B200: the reference's own CUDA extensions (rebuilt unmodified
into oracle/_ref/ by oracle/build_ref.py) inside the reference's own compute graph, on the workload bench.py
uses (BASELINE.json configs[1]).  This is the denominator of the north star's ">= 5x the reference's own
CUDA-extension training rays/s per B200"; it is evidence kept under profiles/, not a bench arm (the driver's
reference arm for this tier is the CPU restatement, see bench.py --impl reference).

The reference's drivers (tile.py, hashgrid/__init__.py) cannot be imported here (easydict, imageio, plyfile,
matplotlib are absent), so the step is sequenced by the same torch code the reference
runs, with every native op going to the REFERENCE kernels:
  pose chain            torch ops of camera.py:84-141 / 37-60 (restated in tile_step.py)          [reference: torch]
  ray generation        torch pinhole math of camera.py:259-281                                    [reference: torch]
  sample placement      reference CUDA_EXT.sample_points_grid / ray_aabb_intersection              [reference kernels]
  valid-ray compaction  boolean indexing + scatter back (hashgrid/__init__.py:419-451)              [reference: torch]
  contraction           torch (hashgrid/__init__.py:394-411)                                        [reference: torch]
  hash encode fwd/bwd   reference HASHGRID embedding_bg_forward/backward_cuda with the zeros_like
                        allocations of PyHashGridBG.py:11-30                                        [reference kernels]
  decoder               torch ShallowMLP (cuBLAS SGEMM + elementwise), network.py:151-190          [reference: torch]
  compositing           torch cumprod chain, hashgrid/__init__.py:344-366, 564-596                 [reference: torch]
  optimisers            dense torch.optim.Adam over the whole table (tile.py:301) + Adam            [reference: torch]

  python tools/ref_cuda_step.py [--steps 10] [--warmup 3] [--out profiles/r1_reference_cuda_step.json]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def build_reference_step(dev, tile_corner, tile_size, Ks, c2w, log2T, grid_resolution, sampler_log2dim, S, S_bg, ply, global_step=10000,
                         lr_table=1e-3, lr_decoder=1e-3, lr_cam=1e-4):
    """A TileStep whose render path is the reference's op-by-op graph on the REFERENCE's kernels (oracle/_ref), with the
    reference's optimisers (dense torch Adam over the table).  pkg.install() must have run."""
    import oracle
    REF_CUDA, REF_HASH = oracle.ref_module("CUDA_EXT"), oracle.ref_module("HASHGRID_EMBED")   # the rebuilt reference modules
    if REF_CUDA is None or REF_HASH is None:
        raise RuntimeError("oracle/_ref/{CUDA_EXT,HASHGRID_EMBED}.so not built (python oracle/build_ref.py)")
    import tile_step as ts
    from hashgrid import HashGrid, TRAIN
    from hashgrid._decoder import ShallowMLP

    class RefEncode(torch.autograd.Function):          # hashgrid/PyHashGridBG.py:9-30, literally
        @staticmethod
        def forward(ctx, points, features, resolution):
            outputs = torch.full((points.shape[0], features.shape[0], 2), 0, dtype=torch.float32, device=points.device)
            REF_HASH.embedding_bg_forward_cuda(points, outputs, features, resolution)
            ctx.save_for_backward(points, features, resolution)
            return outputs

        @staticmethod
        def backward(ctx, grad_in):
            points, features, resolution = ctx.saved_tensors
            grad_points = torch.zeros_like(points)
            grad_features = torch.zeros_like(features)
            REF_HASH.embedding_bg_backward_cuda(points, grad_in.contiguous(), grad_points, grad_features, features, resolution)
            return grad_points, grad_features, None

    class RefGrid(HashGrid):
        """HashGrid with the reference's op-by-op render path (no fused kernels of this repo on it)."""
        fused_decoder = False
        fused_encode = False

        def samplePoints(self, rays_o, rays_d, num_sample, out=None):
            z = torch.full((rays_o.shape[0], num_sample), -1, dtype=torch.float32, device=self.device)
            d = torch.full((rays_o.shape[0], num_sample), -1, dtype=torch.float32, device=self.device)
            REF_CUDA.sample_points_grid(rays_o, rays_d, z, d, self.min_bbox + self.bbox_size / 4.0, self.bbox_size / 2.0,
                                        self.occupied_grid, self.sampler_log2dim)
            return z, d

        def inverse_z_sampling(self, rays_o, rays_d, num_sample, invalid_underground=True, perturb=False, out=None):
            bounds = torch.full((rays_o.shape[0], 2), -1, dtype=torch.float32, device=rays_o.device)
            REF_CUDA.ray_aabb_intersection(rays_o, rays_d, self.bbox_center, self.bbox_size / 2.0, bounds)
            valid = torch.ones_like(rays_d[..., 0]).bool()
            bounds[torch.any(bounds == -1, dim=-1), 1:] = 0.1
            t = torch.linspace(0.0, 1.0, steps=num_sample, device=self.device)[None, :]
            z_vals = 1.0 / (1.0 / (bounds[:, 1:] + 1e-6) * (1.0 - t) + 1.0 / 1e6 * t)
            z_vals = z_vals.expand([rays_o.shape[0], num_sample])
            dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], 1e-6 * torch.ones_like(z_vals[:, :1])], -1)
            return z_vals, dists, valid

        def render_batch_rays(self, rays_o, rays_d, z_vals, dists, decoder, mode, contract_func, out_normal=False, infinity=False, **kw):
            if z_vals.shape[0] == 0:
                return None, False
            R, S_ = z_vals.shape
            samples = rays_o[:, None, :] + z_vals[..., None] * rays_d[:, None, :]
            cx, _ = contract_func(samples.reshape(-1, 3))
            feats = RefEncode.apply(cx.contiguous(), self.HE.features, self.HE.resolution).reshape(R, S_, 32)
            mask32 = self.weight_feature(kw["global_step"])[None, None, :].repeat_interleave(2, dim=-1)
            heads = decoder(torch.cat([feats, rays_d[:, None, :].repeat(1, S_, 1)], -1), weight_feature=mask32)
            weights, T_left = self.cal_integrate_weight(heads["sigma"], z_vals, dists.clone(), rays_d, infinity=infinity)
            out = {"depth": self.accumulate(weights, z_vals[..., None]), "tint": self.accumulate(weights, heads["tint"]),
                   "diffuse": self.accumulate(weights, heads["diffuse"]),
                   "specular": self.accumulate(weights, heads["tint"] * heads["specular"]), "T_left": T_left, "weights": weights}
            out["rgb"] = torch.clamp(out["diffuse"] + out["specular"], 0, 1)
            if mode is TRAIN:
                out["l2_reg_specular"] = torch.mean(torch.sum(weights.detach() * heads["specular"] ** 2, 1))
            return out, True

    class RefPoses(ts.Poses):
        def rays(self, locs):                           # camera.py:259-281 in torch, autograd to se3_refine
            c2w_ = ts.pose_invert(self.get_rts())
            v = locs[:, 0].long()
            K = self.ks[v]
            x = (locs[:, 1].float() + 0.5 - K[:, 0, 2]) / K[:, 0, 0]
            y = (locs[:, 2].float() + 0.5 - K[:, 1, 2]) / K[:, 1, 1]
            d_cam = torch.stack([x, y, torch.ones_like(x)], -1)
            M = c2w_[v]
            return M[:, :, 3].contiguous(), (M[:, :, :3] @ d_cam[..., None])[..., 0].contiguous()

    f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev)
    step = ts.TileStep.__new__(ts.TileStep)
    step.device = dev
    step.featureGrid = RefGrid(dev, f(tile_corner), f(tile_size), log2T, list(grid_resolution), sampler_log2dim, False, ply)
    step.decoder = ShallowMLP(32).to(dev)
    step.poses = RefPoses(Ks, c2w, dev, None)
    step.num_sample, step.num_bg_sample, step.global_step, step.invalid_underground = S, S_bg, global_step, False
    step.consensus = None
    step.warp = None
    step.camera_ids = None
    step.two_streams = False
    step.joint_chains = False
    step.fused_loss = False
    step.featureGrid_optimizer = torch.optim.Adam([{"params": step.featureGrid.parameters(), "lr": lr_table, "betas": (0.9, 0.99), "eps": 1e-15}])
    step.optimizer = torch.optim.Adam([{"params": step.decoder.parameters(), "lr": lr_decoder, "weight_decay": 1e-6},
                                       {"params": step.poses.se3_refine, "lr": lr_cam}])
    return step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_reference_cuda_step.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    import scenes
    import tempfile
    gen = torch.Generator().manual_seed(0)
    c = [cfg["tile_corner"][i] + cfg["tile_size"][i] * f for i, f in enumerate((0.5, 0.25, 0.5))]
    Ks, c2w = scenes.camera_rig(cfg["n_cam"], cfg["H"], cfg["W"], gen, center=tuple(c),
                                radius=0.3 * min(cfg["tile_size"][0], cfg["tile_size"][2]), fx=cfg["fx"])
    ply = os.path.join(tempfile.mkdtemp(prefix="snrf_ref_"), "mesh.ply")
    scenes.write_proxy_mesh_ply(ply, cfg["tile_corner"], cfg["tile_size"], seed=0)
    torch.manual_seed(0)
    step = build_reference_step(dev, cfg["tile_corner"], cfg["tile_size"], Ks, c2w, cfg["log2T"], cfg["grid_resolution"],
                                cfg["sampler_log2dim"], cfg["S"], cfg["S_bg"], ply)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, args.warmup + args.steps, gen)]
    for b in batches[:args.warmup]:
        step.step_device(*b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in batches[args.warmup:]:
        loss = step.step_device(*b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    B = batches[0][0].shape[0]
    line = {"what": "reference CUDA extensions (rebuilt for sm_100a, unmodified) in the reference's torch graph, same workload as bench.py",
            "metric": "train rays/s (fwd+bwd)", "value": B / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": args.steps,
            "warmup": args.warmup, "rays_per_step": B, "loss": float(loss), "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(line))
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
