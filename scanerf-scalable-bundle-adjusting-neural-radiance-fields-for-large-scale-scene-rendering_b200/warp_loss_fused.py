"""Host mirror of the reference's multi-view warp loss (warp_loss.py:147-665, WarpLoss) on the native ops
(the caller right beside the hot path).

What changes against the reference -- the arithmetic does not:
  * the training images stay resident in HBM as uint8 [N,H,W,3]; the neighbour colours are fetched by one kernel
    (cuda.neighbor_sample_forward / _backward) instead of four CPU gathers with a D2H + H2D round trip per step
    (warp_loss.py:441-519);
  * no boolean compaction (`rays_o[selected]`, `nei_origin[nei_valid]`, warp_loss.py:530-535, 627-629): invalid
    rays / neighbours are masked, every shape is static, nothing synchronises with the host.  The re-render of the
    neighbour rays for the visibility score (compute_visibility, :355-377) goes through the masked fused render
    path, which skips masked rays inside every kernel;
  * a batch without any valid ray gives a zero loss instead of None.
The loss value equals the reference's: mean over the selected rays and their K neighbours of
mean_c((pred - neighbour)^2) * score, score = soft_vis * soft_diffuse(neighbour) * soft_diffuse(reference ray).
"""
import torch

from cuda import (computeViewcost, neighbor_sample_backward, neighbor_sample_forward, proj2neighbor_backward,
                  proj2neighbor_forward)
from hashgrid import INFERENCE


class ProjNeighborViewAutoGrad(torch.autograd.Function):
    """warp_loss.py:102-145: (pts [B,3], ks, rts [N,3,4], nei_views, nei_valid [B,K]) -> grid [B,K,3] = K (R p + t),
    nei_origin, nei_direction [B,K,3]; gradients to pts and rts through `grid` only, as in the reference."""

    @staticmethod
    def forward(ctx, pts, ks, rts, nei_views, nei_valid):
        B, K = pts.shape[0], nei_views.shape[1]
        pts, rts = pts.contiguous(), rts.contiguous()
        grid = torch.zeros(B, K, 3, dtype=torch.float32, device=pts.device)
        nei_origin, nei_direction = torch.zeros_like(grid), torch.zeros_like(grid)
        proj2neighbor_forward(pts, ks, rts, nei_views, nei_valid, nei_origin, nei_direction, grid)
        ctx.save_for_backward(pts, ks, rts, nei_views, nei_valid)
        ctx.mark_non_differentiable(nei_origin, nei_direction)
        return grid, nei_origin, nei_direction

    @staticmethod
    def backward(ctx, grad_grid, _go, _gd):
        pts, ks, rts, nei_views, nei_valid = ctx.saved_tensors
        grad_pts = torch.zeros_like(pts)
        grad_rts = torch.zeros(ks.shape[0], 3, 4, dtype=torch.float32, device=pts.device)
        proj2neighbor_backward(pts, ks, rts, nei_views, nei_valid, grad_grid.contiguous(), grad_pts, grad_rts)
        return grad_pts, None, grad_rts, None, None


def proj_neighbor_view(pts, ks, rts, nei_views, nei_valid):
    return ProjNeighborViewAutoGrad.apply(pts, ks, rts, nei_views, nei_valid)


class SampleNeighborColorFn(torch.autograd.Function):
    """(images u8 [N,H,W,3], occlusions bool [N,H,W(,1)] | None, grid [B,K,2], nei_views, nei_valid) ->
    (colour [B,K,3], valid [B,K]); differentiable in `grid` (the bilinear weights), as warp_loss.py:441-519."""

    @staticmethod
    def forward(ctx, images, occlusions, grid, nei_views, nei_valid):
        grid = grid.contiguous()
        color = torch.empty(*grid.shape[:2], 3, dtype=torch.float32, device=grid.device)
        valid = torch.empty(grid.shape[:2], dtype=torch.bool, device=grid.device)
        neighbor_sample_forward(images, occlusions, grid, nei_views, nei_valid, color, valid)
        ctx.save_for_backward(images, grid, nei_views, nei_valid)
        ctx.mark_non_differentiable(valid)
        return color, valid

    @staticmethod
    def backward(ctx, grad_color, _gv):
        images, grid, nei_views, nei_valid = ctx.saved_tensors
        grad_grid = torch.empty_like(grid)
        neighbor_sample_backward(images, grid, nei_views, nei_valid, grad_color.contiguous(), grad_grid)
        return None, None, grad_grid, None, None


class WarpLoss:
    """block: the tile (tile_step.TileStep: .poses, .featureGrid, .render_rays).  images: [N,H,W,3] uint8 (or float in
    [0,1], converted once) -- kept on the device.  alpha / gamma: cfg.TRAINING.LOSS.ALPHA / GAMMA."""

    def __init__(self, block, images, alpha, gamma, topK=10, render_chunk=1 << 16):
        self.block = block
        self.poses = block.poses
        self.device = block.device
        if images.dtype != torch.uint8:
            images = (images.float() * 255.0).round().clamp(0, 255).to(torch.uint8)
        self.images = images.to(self.device).contiguous()
        self.H, self.W = int(images.shape[1]), int(images.shape[2])
        hg = block.featureGrid
        # warp_loss.py:150: the finest cell of the tile (tile_size = the inner half of the doubled box)
        self.voxel_size = float(torch.max((hg.bbox_size / 2.0).cpu() / hg.HE.resolution[-1].cpu().float()))
        self.alpha, self.gamma, self.topK = float(alpha), float(gamma), int(topK)
        self.num_camera = int(self.poses.ks.shape[0])
        self.render_chunk = int(render_chunk)

    def world_to_camera(self):
        """poses.get_rts() (camera_utils.py:76-89), differentiable in se3_refine: the inverse of the fused pose-chain
        kernel's camera-to-world matrices (a handful of launches instead of the ~500 of the torch Taylor chain)."""
        from tile_step import pose_invert
        return pose_invert(self.poses.c2w())

    def soft_vis(self, depth_diff):
        return torch.exp(-self.alpha * depth_diff / self.voxel_size)

    def soft_diffuse(self, specular):
        return torch.exp(-self.gamma * torch.mean(specular, dim=-1, keepdim=True))

    @torch.no_grad()
    def view_selection(self, rays_o, rays_d, pts, rts=None):
        """warp_loss.py:389-413: the topK cheapest views per point; valid where the cost is <= 0.176."""
        B = pts.shape[0]
        cost = torch.full((self.num_camera, B), 1, dtype=torch.float32, device=self.device)
        rts = self.world_to_camera().detach() if rts is None else rts
        computeViewcost(rays_o.contiguous(), rays_d.contiguous(), pts.contiguous(), self.poses.ks, rts.contiguous(), cost, self.H, self.W)
        k = min(self.topK, self.num_camera)
        top, views = torch.topk(cost, k=k, dim=0, largest=False)
        valid = top <= 0.176
        return views.permute(1, 0).int().contiguous(), valid.permute(1, 0).contiguous()

    def projection(self, pts, rts, nei_views, nei_valid):
        """warp_loss.py:415-439."""
        grid, nei_origin, nei_direction = proj_neighbor_view(pts, self.poses.ks, rts, nei_views, nei_valid)
        proj_depth = grid[..., 2:]
        grid = grid[..., :2] / (proj_depth + 1e-8) - 0.5
        return grid, nei_origin, nei_direction, proj_depth

    def sample_neighbor_color(self, grid, nei_views, nei_valid, occlusions):
        return SampleNeighborColorFn.apply(self.images, occlusions, grid, nei_views, nei_valid)

    @torch.no_grad()
    def compute_visibility(self, rays_o, rays_d, proj_depth, valid):
        """warp_loss.py:355-377 over ALL neighbour rays with a validity mask (masked rays are skipped in the kernels)."""
        depth = torch.zeros_like(rays_o[..., :1])
        specular = torch.zeros_like(rays_o)
        for l in range(0, rays_o.shape[0], self.render_chunk):
            s = slice(l, l + self.render_chunk)
            out, ok = self.block.render_rays(rays_o[s].contiguous(), rays_d[s].contiguous(), occlusion_mask=valid[s, None], mode=INFERENCE)
            if ok:
                depth[s] = out["pred_depth"]
                specular[s] = out["pred_specular"]
        return self.soft_vis(torch.abs(depth - proj_depth)), self.soft_diffuse(specular)

    def __call__(self, steps, rays_o, rays_d, depth, diffuse, specular, ray_colors, valid, occlusions, ori_poses_idxs=None):
        """Arguments as warp_loss.py:523; `valid` [B] bool selects the rays that take part."""
        B, K = rays_o.shape[0], min(self.topK, self.num_camera)
        pts = rays_o + depth * rays_d
        rts = self.world_to_camera()
        nei_views, nei_valid = self.view_selection(rays_o.detach(), rays_d.detach(), pts.detach(), rts.detach())
        nei_valid = nei_valid & valid[:, None]
        grid, nei_origin, nei_direction, proj_depth = self.projection(pts, rts, nei_views, nei_valid)
        neighbor_color, nei_valid = self.sample_neighbor_color(grid, nei_views, nei_valid, occlusions)
        flat = nei_valid.reshape(-1)
        vis, nei_diffuse = self.compute_visibility(nei_origin.reshape(-1, 3), nei_direction.reshape(-1, 3),
                                                   proj_depth.detach().reshape(-1, 1), flat)
        with torch.no_grad():
            score = (vis * nei_diffuse * flat[:, None]).reshape(B, K, 1) * self.soft_diffuse(specular)[:, None, :]
            n_sel = valid.sum().clamp_min(1).float()
        pred = torch.clamp(diffuse + specular, 0, 1)
        per_pair = torch.mean((pred[:, None, :] - neighbor_color) ** 2, dim=-1, keepdim=True) * score
        return per_pair.sum() / (n_sel * K)
