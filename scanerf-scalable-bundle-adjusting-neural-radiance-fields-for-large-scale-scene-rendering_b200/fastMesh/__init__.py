"""Drop-in for the reference's `fastMesh` package (fastMesh/__init__.py:9-133): proxy-mesh
depth, occlusion masks and background depth sampling on top of the sm_100a ray/mesh kernels."""
import torch

from .lib.fastMesh import fastMesh

from cuda import background_sampling_cuda, ray_aabb_intersection


class FastMesh:
    def __init__(self, path):
        self.fmesh = fastMesh()
        self.fmesh.build(path)

    def set(self, bbox_center, bbox_size):
        self.bbox_center = bbox_center
        self.bbox_size = bbox_size

    def get_sceneinfo(self):
        return self.fmesh.getSceneBound()

    @torch.no_grad()
    def render_depth(self, rays_o, rays_d):
        depth = torch.zeros_like(rays_o[..., :1])
        self.fmesh.fisrtHit(rays_o, rays_d, depth)
        return depth

    @torch.no_grad()
    def render_mask(self, rays_o, rays_d, trust_mesh=False):
        """True where the ray may see the tile: mesh (entry) depth behind the tile entry, no mesh
        on the ray, or origin inside the tile (fastMesh/__init__.py:28-45)."""
        depth = torch.zeros_like(rays_o[..., :1])
        (self.fmesh.fisrtHit if trust_mesh else self.fmesh.firstEnter)(rays_o, rays_d, depth)
        bounds = torch.ones_like(rays_o[..., :2]) * -1
        ray_aabb_intersection(rays_o, rays_d, self.bbox_center, self.bbox_size, bounds)
        inside = torch.all(torch.abs(rays_o - self.bbox_center) < (self.bbox_size / 2.0), dim=-1, keepdim=True)
        return ((depth > bounds[..., :1]) & (bounds[..., :1] != -1)) | (depth == 0) | inside

    @torch.no_grad()
    def sample_points(self, rays_o, rays_d, start, num_sample):
        z_vals = torch.full((rays_o.shape[0], num_sample), -1, dtype=torch.float32, device=rays_o.device)
        self.fmesh.sample_points(rays_o, rays_d, start, z_vals)
        return z_vals

    @torch.no_grad()
    def compute_bgdepth_batch(self, rays_o, rays_d):
        """Depth of the first mesh hit beyond the tile exit (fastMesh/__init__.py:54-78).
        NOTE: like the reference this advances `rays_o` in place for rays that hit the tile."""
        depth_z = torch.zeros_like(rays_o[..., :1])
        self.fmesh.fisrtHit(rays_o, rays_d, depth_z)
        bounds = torch.full((rays_o.shape[0], 2), -1, dtype=torch.float32, device=rays_o.device)
        ray_aabb_intersection(rays_o, rays_d, self.bbox_center, self.bbox_size, bounds)
        valid = bounds[:, 1] != -1
        rays_o[valid] = rays_o[valid] + bounds[valid, 1:] * rays_d[valid]
        bg_z = torch.zeros_like(rays_o[..., :1])
        self.fmesh.fisrtHit(rays_o, rays_d, bg_z)
        bg_z[depth_z == 0] = 1000
        has_bg = (bg_z[..., 0] > 0) & valid
        bg_z[valid] = bg_z[valid] + bounds[valid, 1:]
        return bg_z, has_bg, bounds

    @torch.no_grad()
    def background_sampling(self, rays_o, rays_d, num_sample, sample_range):
        bg_z, valid, bounds = self.compute_bgdepth_batch(rays_o.clone(), rays_d)
        z_vals = torch.full((rays_o.shape[0], num_sample), -1, dtype=torch.float32, device=rays_o.device)
        background_sampling_cuda(rays_o, rays_d, bounds[:, 1:], bg_z, z_vals, num_sample, sample_range)
        return z_vals, valid

    @torch.no_grad()
    def compute_bgdepth(self, poses, H, W):
        """Per-camera background depth maps [N,H,W] (fastMesh/__init__.py:100-133)."""
        num_camera, device = poses.ks.shape[0], poses.device
        bg_depths = torch.zeros(num_camera, H, W, dtype=torch.float32, device=device)
        all_rays_o, all_rays_d = poses.getRays(H, W)
        for idx in range(num_camera):
            rays_o, rays_d = all_rays_o[idx].contiguous(), all_rays_d[idx].contiguous()
            bounds = torch.full((H * W, 2), -1, dtype=torch.float32, device=device)
            bg_z = torch.zeros((H * W, 1), dtype=torch.float32, device=device)
            ray_aabb_intersection(rays_o, rays_d, self.bbox_center, self.bbox_size, bounds)
            valid = bounds[:, 1] != -1
            rays_o[valid] = rays_o[valid] + bounds[valid, 1:] * rays_d[valid]
            self.fmesh.fisrtHit(rays_o, rays_d, bg_z)
            has_no_bg = bg_z <= 0
            bg_z[valid] = bg_z[valid] + bounds[valid, 1:]
            bg_z[has_no_bg] = 0
            bg_depths[idx] = bg_z.reshape(H, W)
        return bg_depths
