import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, scenes
pkg = importlib.import_module(bench.PKG); pkg.install()
import tile_step as ts
DEV = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
H, W, n_cam = 48, 64, 8
Ks, c2w = scenes.camera_rig(n_cam, H, W, gen, center=(10.0, 3.0, 15.0), radius=5.0, fx=60.0)
locs = torch.stack([torch.randint(0, n_cam, (512,), generator=gen), torch.randint(0, W, (512,), generator=gen), torch.randint(0, H, (512,), generator=gen)], -1).int().to(DEV)
go, gd = torch.randn(512, 3, generator=gen).to(DEV), torch.randn(512, 3, generator=gen).to(DEV)
class RefPoses(ts.Poses):
    def rays(self, locs):
        c2w_ = ts.pose_invert(self.get_rts())
        v = locs[:, 0].long()
        K = self.ks[v]
        x = (locs[:, 1].float() + 0.5 - K[:, 0, 2]) / K[:, 0, 0]
        y = (locs[:, 2].float() + 0.5 - K[:, 1, 2]) / K[:, 1, 1]
        d_cam = torch.stack([x, y, torch.ones_like(x)], -1)
        M = c2w_[v]
        return M[:, :, 3].contiguous(), (M[:, :, :3] @ d_cam[..., None])[..., 0].contiguous()
for scale in (0.0, 1e-4, 1e-3, 1e-2):
    se3 = scale * torch.randn(n_cam, 6, generator=gen)
    out = {}
    for name, cls in (("ours", ts.Poses), ("torch", RefPoses)):
        p = cls(Ks, c2w, DEV, None)
        with torch.no_grad():
            p.se3_refine.copy_(se3.to(DEV))
        o, d = p.rays(locs)
        ((o * go).sum() + (d * gd).sum()).backward()
        out[name] = (o.detach(), d.detach(), p.se3_refine.grad.clone())
    a, b = out["ours"], out["torch"]
    print("scale", scale, "rays_o diff", float((a[0]-b[0]).abs().max()), "rays_d diff", float((a[1]-b[1]).abs().max()),
          "se3 grad rel diff", float((a[2]-b[2]).abs().max() / b[2].abs().max()), flush=True)
    if scale == 1e-3:
        print(" ours grad[0]", a[2][0].tolist()); print(" torch grad[0]", b[2][0].tolist())
