import sys, torch, numpy as np
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import test_render_gpu as T
from test_render_gpu import *
_ops = T._ops
ours, ref = _ops()
nb, B = 3, 6000
sc = dev(make_scene(nb)); o, d = (t.to(DEV) for t in make_rays(B, nb * B))
isect = torch.full((B, nb, 2), MISS, device=DEV); ours.ray_block_intersection(o, d, sc["corners"], sc["sizes"], isect)
tb = torch.argsort(isect[..., 0], dim=-1).int().contiguous()
bgb = torch.full((B, 4), -1, dtype=torch.int16, device=DEV); bgw = torch.zeros(B, 4, device=DEV)
ours.update_outgoing_bidx(o, d, sc["corners"], sc["sizes"], tb, isect, bgb, bgw, 0.12, False)
Sb = 24
bz = torch.full((B, Sb), -1.0, device=DEV); ours.inverse_z_sampling(isect, bgb[..., 0].contiguous(), bz, 1e6)
outs = {}
for name, mod in (("ours", ours), ("ref", ref)):
    a = [torch.full((B, Sb, 3), 0.5, device=DEV), torch.full((B, Sb, 3), 0.5, device=DEV), torch.full((B, Sb, 1), 0.5, device=DEV)]
    mod.bg_pts_inference_v2(o, d, bz, bgb, 0, sc["corners"], sc["sizes"], sc["res"], sc["tables"], sc["params"], *a)
    torch.cuda.synchronize()
    outs[name] = a
    print(name, "nan count", [int(torch.isnan(t).sum()) for t in a])
bad = torch.isnan(outs["ours"][0]).any(-1) | torch.isnan(outs["ref"][0]).any(-1)
idx = bad.nonzero()[:8]
print(idx.tolist())
for r, k in idx.tolist()[:4]:
    print("ray", r, "k", k, "z", float(bz[r, k]), "o", o[r].tolist(), "d", d[r].tolist(), "ours", outs["ours"][2][r, k].item(), "ref", outs["ref"][2][r, k].item())
good = ~bad
print("max err on non-nan", float((outs["ours"][0][good] - outs["ref"][0][good]).abs().max()), float((outs["ours"][2][good[..., None].expand(-1,-1,1)] - outs["ref"][2][good[..., None].expand(-1,-1,1)]).abs().max()))
