import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
import scanerf_b200_capi as capi
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
for two in (1, 2, 1, 2):
    capi.lib().snrf_infer_set_two_pass(capi.c_int(two))
    r = bench.bench_render(step, cfg, dev, frames=3, warm=1)
    print("two_pass", two, "ms/frame", round(r["ms_per_frame"], 1), flush=True)
import render_frame as rf
ts = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, dev).finalize()
K = step.poses.ks[0].clone(); K[0, 2] = 960; K[1, 2] = 540; K[0, 0] *= 2; K[1, 1] *= 2
c2w = step.poses.c2w()[0].detach()
o, d = rf.pinhole_rays(1080, 1920, K, c2w, dev)
from hashgrid.lib import HASHGRID as ops
B, S = o.shape[0], 128
isect = torch.full((B, 1, 2), 1e7, device=dev)
ops.ray_block_intersection(o, d, ts.block_corner, ts.block_size, isect)
order = torch.zeros(B, 1, dtype=torch.int32, device=dev)
z, di = torch.full((B, S), -1.0, device=dev), torch.full((B, S), -1.0, device=dev)
ti, zs = torch.zeros(B, 1, dtype=torch.int32, device=dev), torch.zeros(B, 1, device=dev)
ops.sample_points(o, d, ts.block_corner, ts.block_size, ts.fake_occupied_grid, ts.grid_starts, ts.grid_log2dim, order, isect, ti, zs, z, di)
print("foreground samples placed:", float((z != -1).float().mean()))
