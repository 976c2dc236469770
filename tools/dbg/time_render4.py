import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
import scanerf_b200_capi as capi
import render_frame as rf
from hashgrid._decoder import flatten_for_inference
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
hg, ts = step.featureGrid, rf.TileSet(dev)
flat = flatten_for_inference(step.decoder)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for j in range(NT):
    shift = torch.tensor([0.9 * 0.5 * float(hg.bbox_size[0]) * j, 0.0, 0.0], device=hg.min_bbox.device)
    ts.add_tile(hg.HE.features.detach().half().roll(j, 1), flat * (1.0 + 0.01 * j), hg.HE.resolution, hg.occupied_grid,
                hg.min_bbox + shift, hg.bbox_size, hg.sampler_log2dim)
ts.finalize()
K = step.poses.ks[0].clone(); K[0, 2] = 960; K[1, 2] = 540; K[0, 0] *= 2; K[1, 1] *= 2
c2w = step.poses.c2w()[0].detach()
from torch.profiler import profile, ProfilerActivity
with torch.no_grad():
    rf.render_frame(ts, 1080, 1920, K, c2w); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        rf.render_frame(ts, 1080, 1920, K, c2w); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="device_time_total", row_limit=16, max_name_column_width=60))
