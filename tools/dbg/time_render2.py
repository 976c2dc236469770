import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
import scanerf_b200_capi as capi
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
for two in (1, 0, 1, 0):
    capi.lib().snrf_infer_set_two_pass(capi.c_int(two))
    r = bench.bench_render(step, cfg, dev, frames=3, warm=1)
    print("two_pass", two, "ms/frame", round(r["ms_per_frame"], 1), flush=True)
from torch.profiler import profile, ProfilerActivity
import render_frame as rf
ts = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, dev).finalize()
K = step.poses.ks[0].clone(); K[0, 2] = 960; K[1, 2] = 540; K[0, 0] *= 2; K[1, 1] *= 2
c2w = step.poses.c2w()[0].detach()
capi.lib().snrf_infer_set_two_pass(capi.c_int(1))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    rf.render_frame(ts, 1080, 1920, K, c2w); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="device_time_total", row_limit=10, max_name_column_width=70))
