import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 5, gen)]
for b in batches[:4]:
    step.step_device(*b)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
    step.step_device(*batches[4]); torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.key.startswith("aten::")]
rows.sort(key=lambda r: -r[1])
for k, c, t in rows[:40]:
    print(f"{k:45s} calls {c:4d}  cuda us {t:9.1f}")
print("kernels launched:", sum(e.count for e in prof.key_averages() if e.device_type.name == "CUDA"))
