"""Count launches / CPU vs GPU time of one training step (debug aid)."""
import sys, os, time, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 8, gen)]
for b in batches[:4]:
    step.step_device(*b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for b in batches[4:]:
    step.step_device(*b)
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"4 steps: python issue time {t_cpu*250:.2f} ms/step, wall {t_all*250:.2f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step.step_device(*batches[0])
    torch.cuda.synchronize()
ev = prof.key_averages()
tot_cuda = sum(e.device_time_total for e in ev) / 1e3
n_k = sum(e.count for e in ev if e.device_time_total > 0 and e.cpu_time_total == 0)
print(f"profiled step: device time total {tot_cuda:.2f} ms; kernel launches ~{n_k}")
print(prof.key_averages().table(sort_by="device_time_total", row_limit=22, max_name_column_width=60))
