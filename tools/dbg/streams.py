import sys, os, time, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 40, gen)]
def loop(bs, sync_each=False):
    torch.cuda.synchronize()
    r0 = torch.cuda.memory_reserved(); n0 = torch.cuda.memory_stats()["num_device_alloc"] if "num_device_alloc" in torch.cuda.memory_stats() else torch.cuda.memory_stats().get("segment.all.allocated", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in bs:
        step.step_device(*b)
        if sync_each: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    n1 = torch.cuda.memory_stats().get("num_device_alloc", torch.cuda.memory_stats().get("segment.all.allocated", 0))
    return round(e0.elapsed_time(e1) / len(bs), 2), (torch.cuda.memory_reserved() - r0) >> 20, n1 - n0
for two in (True, False, True, False):
    step.two_streams = two
    loop(batches[:10])
    print("two_streams", two, "async", loop(batches[10:30]), "sync-each", loop(batches[10:30], True), "reserved GiB", torch.cuda.memory_reserved() >> 30)
