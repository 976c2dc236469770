"""Run training steps / a render frame with torch's sync debug mode: every host synchronisation inside is reported."""
import sys, os, importlib, warnings, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
cfg = bench.WORKLOADS["c1-small"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 4, gen)]
step.step_device(*batches[0])
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    for b in batches[1:]:
        step.step_device(*b)
    n_train = len(w)
    for x in w[:10]:
        print("TRAIN SYNC:", str(x.message)[:200], x.filename, x.lineno)
print("train-step synchronisations:", n_train)
import render_frame as rf
ts = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, dev).finalize()
K = step.poses.ks[0].clone()
with torch.no_grad():
    c2w = step.poses.c2w()[0].detach()
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    rf.render_frame(ts, 135, 240, K, c2w)
    for x in w[:10]:
        print("RENDER SYNC:", str(x.message)[:200], x.filename, x.lineno)
    print("render-frame synchronisations:", len(w))
torch.cuda.set_sync_debug_mode("default")
