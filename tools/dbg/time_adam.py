import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
import scanerf_b200_capi as capi
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 9, gen)]
for b in batches[:3]:
    step.step_device(*b)
for name in ("snrf_adam_step", "snrf_field_encode_fwd", "snrf_field_encode_bwd", "snrf_decoder_fwd", "snrf_decoder_bwd"):
    capi.time_calls(name)
    for b in batches[3:]:
        step.step_device(*b)
    ms, _ = capi.timed_results()
    capi.time_calls(None)
    print(name, "avg ms", round(sum(ms) / len(ms), 3), [round(m, 3) for m in ms[:6]], flush=True)
