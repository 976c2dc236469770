"""Time the field-encode kernels inside the real training step for several scatter pass counts."""
import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
import scanerf_b200_capi as capi
cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 6, gen)]
for b in batches[:3]:
    step.step_device(*b)
def measure(name):
    capi.time_calls(name)
    for b in batches[3:]:
        step.step_device(*b)
    ms, units = capi.timed_results()
    capi.time_calls(None)
    return sum(ms) / len(ms), [round(m, 3) for m in ms[:4]]
for run in (0, 2, 4, 8, 0, 4):
    capi.lib().snrf_field_set_run_length(capi.c_int(run))
    print("run_length", run, "bwd avg ms", *measure("snrf_field_encode_bwd"), flush=True)
capi.lib().snrf_field_set_run_length(capi.c_int(0))
for bits in (-1, 0, 1, 2):
    capi.lib().snrf_field_set_passes_log2(capi.c_int(bits))
    print("pass_bits", bits, "bwd avg ms", *measure("snrf_field_encode_bwd"))
capi.lib().snrf_field_set_passes_log2(capi.c_int(-1))
capi.lib().snrf_field_set_run_length(capi.c_int(0))
for agg in (0, 4, 8):
    capi.lib().snrf_field_set_aggregate_levels(capi.c_int(agg))
    print("aggregate_levels", agg, "bwd avg ms", *measure("snrf_field_encode_bwd"))
capi.lib().snrf_field_set_aggregate_levels(capi.c_int(-1))
capi.lib().snrf_field_set_run_length(capi.c_int(0))
print("fwd avg ms", *measure("snrf_field_encode_fwd"))
print("adam", *measure("snrf_adam_step"))
