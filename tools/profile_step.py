#!/usr/bin/env python
"""One training step of the bench workload between cudaProfilerStart / Stop, for
  ncu --profile-from-start off --set full -k regex:'field_|adam_slice|decoder_|composite' ... python tools/profile_step.py
SNRF_PROFILE_OVERLAP=0 runs the scatter / Adam slices serially (default here: ncu serialises kernels anyway)."""
import ctypes
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

pkg = importlib.import_module(bench.PKG)
pkg.install()
import scanerf_b200_capi as capi  # noqa: E402

cfg = dict(bench.WORKLOADS[os.environ.get("SNRF_PROFILE_WORKLOAD", "default.yaml-single-tile")])
if os.environ.get("SNRF_PROFILE_LOG2T"):          # the same workload on a smaller / larger table (L2-residency experiments)
    cfg["log2T"] = int(os.environ["SNRF_PROFILE_LOG2T"])
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
capi.lib().snrf_field_set_overlap(ctypes.c_int(int(os.environ.get("SNRF_PROFILE_OVERLAP", "0"))))
# SNRF_FWD_SPLIT=1: one encode-forward launch per level (per-level L2 hit rate); SNRF_FWD_L2="mode,pin_mib": its L2 policy
capi.lib().snrf_field_set_fwd_split_levels(ctypes.c_int(int(os.environ.get("SNRF_FWD_SPLIT", "0"))))
if os.environ.get("SNRF_FWD_L2"):
    _m, _p = (int(v) for v in os.environ["SNRF_FWD_L2"].split(","))
    capi.lib().snrf_field_set_fwd_l2_policy(ctypes.c_int(_m), ctypes.c_int(_p))
if os.environ.get("SNRF_FWD_PAIR"):
    _m, _p = (int(v) for v in os.environ["SNRF_FWD_PAIR"].split(","))
    capi.lib().snrf_field_set_fwd_pair_loads(ctypes.c_int(_m), ctypes.c_int(_p))
batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 4, gen)]
for b in batches[:3]:
    step.step_device(*b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step.step_device(*batches[3])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled step ok, loss", float(loss))
