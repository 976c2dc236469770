#!/usr/bin/env python
"""A/B of encode-forward scheduling knobs on the bench workload (evidence for DESIGN 4.1; not a bench arm).
  python tools/sweep_fwd.py [--out profiles/r2_fwd_sweep.json]"""
import argparse
import ctypes
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_fwd_sweep.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    rows = []
    for pairing in (0, 1, 0, 1):
        capi.lib().snrf_field_set_fwd_pairing(ctypes.c_int(pairing))
        ms, _ = bench._time_steps(step, batches, 4)
        capi.time_calls(("snrf_field_encode_fwd",))
        for b in batches[:6]:
            step.step_device(*b)
        fwd = capi.timed_by_name().get("snrf_field_encode_fwd", [])
        capi.time_calls(None)
        row = {"fwd_pairing": pairing, "ms_per_step": ms, "encode_fwd_ms": sum(fwd) / max(len(fwd), 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    capi.lib().snrf_field_set_fwd_pairing(ctypes.c_int(0))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows) + "\n")


if __name__ == "__main__":
    main()
