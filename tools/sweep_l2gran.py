#!/usr/bin/env python
"""A/B of the device's L2 fetch granularity (cudaLimitMaxL2FetchGranularity) on the bench workload: training step, its
entry points, and one rendered frame (evidence for DESIGN 4.1; not a bench arm).
  python tools/sweep_l2gran.py [--out profiles/r2_l2gran_sweep.json]"""
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
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_l2gran_sweep.json"))
    ap.add_argument("--render", type=int, default=1)
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    lib = capi.lib()
    default = lib.snrf_l2_fetch_granularity(ctypes.c_int(0))
    print(json.dumps({"default_granularity": default}), flush=True)
    rows = [{"default_granularity": default}]
    names = ("snrf_field_encode_fwd", "snrf_field_encode_bwd_adam", "snrf_decoder_fwd", "snrf_decoder_bwd")
    for gran, fwd_mode in ((default, 0), (32, 0), (128, 0), (default, 0), (32, 0), (32, 1), (default, 1), (32, 1)):
        got = lib.snrf_l2_fetch_granularity(ctypes.c_int(gran))
        lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(fwd_mode), ctypes.c_int(0))
        ms, _ = bench._time_steps(step, batches, 4)
        capi.time_calls(names)
        for b in batches[:8]:
            step.step_device(*b)
        t = capi.timed_by_name()
        capi.time_calls(None)
        row = {"granularity_asked": gran, "granularity": got, "fwd_l2_mode": fwd_mode, "ms_per_step": ms}
        for k in names:
            row[k + "_ms"] = sum(t.get(k, [])) / max(len(t.get(k, [])), 1)
        if args.render:
            r = bench.bench_render(step, cfg, dev, frames=3, warm=1)
            row["render_ms_per_frame"] = r["ms_per_frame"]
            row["render_field_ms"] = r["roofline"]["ms_per_frame"]
        rows.append(row)
        print(json.dumps(row), flush=True)
    lib.snrf_l2_fetch_granularity(ctypes.c_int(default))
    lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(0), ctypes.c_int(0))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows) + "\n")


if __name__ == "__main__":
    main()
