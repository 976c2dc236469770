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
    cfg = dict(bench.WORKLOADS["default.yaml-single-tile"])
    if os.environ.get("SNRF_SWEEP_LOG2T"):            # the same workload on a smaller table (e.g. 22: the city tiles)
        cfg["log2T"] = int(os.environ["SNRF_SWEEP_LOG2T"])
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    rows = []
    lib = capi.lib()
    # (L2 policy mode, pin MiB, x-pair load mode, first level of the pair loads), interleaved with the default
    variants = [(-1, 0, 0, 0), (0, 0, 0, 0), (1, 0, 0, 0), (0, 0, 0, 0), (0, 0, 2, 9)]
    for mode, pin, pair, first in variants:
        if mode >= 0:          # (-1: the library's defaults, untouched)
            lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(mode), ctypes.c_int(pin))
            lib.snrf_field_set_fwd_pair_loads(ctypes.c_int(pair), ctypes.c_int(first))
        ms, _ = bench._time_steps(step, batches, 4)
        capi.time_calls(("snrf_field_encode_fwd", "snrf_field_encode_bwd_adam", "snrf_decoder_fwd"))
        for b in batches[:8]:
            step.step_device(*b)
        t = capi.timed_by_name()
        capi.time_calls(None)
        mean = lambda k: sum(t.get(k, [])) / max(len(t.get(k, [])), 1)
        row = {"l2_mode": mode, "pin_mib": pin, "pair_loads": pair, "pair_first_level": first, "ms_per_step": ms,
               "encode_fwd_ms": mean("snrf_field_encode_fwd"), "bwd_adam_ms": mean("snrf_field_encode_bwd_adam"),
               "decoder_fwd_ms": mean("snrf_decoder_fwd")}
        rows.append(row)
        print(json.dumps(row), flush=True)
    lib.snrf_field_set_fwd_pair_loads(ctypes.c_int(0), ctypes.c_int(0))
    lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(0), ctypes.c_int(0))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows) + "\n")


if __name__ == "__main__":
    main()
