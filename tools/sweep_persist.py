#!/usr/bin/env python
"""Persisting-L2 access-policy window on the gradient scratch of the scatter + update fusion (snrf_field_set_persist_mib) against
the per-instruction evict_last hints, on the bench workload: ms per step (the forward pays for the set-aside), the library's
per-class timing of the fused backward, and the encode forward.  Evidence for DESIGN 4.1c; not a bench arm.
  python tools/sweep_persist.py [--out gpurun_out/persist_sweep.json]"""
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "persist_sweep.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    rt = ctypes.CDLL("libcudart.so.12")
    attrs = {}
    for name, aid in (("max_persisting_l2_bytes", 108), ("max_access_policy_window_bytes", 109), ("l2_bytes", 38)):
        v = ctypes.c_int(0)
        rt.cudaDeviceGetAttribute(ctypes.byref(v), ctypes.c_int(aid), ctypes.c_int(0))
        attrs[name] = v.value
    print(json.dumps(attrs), flush=True)
    rows = [attrs]
    lib = capi.lib()
    for persist, hints in ((0, 1), (64, 1), (64, 0), (72, 1), (48, 1), (0, 1), (64, 1)):
        lib.snrf_field_set_persist_mib(ctypes.c_int(persist))
        lib.snrf_field_set_l2_hints(ctypes.c_int(hints))
        ms, _ = bench._time_steps(step, batches, 4)
        row = {"persist_mib": persist, "l2_hints": hints, "ms_per_step": ms}
        capi.time_calls(("snrf_field_encode_fwd", "snrf_decoder_fwd", "snrf_decoder_bwd", "snrf_field_encode_bwd_adam"))
        for b in batches[:6]:
            step.step_device(*b)
        by = capi.timed_by_name()
        capi.time_calls(None)
        row["entry_ms"] = {k: sum(v) / len(v) for k, v in by.items()}
        lib.snrf_field_set_profile(ctypes.c_int(1))
        acc = [0.0] * 4
        for b in batches[:5]:
            step.step_device(*b)
            out4 = (ctypes.c_float * 4)()
            lib.snrf_field_last_profile(out4)
            acc = [a + v for a, v in zip(acc, out4)]
        lib.snrf_field_set_profile(ctypes.c_int(0))
        row["geom_raygrad_ms"], row["scatter_ms"], row["adam_ms"], row["scatter_and_adam_ms"] = [a / 5 for a in acc]
        rows.append(row)
        print(json.dumps(row), flush=True)
    lib.snrf_field_set_persist_mib(ctypes.c_int(0))
    lib.snrf_field_set_l2_hints(ctypes.c_int(1))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows, indent=1) + "\n")


if __name__ == "__main__":
    main()
