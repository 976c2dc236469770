#!/usr/bin/env python
"""profiles/traffic.json from an ncu CSV of one training step (tools/profile_step.py) taken with
  ncu --profile-from-start off --cache-control none --clock-control none \\
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv python tools/profile_step.py
usage: python tools/traffic_from_csv.py X.csv profiles/traffic.json "<source note>"
bench.py copies the per-class DRAM bytes into roofline*.traffic together with the source string."""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
CLASSES = {
    "encode_fwd": ["field_fwd_kernel"],
    "encode_bwd": ["field_geom_raygrad_kernel", "field_scatter_slice_kernel"],
    "adam_slices": ["adam_slice_kernel"],
    "encode_bwd_adam": ["field_geom_raygrad_kernel", "field_scatter_slice_kernel", "adam_slice_kernel"],
    "decoder": ["decoder_fwd4_kernel", "decoder_fwd_kernel", "decoder_bwd_fold_kernel", "decoder_bwd_kernel", "grad_absmax_kernel"],
}


def main():
    src, dst, note = sys.argv[1], sys.argv[2], sys.argv[3]
    per = collections.defaultdict(float)
    for row in csv.DictReader(l for l in open(src) if not l.startswith("==")):
        name = row.get("Metric Name", "")
        if not name.startswith("dram__bytes"):
            continue
        m = re.search(r"(\w+_kernel)", row["Kernel Name"])
        if not m:
            continue
        per[m.group(1)] += float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
    out = {"_comment": "DRAM bytes per training step and kernel class (sum over the launches of one step of the bench workload); "
                       "bench.py copies them into roofline*.traffic with this source string"}
    for cls, kernels in CLASSES.items():
        out[cls] = {"bytes": int(sum(per.get(k, 0.0) for k in kernels)), "source": note, "kernels": [k for k in kernels if k in per]}
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")
    print(json.dumps({k: v["bytes"] for k, v in out.items() if isinstance(v, dict)}))


if __name__ == "__main__":
    main()
