set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
tail -3 gpurun_out/r3g_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3g_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"])
print("kernel_ms", d.get("kernel_ms"))
print("variants", json.dumps(d.get("variants"))[:1500])
print("ref", d.get("reference_cuda",{}).get("ms_per_step"), "render", d.get("render",{}).get("ms_per_frame"))
for k in ("roofline","roofline_fwd","roofline_decoder","roofline_update"):
    print(k, d[k].get("frac"), d[k].get("avg_launch_ms"))
PY
