set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3k_smoke.log 2>&1; tail -2 gpurun_out/r3k_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err
tail -3 gpurun_out/r3k_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3k_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
print("kernel_ms", {k:round(v,3) for k,v in d.get("kernel_ms").items()})
print("variants", json.dumps(d.get("variants"))[:1200])
print("ref", d.get("reference_cuda",{}).get("ms_per_step"), d.get("reference_cuda",{}).get("ratio"), "render", d.get("render",{}).get("ms_per_frame"), d.get("render",{}).get("roofline",{}).get("frac"))
for k in ("roofline","roofline_fwd","roofline_decoder","roofline_update","roofline_bwd_and_update"):
    print(k, d[k].get("frac"), d[k].get("avg_launch_ms"))
print("cpu", d.get("cpu_baseline"))
PY
