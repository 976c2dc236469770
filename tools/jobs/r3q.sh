set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3q_bench.json 2> gpurun_out/r3q_bench.err || exit 1
tail -2 gpurun_out/r3q_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3q_bench_reference.json 2> gpurun_out/r3q_bench_reference.err
cat gpurun_out/r3q_bench_reference.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r3q_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-variants --no-render --no-cpu-baseline > gpurun_out/r3q_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3q_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d.get("clocks"))
print("ref", d.get("reference_cuda",{}).get("ms_per_step"), "render", d.get("render",{}).get("ms_per_frame"))
print(json.dumps(d.get("variants"))[:900])
PY
