set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r4j_tests.log 2>&1; tail -3 gpurun_out/r4j_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4j_smoke.log 2>&1; tail -1 gpurun_out/r4j_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4j_bench.json 2> gpurun_out/r4j_bench.err || exit 1
python tools/profile_step.py > gpurun_out/r4j_plain.log 2>&1 || exit 1
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r4j_step_kernels_warm.csv python tools/profile_step.py > gpurun_out/r4j_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'decoder_bwd_fold_kernel|decoder_fwd4_kernel' -o gpurun_out/r4j_prof_dec -f python tools/profile_step.py > gpurun_out/r4j_ncu2.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4j_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d.get("clocks"))
print("kernel_ms", {k:round(v,3) for k,v in d.get("kernel_ms").items()})
print("ref", d.get("reference_cuda",{}).get("ms_per_step"), "render", d.get("render",{}).get("ms_per_frame"), d.get("render",{}).get("roofline",{}).get("frac"))
print(json.dumps(d.get("variants"))[:900])
for k in ("roofline","roofline_fwd","roofline_decoder","roofline_update","roofline_bwd_and_update"):
    print(k, d[k].get("frac"), d[k].get("avg_launch_ms"))
PY
