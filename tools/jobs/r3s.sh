set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_warp_loss_gpu.py -x -q > gpurun_out/r3s_tests.log 2>&1; tail -12 gpurun_out/r3s_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-render --no-cpu-baseline > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err
tail -2 gpurun_out/r3s_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3s_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"])
print(json.dumps(d.get("variants"))[:1200])
PY
