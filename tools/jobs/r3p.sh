set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3p_smoke.log 2>&1; tail -2 gpurun_out/r3p_smoke.log
timeout 900 python -m pytest tests/test_tile_step_gpu.py tests/test_sync_free_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/r3p_tests.log 2>&1; tail -4 gpurun_out/r3p_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-render --no-cpu-baseline > gpurun_out/r3p_bench.json 2> gpurun_out/r3p_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3p_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"])
PY
