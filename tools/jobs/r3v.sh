set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_composite_gpu.py tests/test_ert_gpu.py tests/test_tile_step_gpu.py tests/test_field_encode_gpu.py -x -q > gpurun_out/r3v_tests.log 2>&1; tail -3 gpurun_out/r3v_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-render --no-cpu-baseline > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3v_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"])
print({k:round(v,3) for k,v in d["kernel_ms"].items() if "composite" in k or "decoder" in k})
PY
