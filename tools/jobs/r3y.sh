set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_field_encode_gpu.py tests/test_adam_gpu.py tests/test_tile_step_gpu.py tests/test_ert_gpu.py -x -q > gpurun_out/r3y_tests.log 2>&1; tail -3 gpurun_out/r3y_tests.log
timeout 600 python tools/sweep_fused.py --out gpurun_out/r3y_fused_sweep.json 2>&1 | grep "^{"
