set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_ert_gpu.py tests/test_composite_gpu.py tests/test_field_encode_gpu.py tests/test_tile_step_gpu.py tests/test_decoder_gpu.py -x -q > gpurun_out/r3f_tests.log 2>&1; tail -30 gpurun_out/r3f_tests.log
