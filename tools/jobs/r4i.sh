set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_tile_step_gpu.py tests/test_render_gpu.py -x -q > gpurun_out/r4i_tests.log 2>&1; tail -3 gpurun_out/r4i_tests.log
timeout 300 python tools/decoder_overheads.py --out gpurun_out/r4i_decoder_overheads.json 2>&1 | grep "^{"
