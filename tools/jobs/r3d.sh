set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_tile_step_gpu.py tests/test_render_gpu.py -x -q > gpurun_out/r3d_tests.log 2>&1; tail -5 gpurun_out/r3d_tests.log
timeout 300 python tools/sweep_decoder.py --out gpurun_out/r3d_decoder_sweep.json > gpurun_out/r3d_sweep.log 2>&1
grep "^{" gpurun_out/r3d_sweep.log
