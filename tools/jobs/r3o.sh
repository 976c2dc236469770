set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_reference_drivers_gpu.py -x -q -k "allocation" > gpurun_out/r3o_tests.log 2>&1; tail -40 gpurun_out/r3o_tests.log
