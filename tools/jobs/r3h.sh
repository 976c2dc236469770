set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_tests.log 2>&1; tail -15 gpurun_out/r3h_tests.log
