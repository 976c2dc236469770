set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_ert_gpu.py -x -q > gpurun_out/r3m_tests.log 2>&1; tail -3 gpurun_out/r3m_tests.log
timeout 300 python tools/sweep_decoder.py --out gpurun_out/r3m_decoder_sweep.json > gpurun_out/r3m_sweep.log 2>&1
grep "^{" gpurun_out/r3m_sweep.log | cut -c1-330
