cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r5h_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5h_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r5h_smoke.log
timeout 300 python bench.py --no-variants --no-render --no-cpu-baseline > gpurun_out/r5h_bench_short.json 2> gpurun_out/r5h_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r5h_bench_short.json
